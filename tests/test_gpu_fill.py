"""GPU parity tests for the single-pair path: engine (through the C ABI) vs the CPU oracle and the golden
vectors generated from the reference's cpu4 + NwTrace1_Plain.  Bit-exact: integer score, run-length
transcript, trace hash."""
import numpy as np
import pytest

from conftest import case_letters

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(scoring):
    from gpuseqalign_b200 import Engine
    e = Engine(0)
    e.set_scoring(scoring["subst"]["blosum62"], -11)
    yield e
    e.close()


def test_golden_scores(engine, golden):
    for c in golden["cases"]:
        y, x = case_letters(golden, c)
        assert engine.align(y, x, keep_headers=False) == c["score"], (c["y"], c["x"])


def test_golden_scores_i32_entry(engine, golden):
    for c in golden["cases"][::7]:
        y, x = case_letters(golden, c)
        sy = np.concatenate([[0], y]).astype(np.int32)
        sx = np.concatenate([[0], x]).astype(np.int32)
        assert engine.align_i32(sy, sx, keep_headers=True) == c["score"]


def test_golden_traces(engine, golden):
    """All 206 reference cases (pair_debug + pair_generated_1 + substring ranges): score, transcript, hash."""
    for c in golden["cases"]:
        y, x = case_letters(golden, c)
        assert engine.align(y, x, keep_headers=True) == c["score"], (c["y"], c["x"])
        edit, th = engine.trace()
        assert edit == c["edit"], (c["y"], c["x"])
        assert f"{th:08x}" == c["trace_hash"], (c["y"], c["x"])


VARIANTS = [(4, 4, 2, 512), (4, 4, 1, 512), (8, 4, 2, 256), (8, 4, 1, 96), (16, 4, 2, 32), (16, 4, 1, 512),
            (4, 1, 2, 64), (8, 1, 2, 1024), (4, 4, 2, 1024), (8, 4, 2, 64)]


@pytest.mark.parametrize("R,K,Bx", [(4, 2, 32), (8, 2, 64), (4, 1, 96), (16, 2, 128)])
def test_segmented_maps_follow_paths_across_cuts(scoring, oracle, R, K, Bx):
    """The separate map kernel (pairs with more bands than the fused launch shadows) computes a band's map in independent column
    segments that resume from the fill's snapshots; a path that leaves a segment through its left cut is followed through the
    cut labels.  Small snapshot spacings + long horizontal runs (y much shorter than x, and a 3000-column insertion) make
    paths cross many cuts inside one band."""
    import ctypes as C
    from gpuseqalign_b200 import Engine, Params, synth
    subst = scoring["subst"]["blosum62"]
    e = Engine(0)
    e.set_scoring(subst, -11)
    e._L.nwb200_debug_band_stamps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    e._L.nwb200_debug_band_stamps(e._h, 4, 0, None, 0)          # origin maps in their own launch
    x0 = synth.letters(91, 2600)
    cases = [(synth.letters(92, 700), synth.letters(93, 5000)),                                        # 7 columns per row
             (synth.mutated_copy(x0, 94, 2500), np.concatenate([x0[:1200], synth.letters(95, 3000), x0[1200:]])),
             (synth.letters(96, 1500), synth.letters(97, 1400))]
    for y, x in cases:
        exp = oracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
        assert e.align(y, x, params=Params(R, 4, Bx, K)) == exp.score
        edit, th = e.trace()
        assert edit == exp.edit and th == exp.trace_hash, (R, K, Bx, y.size, x.size)
    e.close()


def test_align_with_trace_flag(engine, golden, scoring, oracle):
    """NWB200_WITH_TRACE: the align call enqueues traceback + move copy itself; trace() only formats.  Same results, also when
    the two forms are mixed on one context and for multi-band pairs."""
    from gpuseqalign_b200 import synth
    for c in golden["cases"][::7]:
        y, x = case_letters(golden, c)
        assert engine.align(y, x, with_trace=True) == c["score"], (c["y"], c["x"])
        edit, th = engine.trace()
        assert edit == c["edit"] and f"{th:08x}" == c["trace_hash"], (c["y"], c["x"])
    subst = scoring["subst"]["blosum62"]
    x = synth.letters(71, 3000)
    y = synth.mutated_copy(x, 72, 2900)
    exp = oracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
    for with_trace in (True, False, True):
        assert engine.align(y, x, with_trace=with_trace) == exp.score
        edit, th = engine.trace()
        assert edit == exp.edit and th == exp.trace_hash
        assert engine.trace() == (edit, th)         # a second call returns the cached transcript


@pytest.mark.parametrize("R,W,K,Bx", VARIANTS)
def test_random_shapes_all_kernel_variants(engine, scoring, oracle, R, W, K, Bx):
    from gpuseqalign_b200 import Params
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(R * 100 + W * 10 + K)
    By = R * 32
    shapes = [(1, 1), (1, 40), (5, 3), (31, 33), (By - 1, By + 1), (By, 2 * By), (By + 1, 777), (2 * By + 17, 3 * By + 5),
              (1500, 4100), (2100, 130), (3 * By, 1)]
    for n, m in shapes:
        y = rng.integers(0, 20, n).astype(np.uint8)
        x = rng.integers(0, 20, m).astype(np.uint8)
        exp = oracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
        got = engine.align(y, x, keep_headers=True, params=Params(R, W, Bx, K))
        assert got == exp.score, (n, m)
        assert (engine.info.tile_rows, engine.info.tile_cols) == (By, Bx)
        edit, th = engine.trace()
        assert edit == exp.edit, (n, m)
        assert th == exp.trace_hash, (n, m)


def test_similar_sequences_long(engine, scoring, oracle):
    """A mutated copy (mostly diagonal path), an unrelated pair, and long gap runs at 5000 x 6000."""
    from gpuseqalign_b200 import synth
    subst = scoring["subst"]["blosum62"]
    x = synth.letters(11, 6000)
    y = synth.mutated_copy(x, 12, 5000)
    y2 = synth.letters(13, 5000)
    y3 = np.concatenate([x[:1000], x[3000:5500]])          # a 2000-column deletion: one long horizontal run
    for yy in (y, y2, y3):
        exp = oracle.align_pair(yy, x, subst, -11, want_hash=False, want_trace=True)
        assert engine.align(yy, x, keep_headers=True) == exp.score
        edit, th = engine.trace()
        assert edit == exp.edit and th == exp.trace_hash
    # rows longer than columns (the reference always has X the longer one; the engine does not care)
    exp = oracle.align_pair(x, y2[:700], subst, -11, want_hash=False, want_trace=True)
    assert engine.align(x, y2[:700], keep_headers=True) == exp.score
    edit, th = engine.trace()
    assert edit == exp.edit and th == exp.trace_hash


def test_resident_split_form_repeats(engine, scoring, oracle):
    """upload once, fill + trace many times (the benchmark's device-resident loop); epochs must not leak."""
    from gpuseqalign_b200 import synth
    subst = scoring["subst"]["blosum62"]
    x = synth.letters(21, 3000); y = synth.letters(22, 2500)
    exp = oracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
    engine.upload_pair(y, x)
    for _ in range(5):
        engine.fill_resident(True)
        engine.trace_resident()
        assert engine.fetch_score() == exp.score
        edit, th = engine.fetch_trace()
        assert edit == exp.edit and th == exp.trace_hash


def test_other_scoring(scoring, oracle):
    from gpuseqalign_b200 import Engine
    e = Engine(0)
    try:
        rng = np.random.default_rng(5)
        y = rng.integers(0, 25, 700).astype(np.uint8)
        x = rng.integers(0, 25, 900).astype(np.uint8)
        for name, gap in [("blosum45", -5), ("blosum90", -20), ("blosum50", -1), ("blosum80", 0)]:
            subst = scoring["subst"][name]
            e.set_scoring(subst, gap)
            exp = oracle.align_pair(y, x, subst, gap, want_hash=False, want_trace=True)
            assert e.align(y, x, keep_headers=True) == exp.score, (name, gap)
            edit, th = e.trace()
            assert edit == exp.edit and th == exp.trace_hash, (name, gap)
    finally:
        e.close()


def test_error_behaviour(engine):
    from gpuseqalign_b200 import NwB200Error, NwStat, Params
    with pytest.raises(NwB200Error) as ei:
        engine.align(np.array([99], dtype=np.uint8), np.array([1, 2], dtype=np.uint8))
    assert ei.value.stat == NwStat.errorInvalidValue
    with pytest.raises(NwB200Error) as ei:
        engine.align(np.array([1], dtype=np.uint8), np.array([1, 2], dtype=np.uint8), params=Params(rows_per_lane=3))
    assert ei.value.stat == NwStat.errorInvalidValue
    with pytest.raises(NwB200Error):
        engine.align(np.array([], dtype=np.uint8), np.array([1, 2], dtype=np.uint8))
    engine.align(np.array([1], dtype=np.uint8), np.array([1, 2], dtype=np.uint8), keep_headers=False)
    with pytest.raises(NwB200Error) as ei:          # traceback needs the headers of the last fill
        engine.trace()
    assert ei.value.stat == NwStat.errorInvalidValue


def test_golden_score_hash(engine, golden):
    """score_hash (NwHash1_Plain / NwHash2_Sparse value): every cell recomputed on the GPU, folded by the host."""
    for c in golden["cases"][::4]:
        y, x = case_letters(golden, c)
        engine.align(y, x, keep_headers=False)
        assert f"{engine.score_hash():08x}" == c["score_hash"], (c["y"], c["x"])


def test_score_hash_multi_band_and_variants(engine, scoring, oracle):
    from gpuseqalign_b200 import Params, synth
    subst = scoring["subst"]["blosum62"]
    x = synth.letters(31, 1900); y = synth.letters(32, 1333)
    exp = oracle.align_pair(y, x, subst, -11, want_hash=True, want_trace=False)
    for p in (None, Params(8, 4, 256, 1), Params(16, 4, 64, 2), Params(4, 1, 32, 2)):
        assert engine.align(y, x, keep_headers=True, params=p) == exp.score
        assert engine.score_hash() == exp.score_hash


@pytest.mark.parametrize("kind", ["mutated", "long_indel", "random"])
def test_corridor_maps_hit_miss_and_off(scoring, oracle, monkeypatch, kind):
    """Origin maps restricted to a corridor around the line to the origin (long pairs): a path inside the corridor is found by the
    corridor pass alone, a path that leaves it (a 2 500-letter deletion in the middle) sets the miss flag and the full pass that is
    enqueued behind takes over; same transcript either way, and with the corridor switched off."""
    from gpuseqalign_b200 import Engine, synth
    subst = scoring["subst"]["blosum62"]
    n = 40000                                  # (313 bands: more than the fill launch can host map units for, so pass A is the separate kernel)
    x = synth.letters(611, n)
    if kind == "mutated":
        y = synth.mutated_copy(x, 612, n)
    elif kind == "long_indel":
        y = np.concatenate([x[:19000], x[21500:], synth.letters(613, 2500)])      # the path runs 2 500 columns off the line in the middle
    else:
        y = synth.letters(614, n)
    exp = oracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
    seen = {}
    for d in ("512", "0", None):
        if d is None: monkeypatch.delenv("NWB200_CORRIDOR", raising=False)
        else: monkeypatch.setenv("NWB200_CORRIDOR", d)
        eng = Engine(0)
        try:
            eng.set_scoring(subst, -11)
            assert eng.align(y, x, keep_headers=True) == exp.score
            assert eng.trace() == (exp.edit, exp.trace_hash), (kind, d)
            seen[d] = eng.trace_info()
        finally:
            eng.close()
    assert seen["512"]["corridor_segments"] > 0 and seen["512"]["corridor_segments"] * 2 <= seen["512"]["segments"]
    assert seen["512"]["corridor_missed"] == (kind == "long_indel")
    assert seen["0"]["corridor_segments"] == 0 and not seen["0"]["corridor_missed"]
    assert not seen[None]["corridor_missed"] or kind == "long_indel"


def _path_cells(edit: str, n: int, m: int):
    """Cells of the path described by a run-length transcript, top-left -> bottom-right."""
    import re
    i = j = 0
    cells = [(0, 0)]
    for cnt, op in re.findall(r"(\d*)([=XID])", edit):
        for _ in range(int(cnt) if cnt else 1):
            if op in "=X": i += 1; j += 1
            elif op == "I": i += 1
            else: j += 1
            cells.append((i, j))
    assert (i, j) == (n, m)
    return cells


def test_score_rows_path_values_and_memory_usage(engine, scoring, oracle):
    """The calls behind the plugin's NwPrintScore / calcDebugTrace / peak-memory columns: rows of the full score matrix recomputed
    on the GPU (nwtrace2_sparse.cpp:346-419), the matrix values along the path (nwtrace1_plain.cpp:34-38,107,120-126), and the
    launch resources (nwalign_shared.cpp:5-25)."""
    from gpuseqalign_b200 import Params, synth
    subst = scoring["subst"]["blosum62"]
    for (n, m, p) in ((1333, 1900, None), (77, 3, None), (1, 1, None), (700, 2100, Params(4, 1, 32, 2)), (2500, 301, Params(8, 4, 256, 1))):
        x = synth.letters(900 + n, m); y = synth.mutated_copy(x, 901 + m, n) if n > 3 else synth.letters(5, n)
        H = oracle.fill_full(y, x, subst, -11)
        assert engine.align(y, x, keep_headers=True, params=p) == H[-1, -1]
        assert np.array_equal(engine.score_rows(0, n + 1, m + 1), H)
        for (r0, k) in ((0, 1), (n, 1), (n // 2, min(3, n + 1 - n // 2)), (1, 0)):
            assert np.array_equal(engine.score_rows(r0, k, m + 1), H[r0:r0 + k])
        edit, th = engine.trace()
        vals = engine.trace_values()
        cells = _path_cells(edit, n, m)
        assert np.array_equal(vals, np.array([H[i, j] for i, j in cells], dtype=np.int32))
        assert engine.trace() == (edit, th)                                  # the value pass leaves the cached transcript intact
        mu = engine.memory_usage()
        assert mu["device_bytes"] > 0 and mu["regs_per_thread"] > 0 and mu["blocks"] > 0
        assert mu["register_bytes"] == mu["regs_per_thread"] * 4 * mu["threads_per_block"] * mu["blocks"]
        assert mu["shared_bytes"] > 0
    with pytest.raises(Exception):
        engine.score_rows(n, 5, m + 1)


@pytest.mark.parametrize("R,Bx", [(4, 64), (8, 96), (4, 512)])
def test_exported_headers_feed_reference_style_trace(engine, golden, scoring, oracle, R, Bx):
    """nwb200_copy_headers: headers in the reference's tile layout must let NwTrace2_Sparse (restated, and the live
    reference when oracle/_ref was prebuilt) reproduce the golden transcript: validates every consumed header value."""
    from gpuseqalign_b200 import Params
    subst = scoring["subst"]["blosum62"]
    for c in golden["cases"][::6]:
        y, x = case_letters(golden, c)
        score = engine.align(y, x, keep_headers=True, params=Params(R, 4, Bx, 2))
        assert score == c["score"]
        hrow, hcol = engine.headers()
        info = engine.info
        r = oracle.trace_sparse(hrow, hcol, info.tile_rows, info.tile_cols, y, x, subst, -11)
        assert r.score == c["score"]
        assert r.edit == c["edit"], (c["y"], c["x"])
        if oracle.ref_available() and c["len_y"] > 100:
            r2 = oracle.ref_trace_from_headers(y, x, subst, -11, hrow, hcol, info.tile_rows, info.tile_cols)
            assert (r2.score, r2.edit, f"{r2.trace_hash:08x}") == (c["score"], c["edit"], c["trace_hash"])
        # the engine's own traceback still works after an export
        edit, th = engine.trace()
        assert edit == c["edit"]


@pytest.mark.parametrize("R,block", [(4, 256), (8, 64), (4, 2048), (16, 32)])
def test_column_block_wavefront_loopback(engine, scoring, oracle, R, block):
    """The cross-GPU wavefront code path with world = 1: the blocks hand their border columns to each other through the
    GPU's own receive buffer (same kernel, same flags as over NVLink)."""
    from gpuseqalign_b200 import Params, synth
    from gpuseqalign_b200.wavefront import wave_align
    subst = scoring["subst"]["blosum62"]
    for n, m, seed in [(300, 5000, 41), (1000, 777, 43), (1, 100, 45), (130, 33, 47)]:
        y = synth.letters(seed, n); x = synth.letters(seed + 1, m)
        exp, _, _, _ = oracle.fill_rolling(y, x, subst, -11)
        assert wave_align(engine, y, x, block_cols=block, epoch=seed, params=Params(R, 4, 0, 2)) == exp, (n, m)
