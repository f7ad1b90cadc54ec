"""GPU parity tests for the batch path (BASELINE config 3): scores of many short pairs vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(scoring):
    from gpuseqalign_b200 import Engine
    e = Engine(0)
    e.set_scoring(scoring["subst"]["blosum62"], -11)
    yield e
    e.close()


def _ragged(rng, n_pairs, max_y, max_x, alphabet=20):
    lenY = rng.integers(0, max_y + 1, n_pairs).astype(np.uint32)
    lenX = rng.integers(0, max_x + 1, n_pairs).astype(np.uint32)
    lens = np.empty(2 * n_pairs, dtype=np.uint64)
    lens[0::2] = lenY; lens[1::2] = lenX
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    letters = rng.integers(0, alphabet, int(offs[-1]) + 1).astype(np.uint8)
    return letters, offs[0:-1:2].copy(), lenY, offs[1::2].copy(), lenX


def test_batch_256_synthetic(engine, scoring, oracle):
    from gpuseqalign_b200 import synth
    subst = scoring["subst"]["blosum62"]
    pool, offY, lenY, offX, lenX = synth.batch_pairs(0, 4096, 256, 256)
    got = engine.align_batch(pool, offY, lenY, offX, lenX)
    exp = oracle.score_batch(pool, offY, lenY, offX, lenX, subst, -11)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("max_y,max_x", [(128, 300), (256, 256), (512, 100), (700, 700)])
def test_batch_ragged(engine, scoring, oracle, max_y, max_x):
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(max_y + max_x)
    letters, offY, lenY, offX, lenX = _ragged(rng, 600, max_y, max_x)
    got = engine.align_batch(letters, offY, lenY, offX, lenX)
    exp = oracle.score_batch(letters, offY, lenY, offX, lenX, subst, -11)
    assert np.array_equal(got, exp)


def test_batch_with_transcripts(engine, scoring, oracle):
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(3)
    letters, offY, lenY, offX, lenX = _ragged(rng, 40, 300, 300)
    scores, edits, hashes = engine.align_batch(letters, offY, lenY, offX, lenX, want_trace=True)
    for p in range(40):
        y = letters[int(offY[p]): int(offY[p]) + int(lenY[p])]
        x = letters[int(offX[p]): int(offX[p]) + int(lenX[p])]
        if y.size == 0 or x.size == 0:
            assert scores[p] == -11 * (y.size + x.size)
            continue
        exp = oracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
        assert (scores[p], edits[p], hashes[p]) == (exp.score, exp.edit, exp.trace_hash), p


def test_batch_resident_split_form(engine, scoring, oracle):
    from gpuseqalign_b200 import synth
    subst = scoring["subst"]["blosum62"]
    pool, offY, lenY, offX, lenX = synth.batch_pairs(100, 1000, 256, 256)
    engine.upload_batch(pool, offY, lenY, offX, lenX)
    exp = oracle.score_batch(pool, offY, lenY, offX, lenX, subst, -11)
    for _ in range(3):
        engine.batch_resident()
        assert np.array_equal(engine.fetch_batch_scores(), exp)


def test_batch_rejects_letters_outside_the_alphabet(engine):
    from gpuseqalign_b200 import NwB200Error, NwStat, synth
    pool, offY, lenY, offX, lenX = synth.batch_pairs(0, 50, 64, 64)
    pool = pool.copy(); pool[1234] = 77
    with pytest.raises(NwB200Error) as ei:
        engine.align_batch(pool, offY, lenY, offX, lenX)
    assert ei.value.stat == NwStat.errorInvalidValue
    bad_off = offY.copy(); bad_off[3] = pool.size
    with pytest.raises(NwB200Error):
        engine.align_batch(pool, bad_off, lenY, offX, lenX)


def test_batch_pinned_and_pageable_sources_agree(engine, scoring, oracle):
    """The one-shot call pipelines H2D slices with the kernel; pinned memory goes by DMA, pageable through a staging ring."""
    import torch
    from gpuseqalign_b200 import synth
    subst = scoring["subst"]["blosum62"]
    pool, offY, lenY, offX, lenX = synth.batch_pairs(7, 140000, 256, 256)        # 71 MB: several slices
    exp = oracle.score_batch(pool[: 4000 * 512], offY[:4000], lenY[:4000], offX[:4000], lenX[:4000], subst, -11)
    a = engine.align_batch(pool, offY, lenY, offX, lenX)
    pinned = torch.from_numpy(pool).pin_memory().numpy()
    b = engine.align_batch(pinned, offY, lenY, offX, lenX)
    assert np.array_equal(a, b) and np.array_equal(a[:4000], exp)
    assert int(a.astype(np.int64).sum()) == int(b.astype(np.int64).sum())


@pytest.mark.parametrize("n_pairs,max_y,max_x", [(1, 256, 256), (7, 128, 200), (601, 256, 300), (333, 100, 60), (90, 500, 300)])
def test_batch_packed_and_32bit_kernels_agree(engine, scoring, oracle, monkeypatch, n_pairs, max_y, max_x):
    """Two pairs per warp in 16-bit halves (nw_batch2.cuh) vs one pair per warp (nw_batch.cuh) vs the oracle: odd pair
    counts (a warp whose second half is empty), both band heights, ragged partners in one warp."""
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(1000 + n_pairs)
    letters, offY, lenY, offX, lenX = _ragged(rng, n_pairs, max_y, max_x)
    exp = oracle.score_batch(letters, offY, lenY, offX, lenX, subst, -11)
    monkeypatch.setenv("NWB200_BATCH_PACKED", "1")
    packed = engine.align_batch(letters, offY, lenY, offX, lenX)
    assert np.array_equal(packed, exp)
    # every packed instance: round 1's kernel (0), two pairs per warp (1: the 32-lane groups), mixed IDP routes (2), K = 2 (3),
    # four pairs per warp (4: 16-lane groups, 16 rows per lane; 5: with K = 2; 6: split lanes), expanded profiles (7)
    for v in ("0", "1", "2", "3", "4", "5", "6", "7"):
        monkeypatch.setenv("NWB200_BATCH_VARIANT", v)
        assert np.array_equal(engine.align_batch(letters, offY, lenY, offX, lenX), exp), v
    monkeypatch.delenv("NWB200_BATCH_VARIANT")
    monkeypatch.setenv("NWB200_BATCH_PACKED", "0")
    plain = engine.align_batch(letters, offY, lenY, offX, lenX)
    assert np.array_equal(plain, exp)
    # the 32-bit kernel with its column letters staged by TMA bulk copies (ragged offsets: most pairs are not 16-byte aligned and take
    # the LDG path inside the same launch; the aligned synthetic batch below takes the bulk path for every pair)
    monkeypatch.setenv("NWB200_BATCH_TMA", "1")
    assert np.array_equal(engine.align_batch(letters, offY, lenY, offX, lenX), exp)
    from gpuseqalign_b200 import synth
    pool, oY, lY, oX, lX = synth.batch_pairs(40 + n_pairs, 300, 256, 256)
    assert np.array_equal(engine.align_batch(pool, oY, lY, oX, lX), oracle.score_batch(pool, oY, lY, oX, lX, subst, -11))


def test_batch_packed_halves_at_their_bound(scoring, oracle):
    """s' = 127 on the diagonal and identical 256-letter sequences: P reaches 256 * 127 = 32 512 in both halves of a register;
    one notch more (s' = 128) must take the 32-bit kernel and still agree."""
    from gpuseqalign_b200 import Engine
    S = 25
    rng = np.random.default_rng(77)
    seqs = [rng.integers(0, 20, 256).astype(np.uint8) for _ in range(3)]
    # pairs: (s0,s0) (s1,s1) identical; (s0,s1) unrelated; (s2,s2) with an empty partner in the last warp
    pool = np.concatenate([seqs[0], seqs[0], seqs[1], seqs[1], seqs[0], seqs[1], seqs[2], seqs[2]])
    offY = np.arange(0, 8, 2, dtype=np.uint64) * 256
    offX = offY + 256
    lenY = np.full(4, 256, dtype=np.uint32); lenX = lenY.copy()
    for diag in (105, 106):
        subst = rng.integers(-4, 4, (S, S)).astype(np.int32)
        subst = ((subst + subst.T) // 2).astype(np.int32)
        np.fill_diagonal(subst, diag)
        e = Engine(0)
        try:
            e.set_scoring(subst.ravel(), -11)
            got = e.align_batch(pool, offY, lenY, offX, lenX)
        finally:
            e.close()
        exp = oracle.score_batch(pool, offY, lenY, offX, lenX, subst.ravel(), -11)      # the oracle takes the flat S*S table
        assert np.array_equal(got, exp), diag
        assert got[0] == 256 * diag


@pytest.mark.parametrize("n_pairs,max_y,max_x", [(300, 256, 256), (200, 128, 400), (64, 512, 512), (50, 40, 3000), (24, 300, 2000), (10, 700, 300)])
def test_batch_transcripts_kernel(engine, scoring, oracle, n_pairs, max_y, max_x):
    """Transcripts of a whole batch come from one warp per pair (nw_batch_trace_kernel: move codes of every cell in shared memory,
    then the walk); empty sequences and pairs wider than the shared memory take the single-pair path.  Every pair vs the oracle."""
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(5000 + n_pairs)
    letters, offY, lenY, offX, lenX = _ragged(rng, n_pairs, max_y, max_x)
    # a few near-identical pairs (long diagonal runs) and one-sided pairs (long gap runs)
    for p in range(0, n_pairs, 7):
        k = int(min(lenY[p], lenX[p]))
        letters[int(offY[p]): int(offY[p]) + k] = letters[int(offX[p]): int(offX[p]) + k]
    l0 = engine.launches()
    scores, edits, hashes = engine.align_batch(letters, offY, lenY, offX, lenX, want_trace=True)
    for p in range(n_pairs):
        y = letters[int(offY[p]): int(offY[p]) + int(lenY[p])]
        x = letters[int(offX[p]): int(offX[p]) + int(lenX[p])]
        if y.size == 0 or x.size == 0:
            assert scores[p] == -11 * (y.size + x.size)
            continue
        exp = oracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
        assert (scores[p], edits[p], hashes[p]) == (exp.score, exp.edit, exp.trace_hash), p
    if max_x <= 512 and max_y <= 512:
        assert engine.launches() - l0 < 40, "transcripts must not go pair by pair"


@pytest.mark.parametrize("max_y,max_x", [(256, 256), (200, 400), (512, 300), (40, 40)])
def test_batch_affine_and_local_variants(engine, scoring, oracle, max_y, max_x):
    """The variants the reference lists as future work (README.md:6-29; --gapeCost): Gotoh affine gaps and Smith-Waterman over the
    resident batch against the CPU restatement (oracle/nw_oracle.c, parity unpinned), and the one case the reference DOES define:
    affine with gap_extend == gap_open equals its linear-gap recurrence, i.e. the default batch path and the golden-pinned oracle."""
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(7 * max_y + max_x)
    letters, offY, lenY, offX, lenX = _ragged(rng, 500, max_y, max_x)
    linear = oracle.score_batch(letters, offY, lenY, offX, lenX, subst, -11)
    assert np.array_equal(engine.align_batch_variant(letters, offY, lenY, offX, lenX, "nw_affine", -11, -11), linear)
    assert np.array_equal(engine.align_batch(letters, offY, lenY, offX, lenX), linear)
    for variant, local, go, ge in (("nw_affine", False, -11, -1), ("nw_affine", False, -5, -3), ("sw_affine", True, -11, -1), ("sw_linear", True, -11, -11),
                                   ("sw_affine", True, -4, 0)):
        got = engine.align_batch_variant(letters, offY, lenY, offX, lenX, variant, go, ge)
        exp = oracle.score_batch_gotoh(letters, offY, lenY, offX, lenX, subst, go, ge if variant != "sw_linear" else go, local)
        assert np.array_equal(got, exp), (variant, go, ge, int(np.count_nonzero(got != exp)))


def test_batch_variants_on_similar_sequences_and_errors(engine, scoring, oracle):
    from gpuseqalign_b200 import NwB200Error, synth
    subst = scoring["subst"]["blosum62"]
    # mutated copies: long diagonal runs with indels -- the regime affine gaps are for
    seqs = []
    for k in range(64):
        x = synth.letters(9000 + k, 256)
        seqs += [synth.mutated_copy(x, 9500 + k, 240 + k % 17), x]
    lens = np.array([s.size for s in seqs], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    pool = np.concatenate(seqs + [np.zeros(1, np.uint8)]).astype(np.uint8)
    offY, offX = offs[0:-1:2].copy(), offs[1::2].copy()
    lenY, lenX = lens[0::2].astype(np.uint32), lens[1::2].astype(np.uint32)
    for variant, local in (("nw_affine", False), ("sw_affine", True)):
        got = engine.align_batch_variant(pool, offY, lenY, offX, lenX, variant, -11, -1)
        assert np.array_equal(got, oracle.score_batch_gotoh(pool, offY, lenY, offX, lenX, subst, -11, -1, local)), variant
    with pytest.raises(NwB200Error):
        engine.batch_resident_variant("nw_affine", 3, -1)             # a positive gap cost
    bad = pool.copy(); bad[5] = 99
    with pytest.raises(NwB200Error):
        engine.align_batch_variant(bad, offY, lenY, offX, lenX, "sw_affine", -11, -1)


def _pack_batch(letters, offY, lenY, offX, lenX):
    """Packs the sequences of a batch (Y and X of every pair, in pool order) into one 5-bit pool; returns (packed, offY, offX)."""
    from gpuseqalign_b200 import synth
    n = len(lenY)
    offs = np.empty(2 * n, dtype=np.int64); lens = np.empty(2 * n, dtype=np.int64)
    offs[0::2] = offY; offs[1::2] = offX; lens[0::2] = lenY; lens[1::2] = lenX
    packed, noffs = synth.pack5(letters, offs, lens)
    return packed, noffs[0::2].copy(), noffs[1::2].copy()


@pytest.mark.parametrize("n_pairs,max_y,max_x", [(1, 256, 256), (601, 256, 300), (333, 100, 60), (64, 7, 9)])
def test_batch_packed5_letters(engine, scoring, oracle, n_pairs, max_y, max_x):
    """5-bit packed letters (8 letters in 5 bytes, every sequence on a byte boundary) through both batch forms: ragged lengths (row
    groups that straddle bytes, empty sequences, odd pair counts) and the aligned synthetic batch; equal to the byte-letter scores."""
    from gpuseqalign_b200 import NwB200Error, synth
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(5000 + n_pairs)
    letters, offY, lenY, offX, lenX = _ragged(rng, n_pairs, max_y, max_x)
    exp = oracle.score_batch(letters, offY, lenY, offX, lenX, subst, -11)
    packed, pY, pX = _pack_batch(letters, offY, lenY, offX, lenX)
    assert np.array_equal(engine.align_batch_packed5(packed, pY, lenY, pX, lenX), exp)
    engine.upload_batch_packed5(packed, pY, lenY, pX, lenX)
    engine.batch_resident()
    assert np.array_equal(engine.fetch_batch_scores(), exp)
    with pytest.raises(NwB200Error):
        engine.batch_resident_variant("nw_affine", -11, -1)            # the variants take byte letters
    pool, oY, lY, oX, lX = synth.batch_pairs(77, 70000, 256, 256)           # several H2D slices
    packed, pY, pX = _pack_batch(pool, oY, lY, oX, lX)
    assert packed.size < pool.size * 0.63
    got = engine.align_batch_packed5(packed, pY, lY, pX, lX)
    assert np.array_equal(got, engine.align_batch(pool, oY, lY, oX, lX))
    assert np.array_equal(got[:3000], oracle.score_batch(pool[: 3000 * 512], oY[:3000], lY[:3000], oX[:3000], lX[:3000], subst, -11))
    bad = letters.copy()
    if bad.size > 3 and lenY[0] > 0:
        bad[int(offY[0])] = 27                                           # a 5-bit value outside the 25-letter alphabet
        packed, pY, pX = _pack_batch(bad, offY, lenY, offX, lenX)
        with pytest.raises(NwB200Error):
            engine.align_batch_packed5(packed, pY, lenY, pX, lenX)
    tall_y = np.array([300], dtype=np.uint32)
    with pytest.raises(NwB200Error):
        engine.align_batch_packed5(np.zeros(1024, np.uint8), np.array([0], np.uint64), tall_y, np.array([400], np.uint64), np.array([10], np.uint32))
