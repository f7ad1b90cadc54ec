"""GPU parity at BASELINE.json's FULL sizes (configs 2, 4, 5): known answers computed once by the CPU oracle
(tests/golden/make_big_golden.py -> big_golden.json) plus size-independent properties of the domain:
the transcript must consume exactly (lenY, lenX) with '=' / 'X' agreeing with the letters; NW(y, x) == NW(x, y) for a
symmetric matrix.  (The reference's traceback follows the best-scoring NEIGHBOUR, nwtrace1_plain.cpp:29-100, not the
predecessor that produced the cell, so its transcript does not in general re-score to align_cost -- and neither does ours.)"""
import hashlib
import json
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def big():
    with open(os.path.join(ROOT, "tests", "golden", "big_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def engine(scoring):
    from gpuseqalign_b200 import Engine
    e = Engine(0)
    e.set_scoring(scoring["subst"]["blosum62"], -11)
    yield e
    e.close()


def rescore_transcript(edit, y, x, subst, gap):
    """Walks the run-length transcript over (y, x): returns (score, consumed_y, consumed_x, consistent)."""
    S = int(round(len(subst) ** 0.5))
    sub = np.asarray(subst).reshape(S, S)
    i = j = 0
    score = 0
    ok = True
    for cnt, op in re.findall(r"(\d+)([=XID])", edit):
        k = int(cnt)
        if op in "=X":
            ys, xs = y[i:i + k], x[j:j + k]
            if len(ys) != k or len(xs) != k:
                return score, i, j, False
            eq = ys == xs
            ok &= bool(eq.all()) if op == "=" else bool((~eq).all())
            score += int(sub[ys, xs].sum())
            i += k; j += k
        elif op == "I":          # up: consumes a row letter
            score += k * gap; i += k
        else:                    # 'D' left: consumes a column letter
            score += k * gap; j += k
    return score, i, j, ok


def _inputs(name):
    from gpuseqalign_b200 import synth
    if name == "cfg2_random":
        return synth.letters(2002, 16384), synth.letters(2001, 16384)
    if name == "cfg2_mutated":
        x = synth.letters(2001, 16384)
        return synth.mutated_copy(x, 2003, 16384), x
    if name == "cfg4":
        return synth.letters(4001, 2048), synth.letters(4002, 4194304)
    x = synth.letters(5001, 200000)
    if name == "cfg5_mutated":
        return synth.mutated_copy(x, 5002, 200000), x
    return synth.letters(5004, 200000), x


@pytest.mark.parametrize("name", ["cfg2_random", "cfg2_mutated", "cfg5_mutated", "cfg5_random"])
def test_full_size_score_and_traceback(engine, big, scoring, name):
    g = big[name]
    y, x = _inputs(name)
    assert (y.size, x.size) == (g["len_y"], g["len_x"])
    assert engine.align(y, x, keep_headers=True) == g["score"]
    edit, th = engine.trace()
    assert len(edit) == g["edit_len"] and edit[:64] == g["edit_head"]
    assert hashlib.sha256(edit.encode()).hexdigest() == g["edit_sha256"]
    assert f"{th:08x}" == g["trace_hash"]
    sc, ci, cj, ok = rescore_transcript(edit, y, x, scoring["subst"]["blosum62"], -11)
    assert ok and (ci, cj) == (y.size, x.size)


def test_cfg2_score_hash(engine, big):
    y, x = _inputs("cfg2_random")
    engine.align(y, x, keep_headers=False)
    # the same value the reference's cpu4 + NwHash1_Plain produce for this pair (profiles/r1_reference_gpu9_cpu4_on_b200.jsonl run)
    h = engine.score_hash()
    from oracle import pyoracle
    exp = pyoracle.fill_rolling(y, x, np.asarray(json.load(open(os.path.join(ROOT, "tests", "golden", "scoring.json")))["subst"]["blosum62"], dtype=np.int32),
                                -11, want_hash=True)[3]
    assert h == exp


def test_cfg4_rectangular_score(engine, big):
    g = big["cfg4"]
    y, x = _inputs("cfg4")
    assert engine.align(y, x, keep_headers=False) == g["score"]
    # NW is symmetric for a symmetric substitution matrix: 4M rows x 2k columns exercises the many-band regime
    assert engine.align(x, y, keep_headers=False) == g["score"]


def test_cfg4_cfg5_through_the_column_block_wavefront(engine, big):
    """Same kernel path as the cross-GPU run, on one GPU (blocks exchange borders through the GPU's own receive buffer)."""
    from gpuseqalign_b200.wavefront import wave_align
    y, x = _inputs("cfg4")
    assert wave_align(engine, y, x, block_cols=65536, epoch=7) == big["cfg4"]["score"]
    y, x = _inputs("cfg5_random")
    assert wave_align(engine, y, x, block_cols=4096, epoch=8) == big["cfg5_random"]["score"]


def test_batch_checksum_full_cfg3_slice(engine, scoring, oracle):
    """65 536 pairs of the cfg3 batch: every score vs the oracle, plus the transcript property on a sample."""
    from gpuseqalign_b200 import synth
    subst = scoring["subst"]["blosum62"]
    pool, offY, lenY, offX, lenX = synth.batch_pairs(500000, 65536, 256, 256)
    got = engine.align_batch(pool, offY, lenY, offX, lenX)
    exp = oracle.score_batch(pool, offY, lenY, offX, lenX, subst, -11)
    assert np.array_equal(got, exp)
    k = 64
    sc, edits, hashes = engine.align_batch(pool, offY[:k], lenY[:k], offX[:k], lenX[:k], want_trace=True)
    for p in range(k):
        y = pool[int(offY[p]): int(offY[p]) + 256]; x = pool[int(offX[p]): int(offX[p]) + 256]
        s2, ci, cj, ok = rescore_transcript(edits[p], y, x, subst, -11)
        assert ok and (ci, cj) == (256, 256) and sc[p] == exp[p]


def test_prefix_max_scorer_small_shapes(engine, scoring, oracle):
    """Row-parallel prefix-max formulation (nw_scan.cuh) vs the oracle: chunk boundaries, single rows, many chunks."""
    from gpuseqalign_b200 import synth
    from gpuseqalign_b200.wavefront import scan_align
    subst = scoring["subst"]["blosum62"]
    for n, m, seed in [(1, 1, 1), (3, 4095, 2), (5, 4096, 3), (7, 4097, 4), (64, 20000, 5), (300, 70000, 6), (2048, 9000, 7), (17, 300000, 8)]:
        y = synth.letters(100 + seed, n); x = synth.letters(200 + seed, m)
        exp, _, _, _ = oracle.fill_rolling(y, x, subst, -11)
        assert scan_align(engine, y, x, epoch=seed) == exp, (n, m)


def test_cfg4_prefix_max_scorer(engine, big):
    from gpuseqalign_b200.wavefront import scan_align
    y, x = _inputs("cfg4")
    assert scan_align(engine, y, x, epoch=44) == big["cfg4"]["score"]
