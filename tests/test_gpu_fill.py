"""GPU parity tests for the score-matrix fill: engine (through the C ABI) vs the CPU oracle.

Bit-exact: integer scores and every tile-header value that the traceback consumes.
"""
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


def test_golden_headers_feed_reference_style_trace(engine, golden, scoring, oracle):
    """Headers exported in the reference layout must let the (restated) NwTrace2_Sparse
    reproduce the golden transcript: validates every consumed header value."""
    from gpuseqalign_b200 import Params
    subst = scoring["subst"]["blosum62"]
    for c in golden["cases"][::3]:
        y, x = case_letters(golden, c)
        score = engine.align(y, x, keep_headers=True, params=Params(tile_cols=64))
        assert score == c["score"]
        hrow, hcol = engine.headers()
        info = engine.info
        r = oracle.trace_sparse(hrow, hcol, info.tile_rows, info.tile_cols, y, x, subst, -11)
        assert r.score == c["score"]
        assert r.edit == c["edit"], (c["y"], c["x"])


@pytest.mark.parametrize("R,W,K,Bx", [(4, 4, 2, 512), (4, 4, 1, 512), (8, 4, 1, 256), (8, 4, 2, 96), (4, 8, 2, 32), (8, 8, 1, 1024)])
def test_random_shapes_all_kernel_variants(engine, scoring, oracle, R, W, K, Bx):
    from gpuseqalign_b200 import Params
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(R * 100 + W * 10 + K)
    By = R * 32 * W
    shapes = [(1, 1), (1, 40), (5, 3), (31, 33), (By - 1, By + 1), (By, 2 * By), (By + 1, 777), (2 * By + 17, 3 * By + 5), (1500, 4100)]
    for n, m in shapes:
        y = rng.integers(0, 20, n).astype(np.uint8)
        x = rng.integers(0, 20, m).astype(np.uint8)
        exp, hrow_o, hcol_o, _ = oracle.fill_rolling(y, x, subst, -11, By, Bx)
        got = engine.align(y, x, keep_headers=True, params=Params(R, W, Bx, K))
        assert got == exp, (n, m)
        hrow, hcol = engine.headers()
        info = engine.info
        assert (info.tile_rows, info.tile_cols) == (By, Bx)
        # compare every header entry that lies inside the real matrix
        trows, tcols = info.trows, info.tcols
        hr = hrow.reshape(trows, tcols, 1 + Bx); hro = hrow_o.reshape(trows, tcols, 1 + Bx)
        hc = hcol.reshape(trows, tcols, 1 + By); hco = hcol_o.reshape(trows, tcols, 1 + By)
        for jT in range(tcols):
            kmax = min(Bx, m - jT * Bx)
            assert np.array_equal(hr[:, jT, : kmax + 1], hro[:, jT, : kmax + 1]), (n, m, "hrow", jT)
        for iT in range(trows):
            kmax = min(By, n - iT * By)
            assert np.array_equal(hc[iT, :, : kmax + 1], hco[iT, :, : kmax + 1]), (n, m, "hcol", iT)


def test_similar_sequences_long(engine, scoring, oracle):
    """A mutated copy (mostly diagonal path) and an unrelated pair at 5000 x 6000."""
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(11)
    x = rng.integers(0, 20, 6000).astype(np.uint8)
    y = x[:5000].copy()
    idx = rng.integers(0, 5000, 500)
    y[idx] = rng.integers(0, 20, 500)
    exp, _, _, _ = oracle.fill_rolling(y, x, subst, -11)
    assert engine.align(y, x, keep_headers=False) == exp
    y2 = rng.integers(0, 20, 5000).astype(np.uint8)
    exp2, _, _, _ = oracle.fill_rolling(y2, x, subst, -11)
    assert engine.align(y2, x, keep_headers=False) == exp2


def test_other_scoring(engine, scoring, oracle):
    from gpuseqalign_b200 import Engine
    e = Engine(0)
    try:
        rng = np.random.default_rng(5)
        y = rng.integers(0, 25, 700).astype(np.uint8)
        x = rng.integers(0, 25, 900).astype(np.uint8)
        for name, gap in [("blosum45", -5), ("blosum90", -20), ("blosum50", -1), ("blosum80", 0)]:
            subst = scoring["subst"][name]
            e.set_scoring(subst, gap)
            exp, _, _, _ = oracle.fill_rolling(y, x, subst, gap)
            assert e.align(y, x, keep_headers=False) == exp, (name, gap)
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
