"""CPU-only: pin the oracle (oracle/nw_oracle.c) against the reference's golden vectors.

Golden vectors = outputs of the reference's own cpu4 + NwHash1_Plain + NwTrace1_Plain
(tests/golden/make_golden.py) and the SURVEY.md Appendix B table.
"""
import numpy as np
import pytest

from conftest import case_letters

# SURVEY.md Appendix B (subset; the full generated set is in tests/golden/)
APPENDIX_B = {
    ("len1", "len1"): (5, "7c5d3ce0", "00596829", "1="),
    ("len1", "len2"): (-6, "08ab6870", "7c541c3c", "1=1D"),
    ("len1", "len728"): (-7992, "d25a6eb0", "e1cf965f", "1=727D"),
    ("len4", "len8"): (-35, "644a9b4d", "e1cedb92", "1=3X4D"),
    ("len8", "len16"): (-75, "a6e2c3c1", "3c1ab2b4", "1=2X1=3X1=8D"),
    ("len32", "len33"): (-33, "212cf669", "d7d9ce8d", "26X1=5X1D"),
    ("len256", "len256"): (1320, "8612d3f3", "7c548709", "256="),
    ("len728", "len728"): (3774, "6ddeeb24", "7c52dee5", "728="),
}


def test_golden_matches_appendix_b(golden):
    seen = 0
    for c in golden["cases"]:
        key = (c["y"], c["x"])
        if key in APPENDIX_B and c["y_range"] == [None, None] and c["x_range"] == [None, None]:
            exp = APPENDIX_B[key]
            assert (c["score"], c["score_hash"], c["trace_hash"], c["edit"]) == exp
            seen += 1
    assert seen >= len(APPENDIX_B)


def test_oracle_full_matrix_vs_golden(golden, scoring, oracle):
    subst = scoring["subst"]["blosum62"]
    for c in golden["cases"]:
        y, x = case_letters(golden, c)
        r = oracle.align_pair(y, x, subst, golden["gap"], want_hash=True, want_trace=True)
        assert r.score == c["score"], c
        assert f"{r.score_hash:08x}" == c["score_hash"], c
        assert f"{r.trace_hash:08x}" == c["trace_hash"], c
        assert r.edit == c["edit"], c


def test_oracle_mt_equals_st(golden, scoring, oracle):
    subst = scoring["subst"]["blosum62"]
    for c in golden["cases"][::9]:
        y, x = case_letters(golden, c)
        a = oracle.align_pair(y, x, subst, golden["gap"], want_hash=True, threads=1)
        b = oracle.align_pair(y, x, subst, golden["gap"], want_hash=True, threads=4, blocksz=64)
        assert (a.score, a.score_hash, a.trace_hash, a.edit) == (b.score, b.score_hash, b.trace_hash, b.edit)


@pytest.mark.parametrize("By,Bx", [(128, 209), (32, 32), (512, 256), (7, 5)])
def test_oracle_sparse_headers_and_trace_vs_golden(golden, scoring, oracle, By, Bx):
    """Rolling-row header producer + sparse traceback (gpu9 layout / NwTrace2_Sparse restatement)."""
    subst = scoring["subst"]["blosum62"]
    step = 1 if (By, Bx) == (128, 209) else 5
    for c in golden["cases"][::step]:
        y, x = case_letters(golden, c)
        score, hrow, hcol, sh = oracle.fill_rolling(y, x, subst, golden["gap"], By, Bx, want_hash=True)
        assert score == c["score"]
        assert f"{sh:08x}" == c["score_hash"]
        r = oracle.trace_sparse(hrow, hcol, By, Bx, y, x, subst, golden["gap"])
        assert r.score == c["score"]
        assert r.edit == c["edit"]
        assert f"{r.trace_hash:08x}" == c["trace_hash"]


def test_oracle_batch_scores(golden, scoring, oracle):
    subst = scoring["subst"]["blosum62"]
    cases = [c for c in golden["cases"] if c["len_y"] * c["len_x"] <= 800 * 800]
    letters, offY, lenY, offX, lenX = [], [], [], [], []
    pos = 0
    for c in cases:
        y, x = case_letters(golden, c)
        offY.append(pos); lenY.append(y.size); letters.append(y); pos += y.size
        offX.append(pos); lenX.append(x.size); letters.append(x); pos += x.size
    scores = oracle.score_batch(np.concatenate(letters), offY, lenY, offX, lenX, subst, golden["gap"], threads=2)
    assert scores.tolist() == [c["score"] for c in cases]


def test_synth_generator_is_splitmix64(oracle):
    # independent restatement of SURVEY.md 8(d) in python ints
    def ref(seed, n):
        s, out = seed, []
        M = (1 << 64) - 1
        for _ in range(n):
            s = (s + 0x9E3779B97F4A7C15) & M
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
            z = z ^ (z >> 31)
            out.append((z >> 33) % 20)
        return out
    assert oracle.synth_letters(2001, 64).tolist() == ref(2001, 64)


def test_oracle_vs_live_reference_when_present(golden, scoring, oracle):
    """When oracle/_ref/libnwref.so was prebuilt, the restatement must equal the live reference
    (cpu4 and NwTrace2_Sparse fed with the restated headers) on fresh random inputs."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    subst = scoring["subst"]["blosum62"]
    rng = np.random.default_rng(7)
    for n, m in [(1, 1), (3, 700), (130, 131), (257, 640), (1000, 1300)]:
        y = rng.integers(0, 24, n).astype(np.uint8)
        x = rng.integers(0, 24, m).astype(np.uint8)
        ref = oracle.ref_run("cpu4", y, x, subst, -11)
        mine = oracle.align_pair(y, x, subst, -11, want_hash=True)
        assert (ref.score, ref.score_hash, ref.trace_hash, ref.edit) == (mine.score, mine.score_hash, mine.trace_hash, mine.edit)
        score, hrow, hcol, _ = oracle.fill_rolling(y, x, subst, -11, 128, 209)
        r2 = oracle.ref_trace_from_headers(y, x, subst, -11, hrow, hcol, 128, 209, want_hash=True)
        assert (r2.score, r2.score_hash, r2.trace_hash, r2.edit) == (ref.score, ref.score_hash, ref.trace_hash, ref.edit)


def _gotoh_py(y, x, subst, S, go, ge, local):
    """Independent pure-Python Gotoh (three explicit matrices, no rolling rows): the textbook recurrence the C restatement
    (nwo_score_batch_gotoh) and the GPU kernel implement; a gap of L residues costs go + (L - 1) * ge."""
    NEG = -10 ** 9
    n, m = len(y), len(x)
    H = [[0] * (m + 1) for _ in range(n + 1)]
    E = [[NEG] * (m + 1) for _ in range(n + 1)]
    F = [[NEG] * (m + 1) for _ in range(n + 1)]
    if not local:
        for i in range(1, n + 1): H[i][0] = go + (i - 1) * ge
        for j in range(1, m + 1): H[0][j] = go + (j - 1) * ge
    best = 0
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            E[i][j] = max(E[i][j - 1] + ge, H[i][j - 1] + go)
            F[i][j] = max(F[i - 1][j] + ge, H[i - 1][j] + go)
            h = max(H[i - 1][j - 1] + int(subst[int(y[i - 1]) * S + int(x[j - 1])]), E[i][j], F[i][j])
            if local: h = max(h, 0)
            H[i][j] = h
            best = max(best, h)
    return best if local else H[n][m]


def test_gotoh_restatement_against_an_independent_implementation(oracle, scoring):
    """The affine-gap / Smith-Waterman restatement (parity unpinned: the reference implements neither) against a pure-Python textbook
    Gotoh on small ragged pairs, and its one pinned case: gap_extend == gap_open is the reference's linear-gap recurrence."""
    subst = scoring["subst"]["blosum62"]
    S = int(round(len(subst) ** 0.5))
    rng = np.random.default_rng(11)
    n = 40
    lenY = rng.integers(0, 30, n).astype(np.uint32); lenX = rng.integers(0, 30, n).astype(np.uint32)
    lens = np.empty(2 * n, dtype=np.uint64); lens[0::2] = lenY; lens[1::2] = lenX
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    letters = rng.integers(0, 24, int(offs[-1]) + 1).astype(np.uint8)
    offY, offX = offs[0:-1:2].copy(), offs[1::2].copy()
    for go, ge, local in ((-11, -1, False), (-5, -3, False), (-11, -1, True), (-11, -11, True), (-4, 0, True), (-7, -7, False)):
        got = oracle.score_batch_gotoh(letters, offY, lenY, offX, lenX, subst, go, ge, local)
        for p in range(n):
            y = letters[int(offY[p]): int(offY[p]) + int(lenY[p])]; x = letters[int(offX[p]): int(offX[p]) + int(lenX[p])]
            if len(y) == 0 or len(x) == 0:
                k = len(y) + len(x)
                exp = 0 if (local or k == 0) else go + (k - 1) * ge
            else:
                exp = _gotoh_py(y, x, subst, S, go, ge, local)
            assert got[p] == exp, (go, ge, local, p, len(y), len(x))
    assert np.array_equal(oracle.score_batch_gotoh(letters, offY, lenY, offX, lenX, subst, -11, -11, False),
                          oracle.score_batch(letters, offY, lenY, offX, lenX, subst, -11))
