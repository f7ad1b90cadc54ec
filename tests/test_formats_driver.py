"""The reference's on-disk formats and the TSV driver.  CPU part: readers on small files written here
and (when oracle/_ref/resrc travels along) on the reference's own fixtures; GPU part: the driver over pair_debug.txt must
reproduce the reference's own TSV values (profiles/r1_reference_benchmark_pair_debug_with_b200_plugin.tsv, cpu4 rows)."""
import csv
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RESRC = os.path.join(ROOT, "oracle", "_ref", "resrc")
REF_TSV = os.path.join(ROOT, "profiles", "r1_reference_benchmark_pair_debug_with_b200_plugin.tsv")


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_readers_on_handwritten_files(tmp_path):
    from gpuseqalign_b200 import formats
    subst = _write(tmp_path, "s.json", """// comment
{ "letterMap": { "A": 0, "C": 1, /* inline */ "G": 2 },
  "substMap": { "m": [ 1, -1, -2,  -1, 1, -3,  -2, -3, 1 ] } }""")
    sd = formats.read_subst(subst)
    assert sd.letter_map == {"A": 0, "C": 1, "G": 2} and list(sd.subst_map["m"]) == [1, -1, -2, -1, 1, -3, -2, -3, 1]
    fa = _write(tmp_path, "x.fa", ">s1 some info\nAC GA\nCC\n\n>s2\nGGG\n")
    seqs = formats.read_fasta(fa, sd)
    assert seqs.ids == ["s1", "s2"] and seqs.seqs["s1"].tolist() == [0, 1, 2, 0, 1, 1] and seqs.seqs["s2"].tolist() == [2, 2, 2]
    pairs = formats.read_pairs(_write(tmp_path, "p.txt", "s2 s1\n\ns2[1:] s1[ :4 ]\ns2[:] s1[2:5]\n"), seqs)
    assert [(p.y_id + p.y_range.suffix(), p.x_id + p.x_range.suffix()) for p in pairs] == [("s2", "s1"), ("s2[1:]", "s1[:4]"), ("s2", "s1[2:5]")]
    y, x = formats.pair_letters(pairs[1], seqs)
    assert y.tolist() == [2, 2] and x.tolist() == [0, 1, 2, 0]
    assert formats.with_header(y).tolist() == [0, 2, 2]
    with pytest.raises(formats.FormatError):
        formats.read_fasta(_write(tmp_path, "bad.fa", ">s1\nAXC\n"), sd)            # letter outside the map
    with pytest.raises(formats.FormatError):
        formats.read_fasta(_write(tmp_path, "dup.fa", ">s1\nA\n>s1\nC\n"), sd)      # duplicate id
    with pytest.raises(formats.FormatError):
        formats.read_pairs(_write(tmp_path, "p2.txt", "s2 s9\n"), seqs)             # unknown id
    with pytest.raises(formats.FormatError):
        formats.read_pairs(_write(tmp_path, "p3.txt", "s2[2:9] s1\n"), seqs)        # bad bounds
    params = formats.read_params(_write(tmp_path, "r.json", '{ "A": { "x": [1, 2], "y": [7] }, "B": {} }'))
    assert list(formats.param_combinations(params["A"])) == [{"x": 1, "y": 7}, {"x": 2, "y": 7}]


@pytest.mark.skipif(not os.path.isdir(RESRC), reason="oracle/_ref/resrc not present (built where /root/reference exists)")
def test_readers_on_reference_fixtures(golden):
    from gpuseqalign_b200 import formats
    sd = formats.read_subst(os.path.join(RESRC, "subst.json"))
    assert "".join(sorted(sd.letter_map, key=sd.letter_map.get)) == golden["letters"]
    seqs = formats.read_fasta(os.path.join(RESRC, "seq_generated.fa"), sd)
    for sid, s in golden["seqs"].items():
        assert np.array_equal(seqs.seqs[sid], golden["enc"][sid])
    pairs = formats.read_pairs(os.path.join(RESRC, "pair_debug.txt"), seqs)
    assert len(pairs) == 173
    assert (pairs[-1].y_id + pairs[-1].y_range.suffix(), pairs[-1].x_id + pairs[-1].x_range.suffix()) == ("len512[2:]", "len728[:726]")
    params = formats.read_params(os.path.join(RESRC, "param_best.json"))
    assert params["NwAlign_Gpu9_Mlsp_DiagDiagDiag"]["subtileBx"] == [48]


def test_tsv_columns_match_reference_header():
    from gpuseqalign_b200 import driver
    with open(REF_TSV) as f:
        header = f.readline().rstrip("\n").split("\t")
    assert driver.tsv_columns(True, True) == header


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(RESRC), reason="oracle/_ref/resrc not present")
def test_driver_reproduces_reference_tsv(tmp_path):
    from gpuseqalign_b200 import driver
    out = str(tmp_path / "b200.tsv")
    rows = driver.run(os.path.join(RESRC, "subst.json"), os.path.join(RESRC, "seq_generated.fa"), os.path.join(RESRC, "pair_debug.txt"), out,
                      calc_trace=True, calc_hash=True)
    assert len(rows) == 173
    with open(REF_TSV) as f:
        ref = [r for r in csv.DictReader(f, delimiter="\t") if r["alg_name"] == "NwAlign_Cpu4_Mt_DiagRow"]
    with open(out) as f:
        mine = list(csv.DictReader(f, delimiter="\t"))
    assert len(mine) == len(ref) == 173
    for a, b in zip(mine, ref):
        for col in ("seqY_idx", "seqX_idx", "seqY_id", "seqX_id", "seqY_len", "seqX_len", "subst_name", "gapo_cost", "align_cost",
                    "score_hash", "trace_hash", "edit_trace"):
            assert a[col] == b[col], (col, a["seqY_id"], a["seqX_id"])
    # scores only: the short pairs go through the batch kernel in one call
    rows2 = driver.run(os.path.join(RESRC, "subst.json"), os.path.join(RESRC, "seq_generated.fa"), os.path.join(RESRC, "pair_debug.txt"), None)
    assert [r["align_cost"] for r in rows2] == [int(b["align_cost"]) for b in ref]


class _OracleEngine:
    """Stand-in for gpuseqalign_b200.Engine in the CPU tests of the driver LOOP (parameter sweep, repeats, verification, error steps):
    results come from the CPU oracle (test infrastructure -- the product driver never does this); `wrong_for` makes one parameter
    combination return a wrong score, rowsPerLane = 3 is rejected like the engine rejects it."""

    def __init__(self, oracle, wrong_for=None):
        self.o, self.wrong_for, self.calls, self._lap = oracle, wrong_for, 0, 0.0

    def set_scoring(self, subst, gap):
        self.subst, self.gap = np.asarray(subst, dtype=np.int32).ravel(), gap

    def align(self, y, x, keep_headers=False, params=None, with_trace=False):
        from gpuseqalign_b200 import NwB200Error, NwStat
        self.calls += 1
        if params is not None and params.rows_per_lane == 3:
            raise NwB200Error(int(NwStat.errorInvalidValue), "unsupported tile parameters")
        self._res = self.o.align_pair(y, x, self.subst, self.gap, want_hash=True, want_trace=True)
        self._lap = float(self.calls)                       # a different lap every call: the mean over the sample runs is checked
        bump = 1 if (params is not None and self.wrong_for is not None and params.tile_cols == self.wrong_for) else 0
        return self._res.score + bump

    def timing(self):
        return {"align_cpy_dev": 0.5, "align_calc": self._lap, "align_cpy_host": 0.25, "trace_calc": 2.0, "trace_cpy_host": 1.0}

    def memory_usage(self):
        return {"pinned_host_bytes": 1, "device_bytes": 2, "shared_bytes": 3, "local_bytes": 4, "register_bytes": 5}

    def score_hash(self):
        return self._res.score_hash

    def trace(self):
        return self._res.edit, self._res.trace_hash

    def align_batch(self, pool, offY, lenY, offX, lenX):
        return self.o.score_batch(pool, offY, lenY, offX, lenX, self.subst, self.gap)

    def close(self):
        pass


def _small_inputs(tmp_path):
    subst = _write(tmp_path, "subst.json", '{"letterMap": {"A": 0, "C": 1, "G": 2, "T": 3}, "substMap": {"m": [2,-1,-1,-1, -1,2,-1,-1, -1,-1,2,-1, -1,-1,-1,2]}}')
    fa = _write(tmp_path, "s.fa", ">a\nACGTACGTAC\n>b\nACGTTCGTAC\n>c\nGGGTACCA\n")
    pairs = _write(tmp_path, "p.txt", "a b\nb c\na[2:8] c\n")
    return subst, fa, pairs


def test_driver_loop_param_sweep_repeats_and_verification(tmp_path, oracle):
    """benchmarkAlgs semantics without a GPU: cartesian parameter sweep (last key fastest), warm-up runs discarded and laps averaged over
    the sample runs, first result of a pair is the truth (a combination that disagrees gets err_step 5 and fails the run), a rejected
    combination gets err_step 1 and the sweep goes on."""
    from gpuseqalign_b200 import driver
    subst, fa, pairs = _small_inputs(tmp_path)
    params = _write(tmp_path, "par.json", '// sweep\n{"NwAlign_B200": {"rowsPerLane": [4, 3], "tileCols": [64, 128]}, "Other": {"x": [1]}}')
    eng = _OracleEngine(oracle, wrong_for=128)
    out = str(tmp_path / "o.tsv")
    rep = driver.run(subst, fa, pairs, out, subst_name="m", gapo_cost=-2, calc_trace=True, calc_hash=True, engine=eng, param_path=params,
                     warmup=1, samples=3)
    assert len(rep) == 3 * 4
    combos = [r["alg_params"] for r in rep[:4]]
    assert combos == ['{"rowsPerLane":4,"tileCols":64}', '{"rowsPerLane":4,"tileCols":128}', '{"rowsPerLane":3,"tileCols":64}', '{"rowsPerLane":3,"tileCols":128}']
    first = rep[0]
    assert first["err_step"] == 0 and first["warmup_runs"] == 1 and first["sample_runs"] == 3 and first["last_run_idx"] == 2
    assert first["align.calc"] == pytest.approx((2 + 3 + 4) / 3)            # calls 2..4 are the sample runs, call 1 the discarded warm-up
    assert first["trace.calc"] == pytest.approx(3.0) and first["glmem_peak_allocs"] == 2
    assert rep[1]["err_step"] == 5 and rep[1]["nw_stat"] == 9                # tileCols 128 returns a wrong score: caught by the first result
    assert rep[2]["err_step"] == 1 and rep[2]["nw_stat"] == 8                # rowsPerLane 3 is rejected: bad parameters, skipped
    assert rep.calc_errors == 3
    with open(out) as f:
        rows = list(csv.DictReader(f, delimiter="\t"))
    assert len(rows) == 12 and rows[0]["align.calc"] == "3.0000" and rows[0]["seqY_id"] == "a" and rows[8]["seqY_id"] == "a[2:8]"


def test_driver_verifies_against_a_reference_tsv(tmp_path, oracle):
    from gpuseqalign_b200 import driver
    subst, fa, pairs = _small_inputs(tmp_path)
    good = driver.run(subst, fa, pairs, str(tmp_path / "ref.tsv"), subst_name="m", gapo_cost=-2, calc_trace=True, calc_hash=True, engine=_OracleEngine(oracle))
    assert good.calc_errors == 0 and len(good) == 3
    # the same run verified against that TSV passes; against a doctored TSV it fails on the doctored pair only
    rep = driver.run(subst, fa, pairs, None, subst_name="m", gapo_cost=-2, calc_trace=True, calc_hash=True, engine=_OracleEngine(oracle), verify_tsv=str(tmp_path / "ref.tsv"))
    assert rep.calc_errors == 0
    lines = open(tmp_path / "ref.tsv").read().splitlines()
    cols = lines[0].split("\t"); k = cols.index("align_cost")
    bad = lines[2].split("\t"); bad[k] = str(int(bad[k]) + 1); lines[2] = "\t".join(bad)
    (tmp_path / "bad.tsv").write_text("\n".join(lines) + "\n")
    rep = driver.run(subst, fa, pairs, None, subst_name="m", gapo_cost=-2, engine=_OracleEngine(oracle), verify_tsv=str(tmp_path / "bad.tsv"))
    assert rep.calc_errors == 1 and [r["err_step"] for r in rep] == [0, 5, 0]      # (score-only: the batch path is verified as well)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(RESRC), reason="oracle/_ref/resrc not present")
def test_driver_sweep_and_repeats_on_the_gpu(tmp_path):
    """The engine under the driver's sweep: every kernel shape of the parameter file reproduces the reference's TSV (verification against
    the cpu4 rows), with warm-up and sample repeats; the score-only batch sharded over every GPU of the box."""
    import torch
    from gpuseqalign_b200 import driver
    params = _write(tmp_path, "par.json", '{"NwAlign_B200": {"rowsPerLane": [4, 8], "warpsPerBlock": [0], "tileCols": [64, 256], "skew": [1, 2]}}')
    pairs = _write(tmp_path, "p.txt", "\n".join(open(os.path.join(RESRC, "pair_debug.txt")).read().splitlines()[100:140]) + "\n")
    rep = driver.run(os.path.join(RESRC, "subst.json"), os.path.join(RESRC, "seq_generated.fa"), pairs, str(tmp_path / "o.tsv"),
                     calc_trace=True, calc_hash=True, param_path=params, warmup=1, samples=2, verify_tsv=REF_TSV)
    assert len(rep) % 8 == 0 and len(rep) >= 30 * 8 and rep.calc_errors == 0 and all(r["err_step"] == 0 for r in rep)      # (8 combinations per pair)
    assert all(r["glmem_peak_allocs"] > 0 and r["regmem_peak_allocs"] > 0 and r["align.calc"] > 0 for r in rep)
    devs = list(range(torch.cuda.device_count()))
    rep = driver.run(os.path.join(RESRC, "subst.json"), os.path.join(RESRC, "seq_generated.fa"), os.path.join(RESRC, "pair_debug.txt"), None,
                     warmup=1, samples=2, verify_tsv=REF_TSV, devices=devs)
    assert len(rep) == 173 and rep.calc_errors == 0
