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
