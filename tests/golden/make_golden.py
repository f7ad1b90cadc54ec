#!/usr/bin/env python
"""Generate tests/golden/*.json by RUNNING THE REFERENCE ITSELF (in the build container).

Inputs : /root/reference/resrc/{subst.json, seq_generated.fa, pair_debug.txt, pair_generated_1.txt}
Engine : oracle/_ref/libnwref.so = the unmodified reference sources (NwAlign_Cpu4_Mt_DiagRow,
         blocksz 256 + NwHash1_Plain + NwTrace1_Plain) behind oracle/ref_shim.cpp.
Output : tests/golden/nw_golden_blosum62.json   -- all 173 pair_debug pairs + the pair_generated_1
         pairs up to 5000x5000 (inputs as letter strings, outputs score / hashes / transcript)
         tests/golden/scoring.json              -- letter map + the five substitution matrices

The reference cannot travel to the GPU box, so these vectors are committed; re-run this
script (python tests/golden/make_golden.py) whenever the reference changes.  The md5 digests
printed at the end must equal SURVEY.md section 4 ([probe] values).
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gpuseqalign_b200 import formats  # noqa: E402
from oracle import pyoracle  # noqa: E402

REF = "/root/reference/resrc"
GAP = -11
MAX_CELLS = 5000 * 5000


def main():
    subst = formats.read_subst(os.path.join(REF, "subst.json"))
    seqs = formats.read_fasta(os.path.join(REF, "seq_generated.fa"), subst)
    mat = subst.matrix("blosum62")

    cases = []
    used = set()
    digest_lines = []
    digest_lines_edit = []
    for fname in ("pair_debug.txt", "pair_generated_1.txt"):
        for p in formats.read_pairs(os.path.join(REF, fname), seqs):
            y, x = formats.pair_letters(p, seqs)
            if y.size * x.size > MAX_CELLS:
                continue
            r = pyoracle.ref_run("cpu4", y, x, mat, GAP, want_hash=True, want_trace=True)
            used.update([p.y_id, p.x_id])
            cases.append({
                "src": fname, "y": p.y_id, "x": p.x_id,
                "y_range": [p.y_range.l, p.y_range.r], "x_range": [p.x_range.l, p.x_range.r],
                "len_y": int(y.size), "len_x": int(x.size),
                "score": r.score, "score_hash": f"{r.score_hash:08x}", "trace_hash": f"{r.trace_hash:08x}",
                "edit": r.edit,
            })
            if fname == "pair_debug.txt":
                yid = p.y_id + p.y_range.suffix()
                xid = p.x_id + p.x_range.suffix()
                line = f"{yid} {xid} {r.score} {r.score_hash:08x} {r.trace_hash:08x}"
                digest_lines.append(line)
                digest_lines_edit.append(line + " " + r.edit)

    out = {
        "generator": "tests/golden/make_golden.py (reference NwAlign_Cpu4_Mt_DiagRow blocksz=256 + NwHash1_Plain + NwTrace1_Plain)",
        "subst_name": "blosum62", "gap": GAP, "letters": subst.letters,
        "seqs": {sid: subst.decode(seqs.seqs[sid]) for sid in seqs.ids if sid in used},
        "cases": cases,
    }
    with open(os.path.join(ROOT, "tests", "golden", "nw_golden_blosum62.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    scoring = {"letters": subst.letters, "subst": {k: [int(v) for v in subst.subst_map[k]] for k in subst.subst_map}}
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json"), "w") as f:
        json.dump(scoring, f, separators=(",", ":"))

    md5 = hashlib.md5(("\n".join(digest_lines) + "\n").encode()).hexdigest()
    md5e = hashlib.md5(("\n".join(digest_lines_edit) + "\n").encode()).hexdigest()
    print(f"cases: {len(cases)}  pair_debug md5: {md5}  with edit: {md5e}")


if __name__ == "__main__":
    main()
