#!/usr/bin/env python
"""SASS opcode summary of the kernels in a cubin / shared library (cuobjdump -sass): per kernel the total opcode histogram and the
instruction mix of its longest loop (the unrolled 32-step chunk of the sweep kernels), per step.

    tools/sass_mix.py gpuseqalign_b200/libnwb200.so [kernel-name-substring] [--steps 32] > profiles/rN_sass_opcodes.txt
"""
import collections, re, subprocess, sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, name = None, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur is not None:
                yield name, cur
            name, cur = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    if cur is not None:
        yield name, cur


def opcode(ins):
    t = ins.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


def main():
    path = sys.argv[1]
    want = [a for a in sys.argv[2:] if not a.startswith("--")]
    steps = 32
    if "--steps" in sys.argv:
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
    for name, ins in kernels(path):
        if want and not any(w in name for w in want):
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        tot = collections.Counter(opcode(i).split(".")[0] for _, i in ins)
        print(f"== {dem}\n   {len(ins)} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in tot.most_common(14)))
        full = collections.Counter(opcode(i) for _, i in ins)
        marks = [k for k in full if any(s in k for s in ("VIMNMX3", "VIADDMNMX", "IDP", "UBLKCP", "UTMA", "MAPA", "UCGABAR", "SYNCS", "REDUX", "ATOM", "LDGSTS"))]
        if marks:
            print("   blackwell / DPX / TMA opcodes: " + ", ".join(f"{k} {full[k]}" for k in sorted(marks)))
        # every backward-branch loop of at least 100 instructions (the unrolled 32-step chunk of the sweep kernels is one of them)
        addr_ix = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and tgt in addr_ix and i - addr_ix[tgt] + 1 >= 100:
                    lo, hi = addr_ix[tgt], i
                    n = hi - lo + 1
                    mix = collections.Counter(opcode(t2).split(".")[0] for _, t2 in ins[lo:hi + 1])
                    print(f"   loop of {n} instructions = {n / steps:.2f} per step of {steps}: " + ", ".join(f"{k} {v / steps:.2f}" for k, v in mix.most_common(14)))


if __name__ == "__main__":
    main()
