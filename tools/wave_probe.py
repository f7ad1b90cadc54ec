#!/usr/bin/env python
"""Timing sweep of the cross-GPU column-block wavefront on cfg5 (200 000^2): rows per lane R, lane skew K, block width.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/wave_probe.py [len] [R:K:block ...]
Rank 0 prints one JSON line per setting (device ms = max over ranks of the fill launch, score check against the golden)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from gpuseqalign_b200 import Engine, Params, synth
from gpuseqalign_b200.wavefront import wave_setup


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args = sys.argv[1:]
    n = int(args[0]) if args and args[0].isdigit() else 200000
    settings = [a for a in args if ":" in a] or ["8:2:2048", "4:2:2048", "4:1:2048", "4:2:8192", "4:2:32768", "8:2:8192"]
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
        subst = np.array(json.load(f)["subst"]["blosum62"], dtype=np.int32)
    big = json.load(open(os.path.join(ROOT, "tests", "golden", "big_golden.json")))
    exp = big["cfg5_random"]["score"] if n == 200000 else None
    x = synth.letters(5001, n); y = synth.letters(5004, n)
    eng = Engine(local); eng.set_scoring(subst, -11)
    epoch = 500
    for st in settings:
        R, K, block = (int(v) for v in st.split(":"))
        wave_setup(eng, y, x, rank=rank, world=world, block_cols=block, params=Params(rows_per_lane=R, skew=K))
        times = []
        score = None
        for it in range(4):
            epoch += 1
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            eng.wave_fill(epoch)
            s = eng.wave_fetch()
            if s is not None:
                score = s
            t = torch.tensor([eng.timing()["align_calc"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        sc = torch.tensor([score if score is not None else -(2 ** 62)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(sc, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"n_gpus": world, "len": n, "R": R, "K": K, "block_cols": block, "fill_ms": min(times[1:]), "all_ms": [round(v, 3) for v in times],
                              "gcups": float(n) * n / min(times[1:]) / 1e6, "score_ok": (int(sc.item()) == exp) if exp is not None else None}), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
