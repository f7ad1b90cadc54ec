#!/usr/bin/env python
"""Cross-GPU column-block wavefront check + timing (one process per GPU, launch with torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/wave_check.py [cfg ...]

Every rank aligns the SAME pair cooperatively; rank 0 prints one JSON line per config with the score check against
tests/golden/big_golden.json and the device time (max over ranks).  cfgs: small, cfg2, cfg4, cfg5 (default: small cfg2 cfg5 cfg4)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from gpuseqalign_b200 import Engine, Params, synth
from gpuseqalign_b200.wavefront import wave_align, scan_align


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfgs = sys.argv[1:] or ["small", "cfg2", "cfg5", "cfg4"]
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
        subst = np.array(json.load(f)["subst"]["blosum62"], dtype=np.int32)
    big = json.load(open(os.path.join(ROOT, "tests", "golden", "big_golden.json")))
    eng = Engine(local); eng.set_scoring(subst, -11)
    epoch = 100
    for cfg in cfgs:
        if cfg == "small":
            y = synth.letters(61, 3000); x = synth.letters(62, 20000); block = 512; exp = None
        elif cfg == "cfg2":
            y = synth.letters(2002, 16384); x = synth.letters(2001, 16384); block = 1024; exp = big["cfg2_random"]["score"]
        elif cfg == "scan4":        # cfg4 through the row-parallel prefix-max scorer
            y = synth.letters(4001, 2048); x = synth.letters(4002, 4194304); block = 4096; exp = big["cfg4"]["score"]
        elif cfg == "cfg4":
            y = synth.letters(4001, 2048); x = synth.letters(4002, 4194304); block = 16384; exp = big["cfg4"]["score"]
        else:
            x = synth.letters(5001, 200000); y = synth.letters(5004, 200000); block = int(os.environ.get("BLOCK", "2048")); exp = big["cfg5_random"]["score"]
        if exp is None:
            # single-GPU engine result as the expectation (rank 0 computes it, everyone gets it)
            t = torch.zeros(1, dtype=torch.int64, device="cuda")
            if rank == 0:
                t[0] = eng.align(y, x, keep_headers=False)
            if world > 1:
                dist.broadcast(t, 0)
            exp = int(t.item())
        times = []
        score = None
        for it in range(4):
            epoch += 1
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if cfg == "scan4":
                score = scan_align(eng, y, x, rank=rank, world=world, epoch=epoch)
            else:
                score = wave_align(eng, y, x, rank=rank, world=world, block_cols=block, epoch=epoch)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt, eng.timing()["align_calc"] / 1e3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            times.append(tt.tolist())
        if rank == 0:
            best_wall = min(t[0] for t in times[1:]); best_dev = min(t[1] for t in times[1:])
            cells = float(y.size) * float(x.size)
            print(json.dumps({"cfg": cfg, "n_gpus": world, "len_y": int(y.size), "len_x": int(x.size), "block_cols": block,
                              "score": score, "expected": exp, "ok": score == exp,
                              "fill_ms_device_max": best_dev * 1e3, "gcups_device": cells / best_dev / 1e9,
                              "wall_ms_incl_setup": best_wall * 1e3}), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
