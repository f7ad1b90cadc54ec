#!/usr/bin/env python
"""Developer check of the cross-GPU fill + traceback (wave_trace) on the GPUs of one box, one process per GPU:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/wave_trace_check.py [n] [kinds] [blocks]
Prints one line per step to stderr (progress survives a hang) and one JSON line per case from rank 0."""
import datetime, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def log(*a):
    print(f"[{time.strftime('%H:%M:%S')}] rank {os.environ.get('RANK')}:", *a, file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=60))
    from gpuseqalign_b200 import Engine, synth
    from gpuseqalign_b200.wavefront import wave_trace_setup, wave_trace
    from oracle import pyoracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subst = np.array(json.load(open(os.path.join(root, "tests", "golden", "scoring.json")))["subst"]["blosum62"], dtype=np.int32)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mutated", "random", "long_indel"]
    blocks = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [5120]
    eng = Engine(local)
    eng.set_scoring(subst, -11)
    epoch = 10
    x = synth.letters(611 if n != 200000 else 5001, n)
    for kind in kinds:
        if kind == "mutated": y = synth.mutated_copy(x, 612, n)
        elif kind == "random": y = synth.letters(614 if n != 200000 else 5004, n)
        else: y = np.concatenate([x[: n // 2 - 1000], x[n // 2 + 1500:], synth.letters(613, 2500)])
        exp = None
        if rank == 0 and n <= 60000:
            exp = pyoracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
        for block in blocks:
            epoch += 1
            log(kind, block, "setup")
            wave_trace_setup(eng, y, x, rank=rank, world=world, block_cols=block)
            log("fill")
            t0 = time.perf_counter()
            eng.wave_fill(epoch)
            score = eng.wave_fetch()
            t1 = time.perf_counter()
            log("filled", score, f"{eng.timing()['align_calc']:.3f} ms")
            tr = wave_trace(eng, rank=rank, world=world, cap=1 << 21)
            t2 = time.perf_counter()
            log("traced", tr[2] if tr else None)
            if rank == 0:
                rec = {"n": n, "kind": kind, "block": block, "world": world, "fill_ms": round(eng.timing()["align_calc"], 3), "fill_wall_ms": round((t1 - t0) * 1e3, 3),
                       "gather_trace_wall_ms": round((t2 - t1) * 1e3, 3), "trace_calc_ms": round(eng.timing()["trace_calc"], 3), "info": tr[2],
                       "trace_hash": f"{tr[1]:08x}", "edit_sha256": hashlib.sha256(tr[0].encode()).hexdigest()}
                if exp is not None:
                    rec["matches_oracle"] = bool(tr[0] == exp.edit and tr[1] == exp.trace_hash)
                print(json.dumps(rec), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
