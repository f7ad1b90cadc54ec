python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/wave_probe.py 200000 8:2:100000 8:2:50016 8:2:25024 8:2:12512 2>/dev/null | cut -c1-130
