#!/bin/bash
# experiment 24 (8 GPUs): final strong-scaling lines at HEAD (N = 8, 4, 2) and the 2-rank pytest wrappers
cd /root/repo
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_v5_n$n.json 2> gpurun_out/r02_bench_v5_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v5_n$n.json')); print($n, d['value'], d['phases_ms_per_step'], d['e2e']['value'])"
done
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
