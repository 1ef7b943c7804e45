#!/bin/bash
# experiment 35 (8 GPUs): strong-scaling lines at HEAD for N = 8 and N = 2
cd /root/repo
for n in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_v7_n$n.json 2> gpurun_out/r02_bench_v7_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v7_n$n.json')); print($n, d['value'], d['phases_ms_per_step'], d['e2e']['value'])"
done
