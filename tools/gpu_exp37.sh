#!/bin/bash
# experiment 37 (2 GPUs): decomposed runs after the column swap (migration carries force and oldForce)
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_v8_n2.json 2> gpurun_out/r02_bench_v8_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v8_n2.json')); print(2, d['value'], d['phases_ms_per_step'], d['e2e']['value'])"
