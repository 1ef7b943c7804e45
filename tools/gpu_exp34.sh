#!/bin/bash
# experiment 34 (4 GPUs): split step + force overwrite + two-stage statistics reduction together
cd /root/repo
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR4 --master-port 29514 tools/multi_gpu_check.py 2>&1 | grep -E "ranks|CHECK|rror" | tail -3
timeout 300 $TR4 --master-port 29512 tools/multi_gpu_sph.py --check 2>&1 | grep -E "ranks|CHECK|rror" | tail -3
timeout 600 $TR4 --master-port 29516 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_v7_n4.json 2> gpurun_out/r02_bench_v7_n4.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v7_n4.json')); print(4, d['value'], d['phases_ms_per_step'], d['e2e']['value'])"
