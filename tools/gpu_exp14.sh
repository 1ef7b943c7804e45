#!/bin/bash
# experiment 14 (2 GPUs): SPH across ranks (parity + timing), LJ decomposition regression, strong-scaling bench line
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tools/multi_gpu_sph.py --check 2>&1 | grep -E "ranks|CHECK|rror" | tail -5
timeout 300 $TR --master-port 29513 tools/multi_gpu_sph.py 2>&1 | grep -E "^\{|rror" | tee gpurun_out/exp14_sph_n2.json | cut -c1-400
timeout 300 python tools/multi_gpu_sph.py 2>&1 | grep -E "^\{|rror" | tee gpurun_out/exp14_sph_n1.json | cut -c1-400
timeout 300 $TR --master-port 29514 tools/multi_gpu_check.py 2>&1 | grep -E "ranks|CHECK|rror" | tail -5
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/exp14_bench_n2.json 2> gpurun_out/exp14_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/exp14_bench_n2.json')); print(d['value'], d['phases_ms_per_step'], d['e2e']['value'])"
