#!/bin/bash
# experiment 16 (8 GPUs): strong-scaling bench line at N=8 and N=4, SPH across 8 ranks, LJ decomposition parity at 8 ranks
cd /root/repo
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR8 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_v3_n8.json 2> gpurun_out/r02_bench_v3_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v3_n8.json')); print(8, d['value'], d['phases_ms_per_step'], d['e2e']['value'])"
timeout 300 $TR8 --master-port 29512 tools/multi_gpu_sph.py --check 2>&1 | grep -E "ranks|CHECK|rror" | tail -3
timeout 300 $TR8 --master-port 29513 tools/multi_gpu_sph.py 2>&1 | grep -E "^\{|rror" | tee gpurun_out/exp16_sph_n8.json | cut -c1-400
timeout 300 $TR8 --master-port 29514 tools/multi_gpu_check.py 2>&1 | grep -E "ranks|CHECK|rror" | tail -3
timeout 600 $TR4 --master-port 29516 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_v3_n4.json 2> gpurun_out/r02_bench_v3_n4.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v3_n4.json')); print(4, d['value'], d['phases_ms_per_step'], d['e2e']['value'])"
