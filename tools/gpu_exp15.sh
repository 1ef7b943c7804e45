#!/bin/bash
# experiment 15: packed SPH list kernels; ncu traffic capture of kLJPruned at C3 (16 M) and C2; bench records
cd /root/repo
echo "== parity"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== functors (defaults)"
timeout 600 python tools/bench_functors.py c1 c4 c5 2> gpurun_out/exp15_functors.err | tee gpurun_out/exp15_functors.jsonl | cut -c1-260
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"kSPH.*List" -c 4 -o gpurun_out/r02_sph_list python tools/bench_functors.py c5 > gpurun_out/exp15_ncu_sph.log 2>&1; tail -1 gpurun_out/exp15_ncu_sph.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"kLJPruned" -c 2 -o gpurun_out/r02_kLJPruned_c3_16M python tools/force_only.py 32 2 c3 252 > gpurun_out/exp15_ncu_c3.log 2>&1; tail -1 gpurun_out/exp15_ncu_c3.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"kLJPruned" -c 2 -o gpurun_out/r02_kLJPruned_c2 python tools/force_only.py 32 2 c2 > gpurun_out/exp15_ncu_c2.log 2>&1; tail -1 gpurun_out/exp15_ncu_c2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v3.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_v3_reference.json 2> gpurun_out/r02_bench_v3_reference.err; cut -c1-300 gpurun_out/r02_bench_v3_reference.json
