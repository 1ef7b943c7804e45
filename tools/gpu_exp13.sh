#!/bin/bash
# experiment 13: partner-list SPH kernels, chained scan, halo column refresh
cd /root/repo
echo "== parity default"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== parity list"; APB_LC_KERNEL=list timeout 1200 python -m pytest tests/test_gpu_functors.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -4
echo "== list"
APB_LC_KERNEL=list timeout 600 python tools/bench_functors.py c5 2> gpurun_out/exp13_functors_list.err | tee gpurun_out/exp13_functors_list.jsonl | cut -c1-200
echo "== sph one gpu"
timeout 300 python tools/multi_gpu_sph.py --check 2>&1 | tail -3
APB_LC_KERNEL=list timeout 300 python tools/multi_gpu_sph.py 2>&1 | tail -1
for wl in c2 c3; do
  echo "== $wl";  timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp13_bench.json 2> gpurun_out/exp13_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp13_bench.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['e2e']['value'])"
