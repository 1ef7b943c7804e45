#!/bin/bash
# experiment 21: wire format (f4), SPH leaver test, staged-set stride, final single-GPU records
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
oracle/_ref/shim_test 2>&1 | tail -3
timeout 300 python tools/bench_wire.py 2>&1 | tail -1 | tee gpurun_out/exp21_wire.json | cut -c1-400
for wl in c2 c3; do
  echo "== $wl";  timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v4.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['c2']['value'], d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 600 python tools/bench_functors.py c1 c4 c5 2>/dev/null | tee gpurun_out/exp21_functors.jsonl | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_16M_v3.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 2 > /dev/null 2>&1
grep -c kPrunedStage gpurun_out/r02_launches_c3_16M_v3.csv
