#!/bin/bash
# experiment 36: force / oldForce column swap in the fused integrator
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/exp36_bench.json 2> gpurun_out/exp36_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp36_bench.json')); print(d['value'], d['phases_ms_per_step'], d['c2']['value'], d['upot_last'])"
