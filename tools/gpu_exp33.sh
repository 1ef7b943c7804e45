#!/bin/bash
# experiment 33: two-stage partial reduction from 2048 partials on
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in c2 c3; do
  echo "== $wl";  timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/exp33_bench.json 2> gpurun_out/exp33_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp33_bench.json')); print(d['value'], d['phases_ms_per_step'], d['c2']['value'])"
