#!/bin/bash
# experiment 9: lane-major list assembly in kPrunedFill (9 instructions per entry), WS kernel removed
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for wl in c2 c3; do
  echo "== $wl";  timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl latin";  APB_LIST_SCHEDULE=1 timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
for wl in c2 c3; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/exp9_launches_$wl.csv python tools/force_only.py 32 3 $wl > /dev/null 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp9_bench.json 2> gpurun_out/exp9_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp9_bench.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['e2e']['value'])"
