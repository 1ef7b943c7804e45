#!/bin/bash
# experiment 23: group-cooperative cluster boxes and bin sort in the rebuild
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for wl in c2 c3; do
  echo "== $wl";  timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp23_bench.json 2> gpurun_out/exp23_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp23_bench.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['c2']['phases_ms_per_step'], d['e2e']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_16M_v4.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 2 > /dev/null 2>&1
grep -E "kBinSort|kClusterBoxes" gpurun_out/r02_launches_c3_16M_v4.csv | head -4 | cut -c1-60,200-400
