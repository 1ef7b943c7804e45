#!/bin/bash
# experiment 25: sort keys scattered into bin order (kBinSort without gathers), SPH lists vs deleted particles
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for wl in c2 c3; do
  echo "== $wl";  timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v5.json 2> gpurun_out/r02_bench_v5.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v5.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_16M_v5.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 2 > /dev/null 2>&1
grep -E "kBinSort|kScatterPermBins" gpurun_out/r02_launches_c3_16M_v5.csv | head -4 | cut -c1-60,200-400
