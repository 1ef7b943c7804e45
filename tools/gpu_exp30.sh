#!/bin/bash
# experiment 30: kBinSort with a z-only fast pass
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/exp30_bench.json 2> gpurun_out/exp30_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp30_bench.json')); print(d['value'], d['phases_ms_per_step'], d['c2']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_16M_v6.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 2 --no-other-configs > /dev/null 2>&1
grep -E "kBinSort" gpurun_out/r02_launches_c3_16M_v6.csv | head -4 | cut -c1-60,200-400
