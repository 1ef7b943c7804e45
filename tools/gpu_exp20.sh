#!/bin/bash
# experiment 20: staged sets found once per rebuild (counting pass keeps its result); fresh launch list of the C3 bench
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wl in c2 c3; do
  echo "== $wl";  timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp20_bench.json 2> gpurun_out/exp20_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp20_bench.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['c2']['value'], d['e2e']['value'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c3_16M_v2.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 2 > /dev/null 2>&1
wc -l gpurun_out/r02_launches_c3_16M_v2.csv
