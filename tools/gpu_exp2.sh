#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for wl in c2 c3; do
  echo "== $wl legacy order"; APB_NO_LIST_SCHEDULE=1 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl latin16";      python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c2_latin2.json 2> gpurun_out/r02_bench_c2_latin2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_latin2.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'])"
