#!/bin/bash
# round-2 experiment 1: Latin list layout / doubled z: parity, timings, ncu
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
V=autopas_b200/csrc/build/variants
for wl in c2 c3; do
  echo "== $wl legacy order"; APB_NO_LIST_SCHEDULE=1 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl latin16";      python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl latin16+zdup"; APB_LIB_PATH=$V/lib_zdup.so python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
ncu --set full --clock-control none --import-source on -k regex:kLJPruned -s 3 -c 1 -f -o gpurun_out/r02_kLJPruned_latin_c2 python tools/force_only.py 32 3 c2 > gpurun_out/ncu1.log 2>&1
APB_LIB_PATH=$V/lib_zdup.so ncu --set full --clock-control none --import-source on -k regex:kLJPruned -s 3 -c 1 -f -o gpurun_out/r02_kLJPruned_zdup_c2 python tools/force_only.py 32 3 c2 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kLJPruned -s 3 -c 1 -f -o gpurun_out/r02_kLJPruned_latin_c3 python tools/force_only.py 32 3 c3 > gpurun_out/ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --workload c3 --no-c2 --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c2_latin.json 2> gpurun_out/r02_bench_c2_latin.err
