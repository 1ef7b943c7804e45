#!/bin/bash
# experiment 12: linked-cells kernel variants (thread / warp / deferred) on C1, C4, C5; parity of all three
cd /root/repo
for v in warp deferred; do
  echo "== parity with APB_LC_KERNEL=$v"
  APB_LC_KERNEL=$v timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
done
echo "== default"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in thread warp deferred; do
  echo "== $v"
  APB_LC_KERNEL=$v timeout 600 python tools/bench_functors.py c1 c4 c5 2> gpurun_out/exp12_functors_$v.err | tee gpurun_out/exp12_functors_$v.jsonl | cut -c1-200
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"kSPH.*Deferred|kATMTripletsDeferred" -c 6 -o gpurun_out/r02_functors_lc_deferred python tools/bench_functors.py c4 c5 > gpurun_out/exp12_ncu.log 2>&1
tail -2 gpurun_out/exp12_ncu.log
