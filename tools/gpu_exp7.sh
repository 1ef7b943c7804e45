#!/bin/bash
# experiment 7: warp-specialised persistent force kernel (kLJPrunedWS) against the one-tile-per-CTA kernel
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for wl in c2 c3; do
  echo "== $wl old kernel";  APB_PRUNED_WS=0 timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl ws";  APB_DEBUG=1 timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | grep -v "^\[apb\] pruned build" | tail -1
  APB_DEBUG=1 timeout 300 python tools/force_only.py 32 1 $wl 2>&1 | grep "pruned build" | tail -1
  echo "== $wl ws cap2048";  APB_PRUNED_WS_CAP=2048 timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl ws latin";  APB_LIST_SCHEDULE=1 timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl ws latin cap2048";  APB_PRUNED_WS_CAP=2048 APB_LIST_SCHEDULE=1 timeout 300 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp7_bench.json 2> gpurun_out/exp7_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp7_bench.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['e2e']['value'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kLJPrunedWS -s 3 -c 1 -f -o gpurun_out/r02_kLJPrunedWS_c3 python tools/force_only.py 32 3 c3 > gpurun_out/ncu_ws_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kLJPrunedWS -s 3 -c 1 -f -o gpurun_out/r02_kLJPrunedWS_c2 python tools/force_only.py 32 3 c2 > gpurun_out/ncu_ws_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kPrunedMasks -s 1 -c 1 -f -o gpurun_out/r02_kPrunedMasks_c3 python tools/force_only.py 32 2 c3 > gpurun_out/ncu_masks.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kPrunedFill -s 1 -c 1 -f -o gpurun_out/r02_kPrunedFill_c2 python tools/force_only.py 32 2 c2 > gpurun_out/ncu_fill.log 2>&1
