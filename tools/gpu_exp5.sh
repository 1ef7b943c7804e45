#!/bin/bash
# round-2 session-2 experiment 5: GPU test suite, tile size 16 warps, ncu of the list-refinement kernels at C3 density
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
V=autopas_b200/csrc/build/variants
for wl in c2 c3; do
  echo "== $wl 8 warps";  python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl 8 warps latin";  APB_LIST_SCHEDULE=1 python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl 16 warps"; APB_LIB_PATH=$V/lib_w16.so python tools/force_only.py 32 30 $wl 2>&1 | tail -1
  echo "== $wl 16 warps latin"; APB_LIST_SCHEDULE=1 APB_LIB_PATH=$V/lib_w16.so python tools/force_only.py 32 30 $wl 2>&1 | tail -1
done
for k in kPrunedMasks kPrunedFill kPrunedStage; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_${k}_c3 python tools/force_only.py 32 2 c3 > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:kLJPruned -s 3 -c 1 -f -o gpurun_out/r02_kLJPruned_w16_c3 env APB_LIB_PATH=$V/lib_w16.so python tools/force_only.py 32 3 c3 > gpurun_out/ncu_w16.log 2>&1
