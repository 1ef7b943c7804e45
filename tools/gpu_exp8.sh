#!/bin/bash
cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kPrunedMasks -s 1 -c 1 -f -o gpurun_out/r02_kPrunedMasks_c3 python tools/force_only.py 32 2 c3 > gpurun_out/ncu_masks.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kPrunedFill -s 1 -c 1 -f -o gpurun_out/r02_kPrunedFill_c2 python tools/force_only.py 32 2 c2 > gpurun_out/ncu_fill.log 2>&1
ls -la gpurun_out
