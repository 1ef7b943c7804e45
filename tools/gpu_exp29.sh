#!/bin/bash
# experiment 29: ncu --set full of the rebuild / integration kernels at 16 M particles (HEAD)
cd /root/repo
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"kPrunedMasks|kPrunedFill|kBinSort|kIntegrateVelocitiesPositions|kGatherAll|kPrunedStage|kClusterBoxes|kKeysVCLBins" -c 14 -o gpurun_out/r02_rebuild_kernels_c3_16M python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 1 --no-other-configs > gpurun_out/exp29_ncu.log 2>&1
tail -2 gpurun_out/exp29_ncu.log
