#!/bin/bash
# experiment 10: shim wrappers of the SPH / multi-site functors, leaver columns; baseline ncu of the linked-cells functor kernels
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
oracle/_ref/shim_test 2>&1 | tail -12
oracle/_ref/shim_test_ms 2>&1 | tail -6
timeout 600 python tools/bench_functors.py c1 c4 c5 > gpurun_out/exp10_functors.jsonl 2> gpurun_out/exp10_functors.err
cat gpurun_out/exp10_functors.jsonl | cut -c1-400
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"kSPH|kATMTriplets|kLJLinkedCells" -c 6 -o gpurun_out/r02_functors_lc_baseline python tools/bench_functors.py c1 c4 c5 --small > gpurun_out/exp10_ncu.log 2>&1
tail -3 gpurun_out/exp10_ncu.log
