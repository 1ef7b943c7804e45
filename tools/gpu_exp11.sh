#!/bin/bash
# experiment 11: warp-per-particle linked-cells kernels (LJ, SPH, Axilrod-Teller) against the thread-per-particle ones
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
oracle/_ref/shim_test 2>&1 | grep -v "^FAIL.*leaver" | tail -14
oracle/_ref/shim_test 2>&1 | grep "leaver" | head -3
echo "== thread per particle"
APB_LC_THREAD_KERNEL=1 timeout 600 python tools/bench_functors.py c1 c4 c5 2> gpurun_out/exp11_functors_thread.err | tee gpurun_out/exp11_functors_thread.jsonl | cut -c1-230
echo "== warp per particle"
timeout 600 python tools/bench_functors.py c1 c4 c5 2> gpurun_out/exp11_functors.err | tee gpurun_out/exp11_functors.jsonl | cut -c1-230
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"kSPH.*Warp|kATMTripletsWarp|kLJLinkedCellsWarp" -c 8 -o gpurun_out/r02_functors_lc_warp python tools/bench_functors.py c1 c4 c5 --small > gpurun_out/exp11_ncu.log 2>&1
tail -2 gpurun_out/exp11_ncu.log
