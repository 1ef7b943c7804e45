#!/bin/bash
# experiment 18: masked two-pass Axilrod-Teller kernel; occupancy variants of the SPH hydro list kernel
cd /root/repo
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_functors.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -3
echo "== ATM masked (default)"; timeout 600 python tools/bench_functors.py c4 2>/dev/null | tee gpurun_out/exp18_c4.jsonl | cut -c1-330
echo "== ATM inline"; APB_ATM_INLINE=1 timeout 600 python tools/bench_functors.py c4 2>/dev/null | cut -c1-330
echo "== hydro minblocks 4 (default)"; timeout 600 python tools/bench_functors.py c5 2>/dev/null | tee gpurun_out/exp18_c5.jsonl | cut -c1-200
for b in 5 6 8; do echo "== hydro minblocks $b"; APB_LIB_PATH=/root/repo/autopas_b200/csrc/build/variants/lib_mb$b.so timeout 600 python tools/bench_functors.py c5 2>/dev/null | tail -1 | cut -c1-200; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"kATMTripletsMasked" -c 2 -o gpurun_out/r02_atm_masked python tools/bench_functors.py c4 > gpurun_out/exp18_ncu.log 2>&1; tail -1 gpurun_out/exp18_ncu.log
