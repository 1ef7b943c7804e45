#!/bin/bash
# experiment 19: masked Axilrod-Teller kernel with entry-major partner positions; SPH hydro at 8 blocks per SM
cd /root/repo
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_functors.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -3
echo "== parity, inline ATM"; APB_ATM_INLINE=1 timeout 1200 python -m pytest tests/test_gpu_functors.py -m gpu -x -q -k "atm or ATM or axilrod" 2>&1 | tail -3
echo "== defaults"; timeout 600 python tools/bench_functors.py c1 c4 c5 2>/dev/null | tee gpurun_out/exp19_functors.jsonl | cut -c1-330
echo "== ATM inline"; APB_ATM_INLINE=1 timeout 600 python tools/bench_functors.py c4 2>/dev/null | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"kATMTripletsMasked" -c 2 -o gpurun_out/r02_atm_masked2 python tools/bench_functors.py c4 > gpurun_out/exp19_ncu.log 2>&1; tail -1 gpurun_out/exp19_ncu.log
