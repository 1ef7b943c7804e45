#!/bin/bash
# experiment 17: SPH list kernels with batched gathers (hydro batch 1 / 2 / 4)
cd /root/repo
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_functors.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -3
echo "== hydro batch 2 (default)"; timeout 600 python tools/bench_functors.py c5 2>/dev/null | tee gpurun_out/exp17_c5.jsonl | cut -c1-260
for b in 1 4; do echo "== hydro batch $b"; APB_LIB_PATH=/root/repo/autopas_b200/csrc/build/variants/lib_hb$b.so timeout 600 python tools/bench_functors.py c5 2>/dev/null | cut -c1-200; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"kSPHHydroList" -c 2 -o gpurun_out/r02_sph_hydro_list python tools/bench_functors.py c5 > gpurun_out/exp17_ncu.log 2>&1; tail -1 gpurun_out/exp17_ncu.log
