#!/bin/bash
# experiment 47: checkpoint record with one data array per blockIdx.y and the file writer: parity, shim, timing, ncu
cd /root/repo
echo "== vtk + shim"; timeout 300 python -m pytest tests/test_vtk.py tests/test_gpu_shim.py -m gpu -q -x 2>&1 | tail -8
timeout 200 python tools/bench_vtk.py 1000000 4000000 16003008 2>&1 | tail -4 | tee gpurun_out/r02_vtk_record.jsonl
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"kVtkMeasure|kVtkWrite" -s 2 -c 2 -o gpurun_out/r02_vtk_kernels python tools/bench_vtk.py 4000000 > gpurun_out/exp47_ncu.log 2>&1; tail -2 gpurun_out/exp47_ncu.log
