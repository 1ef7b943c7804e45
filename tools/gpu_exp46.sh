#!/bin/bash
# experiment 46: checkpoint record - timing at 1 M / 4 M / 16 M particles, memcheck of the small cases, ncu of the two kernels
cd /root/repo
timeout 200 python tools/bench_vtk.py 1000000 4000000 16003008 2>&1 | tail -4 | tee gpurun_out/r02_vtk_record.jsonl
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_vtk.py -m gpu -q -k "byte_exact or error_paths" > gpurun_out/exp46_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/exp46_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit|Invalid|Error" gpurun_out/exp46_memcheck.log | head -10
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"kVtkMeasure|kVtkWrite" -c 2 -o gpurun_out/r02_vtk_kernels python tools/bench_vtk.py 4000000 > gpurun_out/exp46_ncu.log 2>&1; tail -2 gpurun_out/exp46_ncu.log
