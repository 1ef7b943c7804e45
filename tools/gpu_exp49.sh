#!/bin/bash
# experiment 49: checkpoint record with chunked scans (64-bit chunk bases): parity incl. small chunks, memcheck, smoke
cd /root/repo
echo "== vtk"; timeout 300 python -m pytest tests/test_vtk.py -m gpu -q 2>&1 | tail -12
timeout 150 env APB_VTK_CHUNK_ROWS=97 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_vtk.py -m gpu -q -k "byte_exact or error_paths" > gpurun_out/exp49_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/exp49_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit|Invalid|Error" gpurun_out/exp49_memcheck.log | head -10
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
