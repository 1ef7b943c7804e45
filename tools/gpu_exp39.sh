#!/bin/bash
# experiment 39: compute-sanitizer memcheck over the parity / dynamics / control tests (small sizes)
cd /root/repo
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dynamics.py tests/test_gpu_control.py -m gpu -q -k "not full_size and not rebuild_timing" > gpurun_out/exp39_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/exp39_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit|Invalid|Error" gpurun_out/exp39_memcheck.log | head -20
