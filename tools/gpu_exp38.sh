#!/bin/bash
# experiment 38: compute-sanitizer memcheck over the round-2 kernels (functor variants, lists, wire format, chained scan,
# rebuild chain of the smoke path)
cd /root/repo
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_functors.py tests/test_wire_format.py -m gpu -q -k "not full_size and not deleted_since" > gpurun_out/exp38_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/exp38_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit|Invalid|Error" gpurun_out/exp38_memcheck.log | head -20
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/exp38_memcheck_smoke.log 2>&1
echo "smoke memcheck exit $?" >> gpurun_out/exp38_memcheck_smoke.log
grep -E "ERROR SUMMARY|OK|memcheck exit|Invalid" gpurun_out/exp38_memcheck_smoke.log | head
