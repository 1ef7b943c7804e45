#!/bin/bash
# experiment 44: Axilrod-Teller warp / two-pass kernels on one shared triplet function
cd /root/repo
for v in warp list; do echo "== $v"; APB_LC_KERNEL=$v timeout 600 python -m pytest tests/test_gpu_functors.py tests/test_gpu_shim.py -m gpu -q -k "atm or ATM or shim or axilrod" 2>&1 | tail -2; done
echo "== inline"; APB_ATM_INLINE=1 timeout 600 python -m pytest tests/test_gpu_functors.py -m gpu -q -k "atm or ATM or axilrod" 2>&1 | tail -2
echo "== default"; timeout 900 python -m pytest tests/test_gpu_functors.py tests/test_gpu_full_size.py -m gpu -q -k "not c3_full and not c5_full" 2>&1 | tail -2
timeout 300 python tools/bench_functors.py c4 2>/dev/null | cut -c1-200
