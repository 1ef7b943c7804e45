#!/bin/bash
# experiment 40: masked Axilrod-Teller kernel with software-pipelined loads
cd /root/repo
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_functors.py tests/test_gpu_shim.py tests/test_gpu_full_size.py -m gpu -q -k "not c3_full and not c5_full" 2>&1 | tail -3
echo "== masked + pipelined (default)"; timeout 300 python tools/bench_functors.py c4 2>/dev/null | cut -c1-330
echo "== inline"; APB_ATM_INLINE=1 timeout 300 python tools/bench_functors.py c4 2>/dev/null | cut -c1-200
