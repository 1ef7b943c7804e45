#!/bin/bash
# experiment 43: SPH hydro kernels on one shared pair function - parity of the three variants, timing
cd /root/repo
for v in thread warp list; do echo "== $v"; APB_LC_KERNEL=$v timeout 600 python -m pytest tests/test_gpu_functors.py tests/test_gpu_shim.py -m gpu -q -k "sph or SPH or shim" 2>&1 | tail -2; done
echo "== default suite (functors, full size, multi-functor shim)"; timeout 900 python -m pytest tests/test_gpu_functors.py tests/test_gpu_full_size.py -m gpu -q -k "not c3_full" 2>&1 | tail -2
timeout 300 python tools/bench_functors.py c5 2>/dev/null | cut -c1-200
