#!/bin/bash
# experiment 42: ncu --set full of the shipped Axilrod-Teller kernel and of the 18-pass integrator at 16 M
cd /root/repo
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"kATMTripletsMasked" -c 1 -o gpurun_out/r02_atm_final python tools/bench_functors.py c4 > gpurun_out/exp42_a.log 2>&1; tail -1 gpurun_out/exp42_a.log
timeout 400 ncu --set full --clock-control none -k regex:"kIntegrateVelocitiesPositions|kLJPruned" -c 4 -o gpurun_out/r02_step_kernels_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 1 --no-other-configs > gpurun_out/exp42_b.log 2>&1; tail -1 gpurun_out/exp42_b.log
