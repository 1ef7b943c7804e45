#!/bin/bash
# experiment 26: force-overwrite mode of apb_run_steps, SPH lists vs deleted particles, bench line with other_configs
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== parity without overwrite"; APB_NO_FORCE_OVERWRITE=1 timeout 900 python -m pytest tests/test_gpu_dynamics.py tests/test_gpu_control.py -m gpu -q 2>&1 | tail -2
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v5.json 2> gpurun_out/r02_bench_v5.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v5.json')); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['e2e']['value'], d['cpu_baseline']['value']); print([ (o['config'][:14], round(o['ms_per_call'],3)) for o in d['other_configs']])"
APB_NO_FORCE_OVERWRITE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c2 --e2e-steps 2 > gpurun_out/exp26_bench_noow.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/exp26_bench_noow.json')); print('no overwrite', d['value'], d['phases_ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_16M_v5.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-c2 --e2e-steps 2 > /dev/null 2>&1
wc -l gpurun_out/r02_launches_c3_16M_v5.csv
