#!/bin/bash
# experiment 31: validation at HEAD (all binaries rebuilt): GPU suite, smoke, default bench line, reference arm
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02_bench_v7.json 2> gpurun_out/r02_bench_v7.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v7.json')); print(d['value'], d['steps'], d['warmup'], d['phases_ms_per_step'], d['roofline']['frac'], d['c2']['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'])"
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_v7_reference.json 2> gpurun_out/r02_bench_v7_reference.err; cut -c1-200 gpurun_out/r02_bench_v7_reference.json
