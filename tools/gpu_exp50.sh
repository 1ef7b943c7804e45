#!/bin/bash
# experiment 50: checkpoint loader on the device (parity with the unmodified reference loader, round trips), bench line with the checkpoint entry
cd /root/repo
echo "== vtk"; timeout 300 python -m pytest tests/test_vtk.py -m gpu -q 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp50_bench.json 2> gpurun_out/exp50_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp50_bench.json')); print(d['value'], d['phases_ms_per_step'], d['c2']['value'], d['e2e']['value'], d.get('checkpoint'))"
