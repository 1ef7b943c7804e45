#!/bin/bash
# experiment 45: device-side VTK checkpoint record (f4) + whole GPU suite + bench line at HEAD
cd /root/repo
echo "== vtk"; timeout 300 python -m pytest tests/test_vtk.py -m gpu -q -x 2>&1 | tail -15
echo "== suite"; timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_vtk.py 2>&1 | tail -4
timeout 120 python tools/bench_vtk.py 1000000 4000000 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp45_bench.json 2> gpurun_out/exp45_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp45_bench.json')); print(d['value'], d['phases_ms_per_step'], d['c2']['value'], d['e2e']['value'], [ (o.get('config'), o.get('ms_per_call')) for o in d.get('other_configs', [])])"
