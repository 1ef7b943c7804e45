#!/bin/bash
# experiment 28: default bench line with other_configs + their CPU reference
cd /root/repo
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v6.json 2> gpurun_out/r02_bench_v6.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v6.json')); print(d['value'], d['phases_ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value']); print([ (o.get('config','')[:16], round(o.get('ms_per_call',0),3), o.get('cpu_reference',{}).get('ms_per_call')) for o in d['other_configs']])"
tail -3 gpurun_out/r02_bench_v6.err
