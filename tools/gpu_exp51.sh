#!/bin/bash
# experiment 51: bench line with the checkpoint entry
cd /root/repo
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp51_bench.json 2> gpurun_out/exp51_bench.err
python -c "
import json; d=json.load(open('gpurun_out/exp51_bench.json')); print(d['value'], d['c2']['value'], d['e2e']['value'], d.get('checkpoint'))"
