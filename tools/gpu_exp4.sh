#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for wl in c2 c3; do
  python tools/force_only.py 32 20 $wl 0 gpuvcl_pruned 2>&1 | tail -1
  python tools/force_only.py 32 20 $wl 0 gpuvcl_pruned n3 2>&1 | tail -1
done
