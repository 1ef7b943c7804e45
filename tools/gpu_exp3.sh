#!/bin/bash
cd /root/repo
export APB_NO_LIST_SCHEDULE=1
for wl in c2 c3; do for M in 4 8 16 32; do python tools/force_only.py $M 20 $wl 2>&1 | tail -1; done; done
