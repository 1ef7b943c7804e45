#!/bin/bash
cd /root/repo
(time timeout 1500 python -m pytest tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -15)
