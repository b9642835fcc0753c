#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_soundtouch.py -m gpu -x -q 2>&1 | tail -n 2
N=256 SECS=60 timeout 200 python tools/prof_st_time.py 2>&1 | tail -n 1
