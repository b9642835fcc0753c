#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_soundtouch.py -m gpu -q -x 2>&1 | tail -2
bash tools/profile_round.sh r2f 2>&1 | tail -10
