#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stft.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
bash tools/profile_round.sh r2 2>&1 | tail -8
ls -la gpurun_out | tail -20
