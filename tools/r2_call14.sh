#!/bin/bash
set -x
mkdir -p gpurun_out




timeout 900 python -m pytest tests/test_gpu_soundtouch.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python tools/kt_sweep.py 2>&1 | tee gpurun_out/r2_kt_sweep.txt
T=256 timeout 300 python tools/chain_trace.py 2>&1 | head -4
T=32 timeout 300 python tools/chain_trace.py 2>&1 | head -4
