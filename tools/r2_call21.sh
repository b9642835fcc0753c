#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_soundtouch.py -m gpu -q -x 2>&1 | tail -2
SECS=30 python tools/tds_phases.py 2>&1 | grep "N=32 CL=4\|N=64 CL=4\|N=256\|rror"
for T in 256 64 32; do echo "T=$T"; T=$T timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2; done
