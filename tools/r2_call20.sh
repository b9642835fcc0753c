#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_soundtouch.py -m gpu -q -x -k "candidates or cluster" 2>&1 | tail -2
for lib in "" $PWD/tools/micro/libnodey_cuda_tds3.so; do
echo "=== lib: ${lib:-default (2 resident)}"
for T in 256 128 64 32; do
echo "T=$T"; NODEY_CUDA_LIB=$lib T=$T timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
done
NODEY_CUDA_LIB=$lib SECS=60 timeout 300 python tools/kt_sweep.py 2>&1 | grep "N= 256"
done
