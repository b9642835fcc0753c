#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -15 gpurun_out/r2_pytest_gpu.log
for T in 32 256; do for NC in 8; do T=$T NC=$NC timeout 300 python tools/chain_overlap.py; done; done > gpurun_out/r2_chain_overlap.txt 2>&1
cat gpurun_out/r2_chain_overlap.txt
for c in 1 8; do echo "== chunks $c"; NODEY_ST_CHUNKS=$c T=32 timeout 300 python tools/chain_trace.py; done > gpurun_out/r2_chain_trace.txt 2>&1
cat gpurun_out/r2_chain_trace.txt
