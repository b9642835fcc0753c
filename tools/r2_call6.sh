#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/e2e_diag.py > gpurun_out/r2_e2e_diag.txt 2>&1; head -12 gpurun_out/r2_e2e_diag.txt; tail -12 gpurun_out/r2_e2e_diag.txt
for c in 2 4 16; do echo "== e2e chunks $c"; NODEY_ST_CHUNKS=$c timeout 600 python tools/e2e_diag.py 2>&1 | head -4; done
for c in 2 4 8 16; do for w in "" 100000; do echo "== value chunks $c wave '$w'"; NODEY_ST_CHUNKS=$c NODEY_WAVE=$w T=256 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2; done; done
for c in 4 8 16; do echo "== 32 tracks chunks $c"; NODEY_ST_CHUNKS=$c T=32 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2; done
