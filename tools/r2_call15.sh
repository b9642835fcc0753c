#!/bin/bash
mkdir -p gpurun_out
for c in 24 32 48 64; do
echo "== e2e NODEY_ST_CHUNKS=$c"
NODEY_ST_CHUNKS=$c timeout 300 python tools/e2e_diag.py 2>&1 | head -4 | tail -2
echo "== resident T=256 NODEY_ST_CHUNKS=$c"
T=256 NODEY_ST_CHUNKS=$c timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
echo "== resident T=32 NODEY_ST_CHUNKS=$c"
T=32 NODEY_ST_CHUNKS=$c timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
done
