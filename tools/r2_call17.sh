#!/bin/bash
for T in 36 40 48; do for cl in 2 4; do
echo "== T=$T NODEY_TDS_CLUSTER=$cl"
T=$T NODEY_TDS_CLUSTER=$cl timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -1
done; done
for T in 72 80 96 112; do for cl in 1 2; do
echo "== T=$T NODEY_TDS_CLUSTER=$cl"
T=$T NODEY_TDS_CLUSTER=$cl timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -1
done; done
