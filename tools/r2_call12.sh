#!/bin/bash
set -x
for p in "" 1; do
  echo "== NO_PRIORITY='$p' 32 tracks";  NODEY_NO_STREAM_PRIORITY=$p T=32 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
  echo "== NO_PRIORITY='$p' 256 tracks"; NODEY_NO_STREAM_PRIORITY=$p T=256 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
  echo "== NO_PRIORITY='$p' e2e";        NODEY_NO_STREAM_PRIORITY=$p timeout 300 python tools/e2e_diag.py 2>&1 | head -4 | tail -2
done
for c in 8 24 32; do echo "== chunks $c 32 tracks"; NODEY_ST_CHUNKS=$c T=32 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2; done
