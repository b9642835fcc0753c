#!/bin/bash
set -x
mkdir -p gpurun_out
for conn in 8 16 32; do
  echo "== connections $conn: 32 tracks"; CUDA_DEVICE_MAX_CONNECTIONS=$conn T=32 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
  echo "== connections $conn: 256 tracks"; CUDA_DEVICE_MAX_CONNECTIONS=$conn T=256 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
  echo "== connections $conn: e2e"; CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python tools/e2e_diag.py 2>&1 | head -4 | tail -2
done
echo "== connections 32, 32 tracks, diag"; CUDA_DEVICE_MAX_CONNECTIONS=32 T=32 timeout 300 python tools/chain_trace.py 2>&1 | tail -12
echo "== connections 8, 32 tracks, diag"; CUDA_DEVICE_MAX_CONNECTIONS=8 T=32 timeout 300 python tools/chain_trace.py 2>&1 | tail -12
# st_post under ncu (source-level), 256 tracks x 60 s
N=256 SECS=60 python tools/prof_st.py
N=256 SECS=60 ncu --set full --clock-control none --import-source on -k regex:st_post -c 1 -o gpurun_out/r2_st_post -f python tools/prof_st.py > gpurun_out/r2_ncu_post.log 2>&1
tail -2 gpurun_out/r2_ncu_post.log
