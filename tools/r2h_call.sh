#!/bin/bash
# experiment: small-batch regime of the WSOLA search; arguments: tracks[:cluster] ...
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_soundtouch.py -m gpu -x -q > $out/r2h_pytest_st.log 2>&1; echo "pytest rc=$?"; tail -n 2 $out/r2h_pytest_st.log
q="--steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-parity --no-e2e"
for item in "$@"; do
  tr=${item%%:*}; cl=""; [[ $item == *:* ]] && cl=${item##*:}
  if [ -n "$cl" ]; then export NODEY_TDS_CLUSTER=$cl; else unset NODEY_TDS_CLUSTER; fi
  timeout 300 python bench.py --tracks $tr $q 2> $out/r2h_t${tr}_${cl}.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tracks $tr cluster ${cl:-auto}: %.2f ms' % d['ms_per_step'])"
done
