#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
T=32 timeout 300 python tools/chain_trace.py > gpurun_out/r2_chain_trace.txt 2>&1
cat gpurun_out/r2_chain_trace.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
NODEY_WAVE=100000 timeout 900 python bench.py --steps 5 --warmup 3 --no-configs --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_n1_onewave.json 2> gpurun_out/r2_bench_n1_onewave.err
tail -2 gpurun_out/r2_bench_n1_onewave.err
timeout 600 python bench.py --tracks 32 --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err
tail -2 gpurun_out/r2_bench32.err
python - <<'PY'
import json
for f in ("r2_bench_n1","r2_bench_n1_onewave","r2_bench32"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f,"ms", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), "parity", d["parity"] and d["parity"]["ok"], d["roofline"]["kernels_ms"])
PY
