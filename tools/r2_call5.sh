#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/e2e_diag.py > gpurun_out/r2_e2e_diag.txt 2>&1; head -80 gpurun_out/r2_e2e_diag.txt
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 600 python tools/e2e_diag.py 2>&1 | head -6
NODEY_ST_CHUNKS=1 timeout 600 python tools/e2e_diag.py 2>&1 | head -6
T=256 timeout 300 python tools/chain_trace.py > gpurun_out/r2_chain_trace256.txt 2>&1; cat gpurun_out/r2_chain_trace256.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
for f in ("r2_bench_n1",):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f,"ms", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), "parity", d["parity"] and d["parity"]["ok"], d["roofline"]["kernels_ms"])
PY
