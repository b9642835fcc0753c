#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
T=32 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
T=256 timeout 300 python tools/chain_trace.py 2>&1 | head -4 | tail -2
timeout 300 python tools/e2e_diag.py 2>&1 | head -4 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["ok"], "memory", d["memory"])
print("kernels", d["roofline"]["kernels_ms"], d["roofline"].get("fp32"))
for k,v in (d["configs"] or {}).items():
    print(k, "ms", round(v["ms"],3), "rt", round(v["value"]), "e2e ms", round(v["e2e"]["ms"],2), "cpu", round(v["cpu_baseline"]["value"],1), "roof", v["roofline"]["kernel"], round(v["roofline"]["frac"],3), v["parity"])
PY
timeout 600 python bench.py --tracks 32 --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err
tail -2 gpurun_out/r2_bench32.err
