#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -8 gpurun_out/r2_pytest_gpu.log
for w in 32 64 "32,64,64,64,32" "64,64,64,32,32" 48; do echo "== e2e waves $w"; NODEY_WAVES=$w timeout 600 python tools/e2e_diag.py 2>&1 | head -4 | tail -2; done
for l in 2 4; do echo "== e2e lanes $l"; NODEY_COMPUTE_LANES=$l timeout 600 python tools/e2e_diag.py 2>&1 | head -4 | tail -2; done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["ok"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["single_thread"]["value"])
print("roofline", {k: d["roofline"][k] for k in ("kernel","frac","launch_ms","share_of_step")}, d["roofline"].get("fp32"))
print("kernels", d["roofline"]["kernels_ms"])
for k,v in (d["configs"] or {}).items():
    print(k, "ms", round(v["ms"],3), "rt", round(v["value"]), "e2e ms", round(v["e2e"]["ms"],2), "cpu", round(v["cpu_baseline"]["value"],1), "roof", v["roofline"]["kernel"], round(v["roofline"]["frac"],3), v["parity"])
print("segments", d["segments"])
PY
timeout 600 python bench.py --tracks 32 --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err
tail -2 gpurun_out/r2_bench32.err
