#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
for c in 1 8; do echo "== chunks $c"; NODEY_ST_CHUNKS=$c T=32 timeout 300 python tools/chain_trace.py; done > gpurun_out/r2_chain_trace.txt 2>&1
cat gpurun_out/r2_chain_trace.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"])
print("kernels", d["roofline"]["kernels_ms"])
for k,v in (d["configs"] or {}).items():
    print(k, "ms", round(v["ms"],3), "rt", round(v["value"]), "e2e ms", round(v["e2e"]["ms"],2), "cpu", round(v["cpu_baseline"]["value"],1), "roof", v["roofline"]["kernel"], round(v["roofline"]["frac"],3), v["parity"])
print("segments", d["segments"])
PY
for T in 32 64 128; do NODEY_ST_CHUNKS=8 timeout 600 python bench.py --tracks $T --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-parity --no-e2e > gpurun_out/r2_bench${T}.json 2> gpurun_out/r2_bench${T}.err; grep -h "bench\]" gpurun_out/r2_bench${T}.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench${T}.json').read().strip().splitlines()[-1]);print($T,'tracks ms',d['ms_per_step'])"; done
