#!/bin/bash
# round 2, first GPU pass: full GPU suite, then the bench with and without chunked chains, at 256 and 32 tracks
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
for chunks in 8 1 4 16; do
  NODEY_ST_CHUNKS=$chunks timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c${chunks}.json 2> gpurun_out/r2_bench_c${chunks}.err
  NODEY_ST_CHUNKS=$chunks timeout 600 python bench.py --tracks 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench32_c${chunks}.json 2> gpurun_out/r2_bench32_c${chunks}.err
done
grep -h "bench\]" gpurun_out/r2_bench*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench*_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "ms  e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), d["roofline"] and d["roofline"]["kernels_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
