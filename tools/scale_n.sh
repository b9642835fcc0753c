#!/bin/bash
# One N-GPU pass (gpurun --gpus N -- bash tools/scale_n.sh N): the multi-rank tests, then bench.py under torchrun.
set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=name --format=csv | head -3
timeout 600 python -m pytest tests/test_gpu_bus.py tests/test_gpu_peer.py -m gpu -q > gpurun_out/r2g_pytest_n${N}.log 2>&1; tail -3 gpurun_out/r2g_pytest_n${N}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2g_bench_n${N}.json 2> gpurun_out/r2g_bench_n${N}.err
echo "bench rc=$?"; grep "bench\]" gpurun_out/r2g_bench_n${N}.err | tail -4; tail -3 gpurun_out/r2g_bench_n${N}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_bench_n${N}.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"])
print("segments", d["segments"])
PY
