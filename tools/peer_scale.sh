#!/bin/bash
# Next-round measurement: the master mix over NVLink peer memory against the NCCL reduce at N ranks.
# usage on an N-GPU box (gpurun --gpus N):  bash tools/peer_scale.sh N [tag]
# Writes gpurun_out/<tag>_bench_n<N>_{nccl,peer}.json (one JSON line each; never run under a profiler) and the
# two-rank bit-exactness test's log.
set -u
n=${1:-2}
tag=${2:-r2}
out=gpurun_out
mkdir -p $out
port=29520
for bus in nccl peer; do
    port=$((port + 1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus "$n" --steps 3 --warmup 3 --bus $bus > $out/${tag}_bench_n${n}_${bus}.json 2> $out/${tag}_bench_n${n}_${bus}.err || exit 1
    python - "$out/${tag}_bench_n${n}_${bus}.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"{d['config']['collective'][:40]:42s} N={d['n_gpus']}: {d['value'] / 1e3:.1f} k audio-s/s, {d['ms_per_step']:.1f} ms per step, e2e {d['e2e']['value'] / 1e3:.1f} k")
PY
done
python -m pytest tests/test_gpu_peer.py tests/test_gpu_bus.py -q > $out/${tag}_peer_bus_pytest.log 2>&1; tail -1 $out/${tag}_peer_bus_pytest.log
