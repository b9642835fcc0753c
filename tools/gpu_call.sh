#!/bin/bash
# one GPU-box call of a development session: the GPU test suite, then whatever the arguments name
# usage: bash tools/gpu_call.sh <tag> [big] [bench] [small]
set -u
tag=${1:-dev}; shift || true
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 $out/${tag}_pytest_gpu.log
for what in "$@"; do
  case $what in
    big)   timeout 600 python tools/big_render.py > $out/${tag}_big_render.json 2> $out/${tag}_big_render.err; echo "big rc=$?"; tail -c 1500 $out/${tag}_big_render.json; tail -n 5 $out/${tag}_big_render.err;;
    bench) timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?"; tail -n 3 $out/${tag}_bench_n1.err; head -c 600 $out/${tag}_bench_n1.json;;
    small) timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $out/${tag}_bench_quick.json 2> $out/${tag}_bench_quick.err; echo "quick rc=$?"; tail -n 3 $out/${tag}_bench_quick.err; head -c 400 $out/${tag}_bench_quick.json;;
  esac
done
