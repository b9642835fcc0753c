#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2f_pytest_gpu.log
python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -3 gpurun_out/r2f_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; tail -c 600 gpurun_out/r2f_bench_reference.json
