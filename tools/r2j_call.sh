#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
FAMILY=soundtouch timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_smoke.py > $out/r2j_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 3 $out/r2j_sanitize_memcheck.log
FAMILY=soundtouch timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_smoke.py > $out/r2j_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 3 $out/r2j_sanitize_racecheck.log
timeout 600 python -m pytest tests -m gpu -q > $out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 $out/r2j_pytest_gpu.log
