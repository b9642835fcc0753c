#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_files.py tests/test_gpu_mp3.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -n 3
