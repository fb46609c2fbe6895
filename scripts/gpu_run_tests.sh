#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vjp.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_tests_vjp.log 2>&1
tail -30 gpurun_out/r2_tests_vjp.log | cut -c1-250
