#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_slab.py tests/test_gpu_api.py -m gpu -q --maxfail=25 > gpurun_out/r2g_tests.log 2>&1
tail -25 gpurun_out/r2g_tests.log
