#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1
tail -30 gpurun_out/r2_tests.log | cut -c1-220
