#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 0 -c 3 -f -o gpurun_out/r2b_prof python bench.py --nx 128 --steps 3 --warmup 3 --e2e-steps 1 --cpu-steps 1 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/r2b_tests.log 2>&1
tail -30 gpurun_out/r2b_tests.log
