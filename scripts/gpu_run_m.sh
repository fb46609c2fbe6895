#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python scripts/validate_dambreak.py 0.02 7.3 > gpurun_out/r02_validation_dambreak.txt 2> gpurun_out/r2m_db.err
head -12 gpurun_out/r02_validation_dambreak.txt | cut -c1-400; tail -3 gpurun_out/r2m_db.err
timeout 600 python -m pytest tests/test_gpu_validation.py tests/test_gpu_duo.py tests/test_simulate.py -m gpu -q -p no:cacheprovider -k "dam_break or duo or state0" 2>&1 | tail -15 | cut -c1-250
