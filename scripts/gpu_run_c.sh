#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity3d.py tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_api.py -m gpu -q --maxfail=25 > gpurun_out/r2e_tests.log 2>&1
tail -15 gpurun_out/r2e_tests.log
python bench.py --steps 20 --warmup 3 --e2e-steps 2 --cpu-steps 1 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python - <<'PY'
import json
f="gpurun_out/r2e_bench.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["passes_ms"], d["config"]["neighbour_search"]["searches"], d["e2e"]["value"])
except Exception as e:
    print(f, "FAILED", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 0 -c 3 -f -o gpurun_out/r2e_prof python bench.py --nx 128 --steps 3 --warmup 3 --e2e-steps 1 --cpu-steps 1 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log | cut -c1-300
