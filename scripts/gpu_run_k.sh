#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/duo_check.py 48 40 > gpurun_out/r2k_duo_check.log 2>&1
grep -v "plan" gpurun_out/r2k_duo_check.log | tail -30
timeout 600 python bench.py --steps 20 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
f="gpurun_out/r2k_bench.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["roofline"]["passes_ms"], d["config"]["plan"], d["config"]["neighbour_search"], d["device_error_word"])
except Exception as e:
    print(f, "FAILED", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_duo -s 0 -c 4 -f -o gpurun_out/r2k_prof python bench.py --nx 128 --steps 3 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log | cut -c1-300
