#!/bin/bash
# last check of the round: duo + api tests, bench line, ncu capture (traffic) of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_duo.py tests/test_gpu_api.py -m gpu -q > gpurun_out/r02_last_tests.txt 2>&1; tail -3 gpurun_out/r02_last_tests.txt
timeout 400 python bench.py --steps 60 --warmup 3 > gpurun_out/r02_bench_tgv3d_256.json 2> gpurun_out/r02_bench.err
python - <<'PY'
import json
f="gpurun_out/r02_bench_tgv3d_256.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["roofline"]["passes_ms"], d["config"]["neighbour_search"]["searches"], "e2e", round(d["e2e"]["value"]/1e6,1), d["device_error_word"])
    s = d["stateless_advance"]; print("stateless", s["ms_per_step"], s["engine_order"]["ms_per_step"], s["scratch_workspace"]["ms_per_step"])
    for k,v in d.get("configs", {}).items(): print(k, round(v.get("value",0)/1e6,1), v.get("ms_per_step"), v.get("searches"), v.get("device_error_word"), v.get("error"))
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/r02_bench.err").read()[-1500:])
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_duo -s 0 -c 4 -f -o gpurun_out/r02_duo_tgv3d_256 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r02_ncu.log 2>&1
tail -1 gpurun_out/r02_ncu.log | cut -c1-200
