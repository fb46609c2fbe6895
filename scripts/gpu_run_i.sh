#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo skip tests

python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
f="gpurun_out/r2j_bench.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["roofline"]["passes_ms"], d["config"]["neighbour_search"]["searches"], "e2e", round(d["e2e"]["value"]/1e6,1))
    print("stateless", d["stateless_advance"]["ms_per_step"], d["stateless_advance"]["scratch_workspace"]["ms_per_step"])
    for k,v in d["configs"].items(): print(k, v.get("n"), round(v.get("value",0)/1e6,1), v.get("ms_per_step"), v.get("searches"), v.get("steps_total"), v.get("device_error_word"), v.get("error"))
    r=d["roofline"]; print("roofline", r["kernel"][:40], r["bound"], r["achieved"], r["peak"], r["frac"], r["traffic"], r["step"])
except Exception as e:
    print(f, "FAILED", e)
PY
