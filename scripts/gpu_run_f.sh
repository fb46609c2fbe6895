#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py --steps 20 --warmup 3 --e2e-steps 1 --cpu-steps 1 "$@" > gpurun_out/r2f_$tag.json 2> gpurun_out/r2f_$tag.err
python - <<PY
import json
f="gpurun_out/r2f_$tag.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("$tag", round(d["value"]/1e6,1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["roofline"]["passes_ms"].items()}, d["config"]["plan"]["tile"], d["config"]["neighbour_search"]["searches"], d["config"]["neighbour_search"]["tiles_without_lists"])
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run t644
run t753 --tile-x 7 --tile-y 5 --tile-z 3
run t744 --tile-x 7 --tile-y 4 --tile-z 4
run t554 --tile-x 5 --tile-y 5 --tile-z 4
run t843 --tile-x 8 --tile-y 4 --tile-z 3
run t644_448 --threads 448
run t753_s12 --tile-x 7 --tile-y 5 --tile-z 3 --skin 0.12
run t644_s08 --skin 0.08
