#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_duo.py tests/test_gpu_parity3d.py tests/test_gpu_reference.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
run() { tag=$1; shift; env "${ENVV[@]}" python bench.py --steps 60 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err
python - <<PY
import json
f="gpurun_out/r2l_$tag.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("$tag", round(d["value"]/1e6,1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["roofline"]["passes_ms"].items()}, d["config"]["plan"]["tile"], d["config"]["plan"]["threads"], d["config"]["neighbour_search"]["searches"], d["config"]["neighbour_search"]["tiles_without_lists"])
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
ENVV=(A=1); run base
ENVV=(A=1); run s12 --skin 0.12
ENVV=(A=1); run s08 --skin 0.08
