#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_duo.*PhysForce" -s 0 -c 1 -f -o gpurun_out/r2l_prof python bench.py --nx 128 --steps 3 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r2l_ncu.log 2>&1
tail -2 gpurun_out/r2l_ncu.log | cut -c1-300
run() { tag=$1; shift; env "${ENVV[@]}" python bench.py --steps 20 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err
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
ENVV=(SPHB200_DUO_TPB=288); run t644_288 --tile-x 6 --tile-y 4 --tile-z 4
ENVV=(SPHB200_DUO_TPB=256); run t544_256 --tile-x 5 --tile-y 4 --tile-z 4
ENVV=(SPHB200_DUO_TPB=384); run t844_384 --tile-x 8 --tile-y 4 --tile-z 4
ENVV=(SPHB200_DUO_TPB=512); run t844_512 --tile-x 8 --tile-y 4 --tile-z 4
ENVV=(SPHB200_DUO_TPB=352); run t853_352 --tile-x 8 --tile-y 5 --tile-z 3
ENVV=(SPHB200_DUO_TPB=352); run t844_s12 --skin 0.12
ENVV=(SPHB200_DUO_TPB=352); run t844_s15 --skin 0.15
