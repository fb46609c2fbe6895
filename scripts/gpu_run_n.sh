#!/bin/bash
# round-2 experiment: A/B of environment switches on the headline bench (passes per step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 100 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs "$@" > gpurun_out/r3n_$name.json 2> gpurun_out/r3n_$name.err
  python - "$name" <<'PY'
import json,sys
f="gpurun_out/r3n_%s.json"%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    p=d["roofline"]["passes_ms"]; c=d["config"]
    print(sys.argv[1], round(d["ms_per_step"],3), {k:round(v,3) for k,v in p.items()}, c["plan"]["tile"], c["plan"]["threads"], c["neighbour_search"]["searches"], c["neighbour_search"]["tiles_without_lists"], d["device_error_word"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f.replace(".json",".err")).read()[-600:])
PY
}
if [ -n "$TESTS" ]; then timeout 900 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -5; fi
while read -r line; do [ -n "$line" ] && run $line; done <<< "$RUNS"
