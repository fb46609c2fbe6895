#!/bin/bash
# relative-drift criterion (SPHB200_REL_DRIFT=1): completeness test, duo suite with it on, A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_duo.py -m gpu -x -q -k "frozen_lists" --durations=3 > gpurun_out/r3r_frozen.log 2>&1; tail -8 gpurun_out/r3r_frozen.log
SPHB200_REL_DRIFT=1 timeout 600 python -m pytest tests/test_gpu_duo.py tests/test_gpu_reference.py -m gpu -q > gpurun_out/r3r_tests_rel.log 2>&1; tail -5 gpurun_out/r3r_tests_rel.log
for v in 1 0; do
  SPHB200_REL_DRIFT=$v timeout 400 python bench.py --steps 100 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs $BENCH_ARGS > gpurun_out/r3r_bench_rel$v.json 2> gpurun_out/r3r_bench_rel$v.err
  python - $v <<'PY'
import json,sys
f="gpurun_out/r3r_bench_rel%s.json"%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("rel",sys.argv[1], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["passes_ms"].items()}, "searches", d["config"]["neighbour_search"]["searches"], "of", d["config"]["neighbour_search"]["steps"], "err", d["device_error_word"], "stateless", round(d["stateless_advance"]["ms_per_step"],2), round(d["stateless_advance"]["engine_order"]["ms_per_step"],2))
except Exception as e:
    print(sys.argv, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done
