#!/bin/bash
# round-2 check: ordered stateless call, pipelined tile loop (SPHB200_DUO_PIPE=1) tests + A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_duo.py tests/test_gpu_slab.py -m gpu -x -q --durations=5 > gpurun_out/r3q_tests.log 2>&1; tail -4 gpurun_out/r3q_tests.log
SPHB200_DUO_PIPE=1 timeout 900 python -m pytest tests/test_gpu_duo.py tests/test_gpu_parity3d.py tests/test_gpu_reference.py tests/test_gpu_slab.py -m gpu -x -q --durations=5 > gpurun_out/r3q_tests_pipe.log 2>&1; tail -4 gpurun_out/r3q_tests_pipe.log
for v in 0 1; do
  SPHB200_DUO_PIPE=$v timeout 400 python bench.py --steps 40 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r3q_bench_pipe$v.json 2> gpurun_out/r3q_bench_pipe$v.err
  python - $v <<'PY'
import json,sys
f="gpurun_out/r3q_bench_pipe%s.json"%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); s=d["stateless_advance"]
    print("pipe",sys.argv[1], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["passes_ms"].items()}, "stateless", round(s["ms_per_step"],3), "ordered", round(s["engine_order"]["ms_per_step"],3), s["engine_order"]["device_error_word"], "scratch", round(s["scratch_workspace"]["ms_per_step"],3), "err", d["device_error_word"])
except Exception as e:
    print(sys.argv, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done
