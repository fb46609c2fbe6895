#!/bin/bash
# round-2 GPU check A: full GPU test-suite, bench with / without skin, FP32 peak, ncu launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 -x --deselect tests/test_gpu_scale.py > gpurun_out/r2a_tests.log 2>&1
tail -5 gpurun_out/r2a_tests.log
python scripts/fp32_peak.py > gpurun_out/r2a_fp32_peak.json 2> gpurun_out/r2a_fp32_peak.err
cat gpurun_out/r2a_fp32_peak.json
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2a_bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["passes_ms"], d["config"]["neighbour_search"], d["e2e"]["value"])
    except Exception as e:
        print(f, "FAILED", e)
PY
for sk in -1 0.05 0.15 0.2; do
python bench.py --steps 20 --warmup 3 --skin $sk --e2e-steps 1 --cpu-steps 1 > gpurun_out/r2a_bench_skin$sk.json 2> gpurun_out/r2a_bench_skin$sk.err
python - <<PY
import json
f="gpurun_out/r2a_bench_skin$sk.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["passes_ms"], d["config"]["neighbour_search"])
except Exception as e:
    print(f, "FAILED", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 8 --warmup 3 --e2e-steps 1 --cpu-steps 1 > gpurun_out/r2a_ncu_bench.log 2>&1
tail -2 gpurun_out/r2a_tests.log
