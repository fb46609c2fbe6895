#!/bin/bash
# Round-end evidence run (GPU box, 1 GPU): GPU test-suite, parity table, bench lines (60 and 200
# timed steps), ncu launch list and full-set capture of the duo kernels at the benchmark size
# -> gpurun_out/, copied to profiles/ by hand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_gpu_tests.txt 2>&1; tail -3 gpurun_out/r02_gpu_tests.txt
timeout 900 python scripts/parity_table.py > gpurun_out/r02_parity_table.txt 2> gpurun_out/r02_parity_table.err
tail -2 gpurun_out/r02_parity_table.txt; tail -3 gpurun_out/r02_parity_table.err
timeout 900 python bench.py --steps 60 --warmup 3 > gpurun_out/r02_bench_tgv3d_256.json 2> gpurun_out/r02_bench.err
tail -3 gpurun_out/r02_bench.err
timeout 600 python bench.py --steps 200 --warmup 3 --no-configs > gpurun_out/r02_bench_tgv3d_256_200steps.json 2> gpurun_out/r02_bench200.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_tgv3d_256.json", "gpurun_out/r02_bench_tgv3d_256_200steps.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["roofline"]["passes_ms"], d["config"]["neighbour_search"]["searches"], "e2e", round(d["e2e"]["value"]/1e6,1), d["clocks"])
        s = d["stateless_advance"]; print("stateless", s["ms_per_step"], s["engine_order"]["ms_per_step"], s["scratch_workspace"]["ms_per_step"])
        for k,v in d.get("configs", {}).items(): print(k, v.get("n"), round(v.get("value",0)/1e6,1), v.get("ms_per_step"), v.get("searches"), v.get("steps_total"), v.get("device_error_word"), v.get("error"))
        r=d["roofline"]; print("roofline", r["kernel"][:40], r["bound"], r["achieved"], r["peak"], r["frac"], r["traffic"], r["step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_tgv3d_256.csv python bench.py --steps 6 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo -s 0 -c 4 -f -o gpurun_out/r02_duo_tgv3d_256 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r02_ncu.log 2>&1
tail -2 gpurun_out/r02_ncu.log | cut -c1-200
