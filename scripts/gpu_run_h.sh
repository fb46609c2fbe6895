#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab_nccl.py tests/test_gpu_slab.py -m gpu -q -p no:cacheprovider > gpurun_out/r2h_nccl_tests.log 2>&1
tail -8 gpurun_out/r2h_nccl_tests.log | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 2 > gpurun_out/r2h_bench2.json 2> gpurun_out/r2h_bench2.err
python - <<'PY'
import json
f="gpurun_out/r2h_bench2.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["roofline"]["passes_ms"], d["config"].get("slab",{}).get("exchange_ms_bytes_by_phase"), d["device_error_word"], d["e2e"]["value"])
except Exception as e:
    print(f, "FAILED", e)
PY
tail -5 gpurun_out/r2h_bench2.err
