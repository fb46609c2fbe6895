#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -q -p no:cacheprovider > gpurun_out/r2n_nccl_tests.log 2>&1
tail -3 gpurun_out/r2n_nccl_tests.log | cut -c1-200
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 40 --warmup 3 --e2e-steps 2 > gpurun_out/r2n_bench$N.json 2> gpurun_out/r2n_bench$N.err
python - <<PY
import json
f="gpurun_out/r2n_bench$N.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["roofline"]["passes_ms"], d["config"].get("slab",{}).get("exchange_ms_bytes_by_phase"), d["device_error_word"], round(d["e2e"]["value"]/1e6,1), d["config"].get("plan"))
except Exception as e:
    print(f, "FAILED", e)
PY
tail -2 gpurun_out/r2n_bench$N.err | cut -c1-300
done
