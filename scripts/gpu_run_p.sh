#!/bin/bash
# N-GPU check of the slab transports: parity test (both transports) + bench per transport
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out

timeout 900 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -x -q -k "${N}-" > gpurun_out/r3p_tests_$N.log 2>&1; tail -5 gpurun_out/r3p_tests_$N.log
for tr in direct nccl; do
  SPHB200_SLAB_TRANSPORT=$tr timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 3 --e2e-steps 1 --cpu-steps 1 > gpurun_out/r3p_bench${N}_$tr.json 2> gpurun_out/r3p_bench${N}_$tr.err
  python - $N $tr <<'PY'
import json,sys
f="gpurun_out/r3p_bench%s_%s.json"%(sys.argv[1],sys.argv[2])
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[2], "N", d["n_gpus"], round(d["ms_per_step"],3), "ms", round(d["value"]/1e9,3), "G/s", d["config"]["slab"]["exchange_ms_bytes_by_phase"], d["config"].get("transport"), "enqueue", round(d["config"]["slab"]["host_enqueue_ms_per_step"],3), d.get("device_error_word"), d.get("particles_total_after_run"))
except Exception as e:
    print(sys.argv, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done
