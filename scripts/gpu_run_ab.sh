#!/bin/bash
# A/B of two builds of the library on the headline bench: libsphb200.so (new) against libsphb200_prev.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 900 python -m pytest $TESTS -m gpu -x -q > gpurun_out/r3ab_tests.log 2>&1; tail -4 gpurun_out/r3ab_tests.log; fi
bench() {
  timeout 400 python bench.py --steps 40 --warmup 3 --e2e-steps 1 --cpu-steps 1 --no-configs > gpurun_out/r3ab_$1.json 2> gpurun_out/r3ab_$1.err
  python - $1 <<'PY'
import json,sys
f="gpurun_out/r3ab_%s.json"%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["passes_ms"].items()}, "err", d["device_error_word"])
except Exception as e:
    print(sys.argv, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
bench new
cp jax_sph_b200/libsphb200.so /tmp/new.so; cp jax_sph_b200/libsphb200_prev.so jax_sph_b200/libsphb200.so; touch jax_sph_b200/libsphb200.so
bench prev
cp /tmp/new.so jax_sph_b200/libsphb200.so
bench new2
