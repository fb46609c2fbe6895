"""One worker of the CPU arm of bench.py (test / measurement infrastructure, like the rest of
oracle/: never on the product path).

    python -m oracle.cpu_arm --workload tgv3d --nx 32 --warmup 1 --steps 3

builds the bounded sample of the workload with the oracle's case generator, runs `warmup`
untimed and `steps` timed advance() calls of the NumPy restatement of the reference's
edge-list algorithm (jax_sph/integrator.py:22-56, solver.py:702-951, partition.py) on ONE
core, and prints one JSON line {n, steps, t0, t1} (wall-clock epoch seconds of the timed
region).  bench.py launches one worker per host core -- independent periodic boxes, the way
its GPU arm runs independent engines -- and reports the aggregate over the common window.
"""

import argparse
import json
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_setup(workload, nx):
    from oracle import cases

    if workload == "tgv3d":
        return cases.make_case("tgv", dim=3, dx=2 * np.pi / nx, dtype=np.float32, tvf=1.0,
                               viscosity=0.02)
    if workload == "ht3d":
        return cases.make_case("ht", dim=3, dx=1.0 / nx, dtype=np.float32)
    if workload == "tgv2d":
        return cases.make_case("tgv", dim=2, dx=1.0 / nx, dtype=np.float32, tvf=1.0)
    raise ValueError(workload)


def main():
    from oracle import integrator

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="tgv3d")
    ap.add_argument("--nx", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    setup = make_setup(a.workload, a.nx)
    if a.warmup > 0:
        integrator.simulate(setup, a.warmup, fast_segment_sum=True)
    t0 = time.time()
    integrator.simulate(setup, a.steps, fast_segment_sum=True)
    t1 = time.time()
    print(json.dumps({"n": len(setup.state["r"]), "steps": a.steps, "t0": t0, "t1": t1}))


if __name__ == "__main__":
    main()
