"""Measured parity errors of the CUDA engine, per case and field (run on the GPU box).

    python scripts/parity_table.py > profiles/r02_parity_table.txt

For every reference golden (tests/golden/ref_*.npz: output of the UNMODIFIED reference sources,
tests/golden/make_reference_golden.py) and every 200-step golden (ref200_*.npz) the engine runs
from the reference's state0 through the C ABI, and each field is reported as

    err   max |engine - ref32|                      (ref32: the reference's float32 run)
    tol   the bound the tests assert (tests/_util.py: 1e-5 max|ref| + float32 EoS noise floor,
          times the step factor of the test)
    err/tol
    |e-64|   max |engine - ref64|   and   |32-64|   max |ref32 - ref64|   (drift bound: the engine
          is asserted to be no further from the float64 run than 2-3 x the float32 run is)

so that one can see where the engine sits inside its tolerances.  The oracle-based 3D case of
tests/test_gpu_parity3d.py (interior tiles, frozen steps) is appended.
"""

import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _util import GOLDEN, max_err, periodic_err, tolerance  # noqa: E402

FWD_KEYS = ("rho", "p", "u", "v", "dudt", "dvdt", "drhodt", "T", "dTdt")
ADV_KEYS = ("r", "u", "v", "rho", "p", "T", "dudt", "dvdt")


def engine_for(setup, **tuning):
    from jax_sph_b200 import Engine, config_from_setup

    return Engine(config_from_setup(setup, **tuning), len(setup.state["r"]))


def row(case, what, key, got, ref32, ref64, setup, factor, same_start=True):
    err = periodic_err(got, ref32, setup.box_size) if key == "r" else max_err(got, ref32)
    tol = factor * tolerance(key, ref32, setup)
    line = f"{case:16s} {what:9s} {key:7s} err {err:10.3e}  tol {tol:10.3e}  err/tol {err / tol:6.3f}  "
    if same_start:
        e64 = periodic_err(got, ref64, setup.box_size) if key == "r" else max_err(got, ref64)
        d3264 = periodic_err(ref32, ref64, setup.box_size) if key == "r" else max_err(ref32, ref64)
        line += f"|e-64| {e64:10.3e}  |32-64| {d3264:10.3e}"
    else:
        # the reference draws its position noise per dtype: its float64 run starts elsewhere
        line += "|e-64|        n/a  |32-64|        n/a  (float64 run: another noise sample)"
    print(line)
    return err / tol


def same_start(z):
    return float(np.abs(z["state0_f32_r"].astype(np.float64) - z["state0_f64_r"]).max()) < 1e-5


def main():
    from oracle import cases

    worst = 0.0
    print("# engine vs the reference's own output (tests/golden/ref_*.npz), float32; tolerances of "
          "tests/test_gpu_reference.py")
    for path in sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz"))):
        name = os.path.basename(path)[4:-4]
        z = np.load(path)
        kw = json.loads(str(z["make_case_kwargs"]))
        meta = json.loads(str(z["meta_f32"]))
        setup = cases.make_case(dtype=np.float32, **kw)
        state0 = {k: np.array(z[f"state0_f32_{k}"]) for k in setup.state}
        setup.state = state0
        eng = engine_for(setup)
        eng.upload(state0)
        eng.step(0.0, 1)
        got = eng.download(host=True)
        for k in FWD_KEYS:
            worst = max(worst, row(name, "forward", k, got[k].numpy(), z[f"forward_f32_{k}"],
                                   z[f"forward_f64_{k}"], setup, 1.0, same_start(z)))
        eng.upload(state0)
        eng.step(meta["dt"], meta["nsteps"])
        got = eng.download(host=True)
        cnt = eng.counters()
        for k in ADV_KEYS:
            worst = max(worst, row(name, f"{meta['nsteps']} steps", k, got[k].numpy(),
                                   z[f"advance_f32_{k}"], z[f"advance_f64_{k}"], setup, 5.0, same_start(z)))
        print(f"{name:16s} searches {cnt['searches']} of {cnt['steps']} steps, device error {eng.error()}")
        eng.close()
    print("# 200-step trajectories (tests/golden/ref200_*.npz), tolerance factor 15")
    for path in sorted(glob.glob(os.path.join(GOLDEN, "ref200_*.npz"))):
        name = os.path.basename(path)[7:-4]
        z = np.load(path)
        kw = json.loads(str(z["make_case_kwargs"]))
        meta = json.loads(str(z["meta_f32"]))
        setup = cases.make_case(dtype=np.float32, **kw)
        setup.state = {k: np.array(z[f"state0_f32_{k}"]) for k in setup.state}
        eng = engine_for(setup)
        eng.upload(setup.state)
        eng.step(meta["dt"], meta["nsteps"])
        got = eng.download(host=True)
        cnt = eng.counters()
        for k in ("r", "u", "v", "rho", "T"):
            worst = max(worst, row(name, "200 steps", k, got[k].numpy(), z[f"advance_f32_{k}"],
                                   z[f"advance_f64_{k}"], setup, 15.0, same_start(z)))
        print(f"{name:16s} searches {cnt['searches']} of {cnt['steps']} steps, device error {eng.error()}")
        eng.close()
    print("# oracle-based 3D case with interior tiles and frozen steps (tests/test_gpu_parity3d.py)")
    from oracle import integrator
    from test_gpu_parity3d import CASES_3D, _oracle_forward, interior_tiles

    kw, nsteps = CASES_3D["tgv3d_tvf_40"]
    setup = cases.make_case(dtype=np.float32, **kw)
    setup64 = cases.make_case(dtype=np.float64, **kw)
    for k, v in setup.state.items():
        setup64.state[k] = v.astype(np.float64) if v.dtype == np.float32 else v.copy()
    eng = engine_for(setup)
    inner, total = interior_tiles(eng.plan())
    ref, ref64 = _oracle_forward(setup, np.float32), _oracle_forward(setup64, np.float64)
    eng.upload(setup.state)
    eng.step(0.0, 1, integrate=False, bc=False)
    got = eng.download(host=True)
    for k in FWD_KEYS:
        worst = max(worst, row("tgv3d_tvf_40", "forward", k, got[k].numpy(), ref[k], ref64[k], setup, 1.0))
    ref = integrator.simulate(setup, nsteps, fast_segment_sum=True)
    ref64 = integrator.simulate(setup64, nsteps, fast_segment_sum=True)
    eng.upload(setup.state)
    eng.step(setup.dt, nsteps)
    got = eng.download(host=True)
    cnt = eng.counters()
    for k in ADV_KEYS:
        worst = max(worst, row("tgv3d_tvf_40", f"{nsteps} steps", k, got[k].numpy(), ref[k], ref64[k],
                               setup, 4.0))
    print(f"tgv3d_tvf_40     {inner} of {total} tiles interior, searches {cnt['searches']} of "
          f"{cnt['steps']} steps, device error {eng.error()}")
    print(f"# worst err/tol over all rows: {worst:.3f}")


if __name__ == "__main__":
    main()
