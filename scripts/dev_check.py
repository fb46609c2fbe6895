"""Development check (GPU box): compare the CUDA path against the oracle on a set
of small cases and print error metrics.  Not a test; tests/ holds the asserted
versions.  Usage: python scripts/dev_check.py [case ...]"""

import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jax_sph_b200 import Engine, config_from_setup  # noqa: E402
from oracle import cases, integrator, partition  # noqa: E402
from oracle.solver import WCSPH  # noqa: E402


def relerr(a, b, extra_scale=0.0):
    scale = float(np.abs(b).max()) + extra_scale
    return float(np.abs(a - b).max()) / max(scale, 1e-30)


def check_neighbors(name, r, box, cutoff_h, dim, kernel="QSK", mask_self=False, **tune):
    from jax_sph_b200 import make_config

    h = cutoff_h
    cfg = make_config(dim, box, h, 0.0, kernel=kernel, **tune)
    n = len(r)
    eng = Engine(cfg, n)
    eng.upload({"r": r.astype(np.float32)})
    cutoff = (3.0 if kernel == "QSK" else 2.0) * h
    ref = partition.neighbor_pairs(r.astype(np.float32), np.asarray(box, dtype=np.float64), cutoff,
                                   mask_self=mask_self)
    cap = int(ref.shape[1] * 1.25) + 8
    idx, cnt = eng.neighbor_list(cap, mask_self=mask_self)
    err = eng.error()
    got = idx.cpu().numpy()
    got = got[:, got[0] < n]
    same = got.shape == ref.shape and bool((got == ref).all())
    tb = partition.tie_band(r.astype(np.float32), np.asarray(box, dtype=np.float64), cutoff, ref)
    print(f"[nl] {name}: N={n} edges ref={ref.shape[1]} got={cnt} err={err} exact={same} "
          f"tie_band={tb} plan={eng.plan()}")
    if not same:
        a = set(map(tuple, got.T.tolist()))
        b = set(map(tuple, ref.T.tolist()))
        print("   missing", len(b - a), "extra", len(a - b), list(b - a)[:5], list(a - b)[:5])
    return same


def check_case(name, setup, nsteps=3, **tune):
    n = len(setup.state["r"])
    cfg = config_from_setup(setup, **tune)
    eng = Engine(cfg, n)
    print(f"[case] {name}: N={n} plan={eng.plan()}")
    # forward only
    solver = WCSPH(setup.displacement_fn, setup.eos, setup.g_ext_fn, setup.dx, setup.dim, setup.dt,
                   setup.c_ref, setup.eta_limiter, 0.0, 0.0, setup.solver, setup.kernel,
                   setup.h_factor, setup.is_bc_trick, setup.density_evolution,
                   setup.artificial_alpha, setup.free_slip, setup.density_renormalize,
                   setup.heat_conduction, dtype=np.float32, fast_segment_sum=False)
    nfn = integrator.make_neighbors_fn(setup.box_size, solver._kernel_fn.cutoff)
    idx = nfn(setup.state["r"])
    ref = solver.forward({k: v.copy() for k, v in setup.state.items()}, idx)
    eng.upload(setup.state)
    eng.step(0.0, 1, integrate=False, bc=False)
    got = eng.download(host=True)
    err = eng.error()
    line = []
    for k in ("rho", "p", "u", "v", "dudt", "dvdt", "drhodt", "T", "dTdt"):
        if k in got:
            line.append(f"{k}={relerr(got[k].numpy(), ref[k], setup.p_ref if k == 'p' else 0):.2e}")
    print(f"   forward err={err}: " + " ".join(line))
    # advance nsteps
    t0 = time.time()
    ref = integrator.simulate(setup, nsteps, fast_segment_sum=False)
    t1 = time.time()
    eng.upload(setup.state)
    eng.step(setup.dt, nsteps)
    got = eng.download(host=True)
    err = eng.error()
    line = []
    for k in ("r", "u", "v", "rho", "p", "dudt", "dvdt", "T", "dTdt"):
        if k in got:
            line.append(f"{k}={relerr(got[k].numpy(), ref[k], setup.p_ref if k == 'p' else 0):.2e}")
    print(f"   advance x{nsteps} err={err} (oracle {t1 - t0:.1f}s): " + " ".join(line))
    ek, um = eng.stats()
    print(f"   ekin={ek:.6e} umax={um:.6e} launches={eng.launches()}")


def main():
    want = set(sys.argv[1:])

    def run(name, fn):
        if want and name not in want:
            return
        try:
            fn()
        except Exception:
            print(f"[FAIL] {name}")
            traceback.print_exc()
        torch.cuda.synchronize()

    rng = np.random.default_rng(0)
    # KATs (tiny boxes -> exact_all mode)
    r1 = np.array([[0.1, 0.1], [0.1, 0.3], [0.1, 0.9], [0.6, 0.5]])
    run("kat1", lambda: check_neighbors("kat1", r1, [1.0, 1.0], 0.11, 2))
    run("kat1m", lambda: check_neighbors("kat1m", r1, [1.0, 1.0], 0.11, 2, mask_self=True))
    r2 = rng.random((2000, 2))
    run("rand2d", lambda: check_neighbors("rand2d", r2, [1.0, 1.0], 0.02, 2))
    run("rand2d_s2", lambda: check_neighbors("rand2d_s2", r2, [1.0, 1.0], 0.02, 2, cell_sub=[2, 2, 0]))
    r3 = rng.random((4000, 3)) * np.array([1.0, 0.7, 0.5])
    run("rand3d", lambda: check_neighbors("rand3d", r3, [1.0, 0.7, 0.5], 0.03, 3))
    run("rand3d_s2", lambda: check_neighbors("rand3d_s2", r3, [1.0, 0.7, 0.5], 0.03, 3,
                                             cell_sub=[2, 2, 2]))
    run("rand3d_wc2", lambda: check_neighbors("rand3d_wc2", r3, [1.0, 0.7, 0.5], 0.04, 3, kernel="WC2K"))
    lat = cases.make_case("tgv", dim=3, dx=2 * np.pi / 24, dtype=np.float32)
    run("lat3d", lambda: check_neighbors("lat3d", lat.state["r"], lat.box_size, lat.dx, 3))
    lat2 = cases.make_case("tgv", dim=2, dx=0.02, dtype=np.float32)
    run("lat2d", lambda: check_neighbors("lat2d", lat2.state["r"], lat2.box_size, lat2.dx, 2))

    run("tgv2d", lambda: check_case("tgv2d", cases.make_case("tgv", dim=2, dx=0.02, dtype=np.float32)))
    run("tgv2d_tvf", lambda: check_case("tgv2d_tvf", cases.make_case("tgv", dim=2, dx=0.02, dtype=np.float32, tvf=1.0)))
    run("tgv2d_rie", lambda: check_case("tgv2d_rie", cases.make_case(
        "tgv", dim=2, dx=0.02, dtype=np.float32, solver="RIE", density_evolution=True)))
    run("tgv3d", lambda: check_case("tgv3d", cases.make_case(
        "tgv", dim=3, dx=2 * np.pi / 20, dtype=np.float32, tvf=1.0, viscosity=0.02)))
    run("tgv3d_s2", lambda: check_case("tgv3d_s2", cases.make_case(
        "tgv", dim=3, dx=2 * np.pi / 20, dtype=np.float32, tvf=1.0, viscosity=0.02),
        cell_sub=[2, 2, 2]))
    run("db", lambda: check_case("db", cases.make_case("db", dim=2, dx=0.04, dtype=np.float32)))
    run("pf", lambda: check_case("pf", cases.make_case("pf", dim=2, dx=0.05, dtype=np.float32)))
    run("cf", lambda: check_case("cf", cases.make_case("cf", dim=2, dx=0.05, dtype=np.float32)))
    run("ht", lambda: check_case("ht", cases.make_case("ht", dim=2, dx=0.02, dtype=np.float32)))
    run("ht3d", lambda: check_case("ht3d", cases.make_case("ht", dim=3, dx=0.04, dtype=np.float32)))
    run("pf_rie", lambda: check_case("pf_rie", cases.make_case(
        "pf", dim=2, dx=0.05, dtype=np.float32, solver="RIE", density_evolution=True)))
    run("tgv2d_wc2", lambda: check_case("tgv2d_wc2", cases.make_case(
        "tgv", dim=2, dx=0.02, dtype=np.float32, kernel="WC2K", h_factor=1.3)))


if __name__ == "__main__":
    main()
