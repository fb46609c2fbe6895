"""Generate the golden vectors in tests/golden/*.npz.

    python tests/golden/make_golden.py

The reference itself (tumaer/jax-sph) cannot be executed in the build image
(jax / jaxlib are not installable), so the vectors come
from the NumPy oracle (oracle/), which is pinned against the reference's own
known-answer tests by tests/test_oracle_pins.py.  Each file holds, for one
case: the initial state, WCSPH.forward of it in float32 and float64, and the
state after NSTEPS advance() calls in float32 and float64 -- all in the
reference's state-dict layout.  Neighbour-list known answers of the reference's
tests/test_neighbors.py:89-121 are stored verbatim in neighbors_kat.npz.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import cases, integrator  # noqa: E402
from oracle.solver import WCSPH  # noqa: E402

NSTEPS = 20
OUT_KEYS = ("r", "u", "v", "rho", "p", "dudt", "dvdt", "drhodt", "T", "dTdt")

CASES = {
    # name: (make_case kwargs)
    "tgv2d_sph": dict(case="tgv", dim=2, dx=0.02),
    "tgv2d_tvf": dict(case="tgv", dim=2, dx=0.02, tvf=1.0),
    "tgv2d_rie": dict(case="tgv", dim=2, dx=0.02, solver="RIE", density_evolution=True),
    "tgv3d_tvf": dict(case="tgv", dim=3, dx=2 * np.pi / 16, tvf=1.0, viscosity=0.02),
    "db2d": dict(case="db", dim=2, dx=0.04),
    "ht2d": dict(case="ht", dim=2, dx=0.02),
    "cf2d_wc2k": dict(case="cf", dim=2, dx=0.04, kernel="WC2K", h_factor=1.3),
}


def solver_of(setup, dtype):
    return WCSPH(setup.displacement_fn, setup.eos, setup.g_ext_fn, setup.dx, setup.dim, setup.dt,
                 setup.c_ref, setup.eta_limiter, setup.diff_delta, setup.diff_alpha, setup.solver, setup.kernel,
                 setup.h_factor, setup.is_bc_trick, setup.density_evolution,
                 setup.artificial_alpha, setup.free_slip, setup.density_renormalize,
                 setup.heat_conduction, dtype=dtype)


def main():
    for name, kw in CASES.items():
        out = {}
        for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
            setup = cases.make_case(dtype=dtype, **kw)
            if dtype == np.float64:
                # same initial particles as the float32 run (cast up), so that the
                # float64 result is the "exact" answer for the float32 input
                s32 = cases.make_case(dtype=np.float32, **kw)
                for k, v in s32.state.items():
                    setup.state[k] = v.astype(np.float64) if v.dtype == np.float32 else v.copy()
            else:
                for k, v in setup.state.items():
                    out["state0_" + k] = v
            solver = solver_of(setup, dtype)
            nfn = integrator.make_neighbors_fn(setup.box_size, solver._kernel_fn.cutoff)
            fwd = solver.forward({k: v.copy() for k, v in setup.state.items()},
                                 nfn(setup.state["r"]))
            adv = integrator.simulate(setup, NSTEPS)
            for k in OUT_KEYS:
                out[f"forward_{tag}_{k}"] = fwd[k]
                out[f"advance_{tag}_{k}"] = adv[k]
        out["nsteps"] = np.int32(NSTEPS)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, len(out["state0_r"]), "particles ->", os.path.getsize(path) // 1024, "KiB")

    # reference tests/test_neighbors.py:89-121, verbatim targets
    r1 = np.array([[0.1, 0.1], [0.1, 0.3], [0.1, 0.9], [0.6, 0.5]])
    r2 = np.array([[0.5, 0.2], [0.2, 0.5], [0.5, 0.5], [0.8, 0.5], [0.5, 0.8]])
    np.savez(
        os.path.join(HERE, "neighbors_kat.npz"), cutoff=0.33, box=np.array([1.0, 1.0]),
        r1=r1, r2=r2,
        t1_self=np.array([[0, 1, 2, 0, 1, 0, 2, 3], [0, 0, 0, 1, 1, 2, 2, 3]]),
        t1_mask=np.array([[1, 2, 0, 0], [0, 0, 1, 2]]),
        t2_self=np.array([[0, 2, 1, 2, 0, 1, 2, 3, 4, 2, 3, 2, 4],
                          [0, 0, 1, 1, 2, 2, 2, 2, 2, 3, 3, 4, 4]]),
        t2_mask=np.array([[2, 2, 0, 1, 3, 4, 2, 2], [0, 1, 2, 2, 2, 2, 3, 4]]))


if __name__ == "__main__":
    main()
