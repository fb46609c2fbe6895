"""GPU parity: CUDA engine (through the C ABI) vs the oracle / golden vectors.

Covers WCSPH.forward (solver.py:705-949) and si_euler.advance
(integrator.py:22-56) for every solver variant the benchmark configs use:
SPH summation (C1), +tvf (C2a, C4), RIE + density evolution (C2b), dam break
(bc_trick + density evolution + artificial viscosity + g_ext, C3), heated
channel (bc_trick + heat + band g_ext + Dirichlet wall, C5), Couette with
Wendland C2 and a moving wall.  Tolerances: tests/_util.py.
"""

import numpy as np
import pytest

from _util import assert_close, drift_ok, load_golden

pytestmark = pytest.mark.gpu

GOLDEN_CASES = {
    "tgv2d_sph": dict(case="tgv", dim=2, dx=0.02),
    "tgv2d_tvf": dict(case="tgv", dim=2, dx=0.02, tvf=1.0),
    "tgv2d_rie": dict(case="tgv", dim=2, dx=0.02, solver="RIE", density_evolution=True),
    "tgv3d_tvf": dict(case="tgv", dim=3, dx=2 * np.pi / 16, tvf=1.0, viscosity=0.02),
    "db2d": dict(case="db", dim=2, dx=0.04),
    "ht2d": dict(case="ht", dim=2, dx=0.02),
    "cf2d_wc2k": dict(case="cf", dim=2, dx=0.04, kernel="WC2K", h_factor=1.3),
}
FWD_KEYS = ("rho", "p", "u", "v", "dudt", "dvdt", "drhodt", "T", "dTdt")
ADV_KEYS = ("r", "u", "v", "rho", "p", "T", "dudt", "dvdt")
# default: per-step neighbour lists built by the density sweep and consumed by the later sweeps;
# no_lists: every sweep searches on its own; tiny_lists: rows of 16 overflow, so every tile falls
# back to its own search (the list is an accelerator, never a correctness dependency)
PLANS = {"default": {}, "coarse_cells": dict(cell_sub=[1, 1, 1], threads=128, list_cap=96),
         "no_lists": dict(nl_cap=-1), "tiny_lists": dict(nl_cap=16)}


def _setup(name):
    from oracle import cases

    return cases.make_case(dtype=np.float32, **GOLDEN_CASES[name])


def _engine(setup, **tuning):
    from jax_sph_b200 import Engine, config_from_setup

    return Engine(config_from_setup(setup, **tuning), len(setup.state["r"]))


@pytest.mark.parametrize("plan", list(PLANS))
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_forward_matches_golden(name, plan):
    z, state0 = load_golden(name)
    setup = _setup(name)
    for k in state0:  # the committed vectors must describe this very setup
        assert np.array_equal(state0[k], setup.state[k]), f"golden state0[{k}] is stale"
    eng = _engine(setup, **PLANS[plan])
    eng.upload(state0)
    eng.step(0.0, 1, integrate=False, bc=False)
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in FWD_KEYS:
        g = got[k].numpy()
        assert np.isfinite(g).all(), k
        assert_close(k, g, z[f"forward_f32_{k}"], setup, what=f"{name} forward")
        drift_ok(k, g, z[f"forward_f32_{k}"], z[f"forward_f64_{k}"], setup)
    # untouched fields come back bit-identical
    for k in ("r", "mass", "eta", "tag"):
        assert np.array_equal(got[k].numpy(), state0[k]), k


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_advance_20_steps_matches_golden(name):
    z, state0 = load_golden(name)
    setup = _setup(name)
    nsteps = int(z["nsteps"])
    eng = _engine(setup)
    eng.upload(state0)
    eng.step(setup.dt, nsteps)
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in ADV_KEYS:
        g = got[k].numpy()
        # 20 steps let the per-step float32 noise accumulate: allow sqrt(20) ~ 5 units
        assert_close(k, g, z[f"advance_f32_{k}"], setup, factor=5.0, what=f"{name} advance")
        drift_ok(k, g, z[f"advance_f32_{k}"], z[f"advance_f64_{k}"], setup, factor=3.0)


@pytest.mark.parametrize("name,kw", [
    ("tgv2d_tvf", dict(case="tgv", dim=2, dx=0.02, tvf=1.0)),
    ("tgv2d_rie", dict(case="tgv", dim=2, dx=0.02, solver="RIE", density_evolution=True)),
])
def test_200_step_trajectory_vs_oracle(name, kw):
    """north_star: tolerance-matched 200-step trajectories (2D TGV, N = 2500)."""
    from oracle import cases, integrator

    setup = cases.make_case(dtype=np.float32, **kw)
    ref = integrator.simulate(setup, 200, fast_segment_sum=True)
    setup64 = cases.make_case(dtype=np.float64, **kw)
    for k, v in setup.state.items():
        setup64.state[k] = v.astype(np.float64) if v.dtype == np.float32 else v.copy()
    ref64 = integrator.simulate(setup64, 200, fast_segment_sum=True)
    eng = _engine(setup)
    eng.upload(setup.state)
    eng.step(setup.dt, 200)
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in ("r", "u", "rho"):
        g = got[k].numpy()
        assert_close(k, g, ref[k], setup, factor=15.0, what=f"{name} 200 steps")
        drift_ok(k, g, ref[k], ref64[k], setup, factor=3.0)
    # energy decay (validation/validate.py:181-195): E_kin of the engine state vs oracle state
    ek = 0.5 * float((got["mass"].numpy()[:, None] * got["u"].numpy() ** 2).sum())
    ek_ref = 0.5 * float((ref["mass"][:, None] * ref["u"] ** 2).sum())
    assert abs(ek - ek_ref) <= 1e-5 * ek_ref
    ek_dev, umax = eng.stats()
    assert abs(ek_dev - ek) <= 1e-5 * ek and abs(umax - np.linalg.norm(got["u"].numpy(), axis=1).max()) < 1e-5


@pytest.mark.parametrize("kw", [
    dict(case="pf", dim=2, dx=0.05),                      # tiny box: exact_all path, band g_ext
    dict(case="pf", dim=2, dx=0.05, solver="RIE", density_evolution=True, is_bc_trick=True),
    dict(case="ht", dim=3, dx=0.04),                      # 3D walls + heat
    dict(case="db", dim=2, dx=0.05, density_renormalize=True),
    dict(case="cf", dim=2, dx=0.05, free_slip=True),
    dict(case="tgv", dim=3, dx=2 * np.pi / 12, kernel="WC2K", h_factor=1.3, tvf=1.0),
    dict(case="tgv", dim=3, dx=2 * np.pi / 12, solver="DELTA", density_evolution=True),  # 3x3 L
    dict(case="db", dim=2, dx=0.05, solver="DELTA", gamma=7.0, artificial_alpha=0.0,
         density_renormalize=True),
])
def test_variants_vs_oracle(kw):
    """Variants without a committed golden file: oracle evaluated on the fly."""
    from oracle import cases, integrator
    from oracle.solver import WCSPH

    setup = cases.make_case(dtype=np.float32, **kw)
    solver = WCSPH(setup.displacement_fn, setup.eos, setup.g_ext_fn, setup.dx, setup.dim, setup.dt,
                   setup.c_ref, setup.eta_limiter, setup.diff_delta, setup.diff_alpha, setup.solver, setup.kernel,
                   setup.h_factor, setup.is_bc_trick, setup.density_evolution,
                   setup.artificial_alpha, setup.free_slip, setup.density_renormalize,
                   setup.heat_conduction, dtype=np.float32)
    nfn = integrator.make_neighbors_fn(setup.box_size, solver._kernel_fn.cutoff)
    ref = solver.forward({k: v.copy() for k, v in setup.state.items()}, nfn(setup.state["r"]))
    eng = _engine(setup)
    eng.upload(setup.state)
    eng.step(0.0, 1, integrate=False, bc=False)
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in FWD_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, what=f"{kw} forward")
    ref = integrator.simulate(setup, 5)
    eng.upload(setup.state)
    eng.step(setup.dt, 5)
    got = eng.download(host=True)
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, factor=3.0, what=f"{kw} advance")
