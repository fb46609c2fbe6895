"""GPU parity at sizes where the engine's fast paths actually run (3D, >= 32 cells per axis).

The small goldens (<= 20^3 particles) have at most 13 cells per axis: every tile's stencil wraps
around the periodic box, so the INTERIOR code of sweep.cuh (no periodic fold, single-segment row
walk) and the frozen-sort steps between two searches were never compared with the oracle on
physics there.  Here the oracle (oracle/, the CPU restatement of solver.py:705-949 and
integrator.py:22-56) runs beside the engine on noisy 3D lattices large enough that

* interior tiles exist (asserted from the engine's own plan: 16 of 108 at 40^3, 48 of 168 for the
  channel),
* the step sequence contains searches AND frozen steps (asserted from the device counters),

plus the degraded modes at 3D scale: a stencil that needs several staging groups, list rows that
overflow, skin off / tiny / large.  Tolerances: tests/_util.py.
"""

import numpy as np
import pytest

from _util import assert_close, drift_ok

pytestmark = pytest.mark.gpu

FWD_KEYS = ("rho", "p", "u", "v", "dudt", "dvdt", "drhodt", "T", "dTdt")
ADV_KEYS = ("r", "u", "v", "rho", "p", "T", "dudt", "dvdt")


def _engine(setup, **tuning):
    from jax_sph_b200 import Engine, config_from_setup

    return Engine(config_from_setup(setup, **tuning), len(setup.state["r"]))


def interior_tiles(plan):
    """Tiles whose stencil holds no periodic image, from the engine's plan (the rule of
    sweep.cuh: stencil [t T - S, t T + T + S) inside [0, n) and n >= 2 S + 2 on every axis)."""
    total, inner = 1, 1
    for n, S, T in zip(plan["cells"], plan["sub"], plan["tile"]):
        if n <= 1:
            continue
        nt = (n + T - 1) // T
        ok = 0
        for t in range(nt):
            lo, hi = t * T - S, t * T + min(T, n - t * T) + S
            ok += int(lo >= 0 and hi <= n and n >= 2 * S + 2)
        total *= nt
        inner *= ok
    return inner, total


def _oracle_forward(setup, dtype):
    from oracle import cases, integrator
    from oracle.solver import WCSPH

    solver = WCSPH(setup.displacement_fn, setup.eos, setup.g_ext_fn, setup.dx, setup.dim, setup.dt,
                   setup.c_ref, setup.eta_limiter, setup.diff_delta, setup.diff_alpha, setup.solver,
                   setup.kernel, setup.h_factor, setup.is_bc_trick, setup.density_evolution,
                   setup.artificial_alpha, setup.free_slip, setup.density_renormalize,
                   setup.heat_conduction, dtype=dtype)
    nfn = integrator.make_neighbors_fn(setup.box_size, solver._kernel_fn.cutoff)
    return solver.forward({k: v.copy() for k, v in setup.state.items()}, nfn(setup.state["r"]))


CASES_3D = {
    # BASELINE configs[3] (validation/tgv3d.sh) at 40^3 with the lattice noise of case_setup.py:139-144
    # (23 cells per axis: the 9 x 4 x 4 tiles of the uniform-viscosity duo sweeps leave 16 interior)
    "tgv3d_tvf_40": (dict(case="tgv", dim=3, dx=2 * np.pi / 40, tvf=1.0, viscosity=0.02,
                          r0_noise_factor=0.25), 8),
    # BASELINE configs[4] (cases/ht.yaml, case.dim=3): walls + heat + band force, 89 600 particles
    "ht3d_80": (dict(case="ht", dim=3, dx=0.0125), 4),
}


@pytest.mark.parametrize("name", list(CASES_3D))
def test_interior_fast_path_vs_oracle(name):
    from oracle import cases, integrator

    kw, nsteps = CASES_3D[name]
    setup = cases.make_case(dtype=np.float32, **kw)
    eng = _engine(setup)
    inner, total = interior_tiles(eng.plan())
    assert inner >= 4, f"{name}: only {inner} of {total} tiles interior"
    # forward: engine vs float32 oracle, drift bound against the float64 oracle
    ref = _oracle_forward(setup, np.float32)
    setup64 = cases.make_case(dtype=np.float64, **kw)
    for k, v in setup.state.items():
        setup64.state[k] = v.astype(np.float64) if v.dtype == np.float32 else v.copy()
    ref64 = _oracle_forward(setup64, np.float64)
    eng.upload(setup.state)
    eng.step(0.0, 1, integrate=False, bc=False)
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in FWD_KEYS:
        g = got[k].numpy()
        assert np.isfinite(g).all(), k
        # (ref64: the wall particles of the channel whose fluid neighbours sit ON the cutoff are
        # ill-conditioned in the reference itself, see _util.assert_close)
        assert_close(k, g, ref[k], setup, what=f"{name} forward", ref64=ref64[k])
        drift_ok(k, g, ref[k], ref64[k], setup)
    # advance: searches and frozen steps in one sequence
    ref = integrator.simulate(setup, nsteps, fast_segment_sum=True)
    ref64 = integrator.simulate(setup64, nsteps, fast_segment_sum=True)
    eng.upload(setup.state)
    eng.step(setup.dt, nsteps)
    got = eng.download(host=True)
    assert eng.error() == 0
    cnt = eng.counters()
    assert cnt["tiles_without_lists"] == 0, cnt
    assert 1 <= cnt["searches"] < nsteps + 1, f"{name}: no frozen step in the sequence {cnt}"
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, factor=4.0, what=f"{name} {nsteps} steps",
                     ref64=ref64[k])


SKINS = {"off": -1.0, "tiny": 0.01, "default": 0.0, "large": 0.35}


@pytest.mark.parametrize("skin", list(SKINS))
def test_skin_does_not_change_the_result(skin):
    """The search is amortised, membership is not: any skin gives the oracle's trajectory."""
    from oracle import cases, integrator

    kw = dict(case="tgv", dim=3, dx=2 * np.pi / 20, tvf=1.0, viscosity=0.02, r0_noise_factor=0.25)
    nsteps = 30
    setup = cases.make_case(dtype=np.float32, **kw)
    ref = integrator.simulate(setup, nsteps, fast_segment_sum=True)
    eng = _engine(setup, skin=SKINS[skin])
    eng.upload(setup.state)
    eng.step(setup.dt, nsteps)
    got = eng.download(host=True)
    assert eng.error() == 0
    cnt = eng.counters()
    if skin in ("off", "tiny"):
        assert cnt["searches"] >= nsteps - 2, cnt
    if skin == "large":
        assert cnt["searches"] <= nsteps // 3, cnt
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, factor=6.0, what=f"skin {skin}")


DEGRADED = {
    # the stencil of a tile (about 3000 particles) does not fit 1024 staged particles: several
    # staging groups, no lists, every sweep searches on its own -- on frozen cells too
    "staging_groups": dict(stage_cap=1024),
    # rows of 64 entries overflow (a 3D particle has ~150 skin neighbours): same fall-back
    "row_overflow": dict(nl_cap=64),
    "lists_off": dict(nl_cap=-1),
}


@pytest.mark.parametrize("mode", list(DEGRADED))
def test_degraded_modes_3d(mode):
    from oracle import cases, integrator

    kw = dict(case="tgv", dim=3, dx=2 * np.pi / 24, tvf=1.0, viscosity=0.02, r0_noise_factor=0.25)
    nsteps = 10
    setup = cases.make_case(dtype=np.float32, **kw)
    ref = integrator.simulate(setup, nsteps, fast_segment_sum=True)
    eng = _engine(setup, **DEGRADED[mode])
    eng.upload(setup.state)
    eng.step(setup.dt, nsteps)
    got = eng.download(host=True)
    assert eng.error() == 0
    cnt = eng.counters()
    assert cnt["tiles_without_lists"] == cnt["tiles"], cnt
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, factor=4.0, what=mode)


def test_neighbor_list_after_frozen_steps_is_exact():
    """The drop-in `.idx` asked for in the middle of a run (particles drifted out of the frozen
    cells) is the brute-force set of the CURRENT positions, and the run continues correctly."""
    from oracle import cases, integrator, partition

    kw = dict(case="tgv", dim=3, dx=2 * np.pi / 16, tvf=1.0, viscosity=0.02, r0_noise_factor=0.25)
    setup = cases.make_case(dtype=np.float32, **kw)
    eng = _engine(setup)
    eng.upload(setup.state)
    eng.step(setup.dt, 4)
    n = len(setup.state["r"])
    _, count = eng.neighbor_list(0)
    idx, count2 = eng.neighbor_list(count)
    assert count2 == count
    r_now = eng.download(host=True)["r"].numpy()
    want = partition.neighbor_pairs(r_now, setup.box_size, 3.0 * setup.dx)
    got = idx.cpu().numpy()
    keys = np.sort(got[1].astype(np.int64) * n + got[0].astype(np.int64))
    wkeys = np.sort(np.asarray(want[1], dtype=np.int64) * n + np.asarray(want[0], dtype=np.int64))
    assert np.array_equal(keys, wkeys)
    eng.step(setup.dt, 4)
    ref = integrator.simulate(setup, 8, fast_segment_sum=True)
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, factor=3.0, what="after neighbor_list")
