"""GPU tests of the duo sweeps (jax_sph_b200/csrc/sweep2.cuh).

The headline SPH variants (summation density, compact force record: BASELINE configs[0], [1],
[3]) run with two slot neighbours per thread on one union neighbour list, staged by bulk copies
from per-tile descriptors.  The parity suites (test_gpu_reference.py, test_gpu_parity*.py) already
run those variants through this path against the reference's vectors and the oracle; here:

* which variants take the path and which do not (device counters),
* duo == per-particle sweeps of the same library (SPHB200_DUO=0) within the parity tolerance,
  forward and over a run that re-sorts several times, on disordered and CLUSTERED particles
  (runs of odd length, duos whose particles sit in different cells, half-empty tiles),
* a mix of tiles with and without lists in one step (a staging capacity that only the crowded
  tiles exceed): the not-ok tiles are swept by sweep.cuh kernels launched over them alone,
* the engine against the oracle (solver.py:705-949, integrator.py:22-56) on the clustered state.

Tolerances: tests/_util.py.
"""

import os

import numpy as np
import pytest

from _util import assert_close

pytestmark = pytest.mark.gpu

FWD_KEYS = ("rho", "p", "dudt", "dvdt")
ADV_KEYS = ("r", "u", "v", "rho", "p", "dudt", "dvdt")


def _engine(setup, duo=True, **tuning):
    from jax_sph_b200 import Engine, config_from_setup

    old = os.environ.get("SPHB200_DUO")
    os.environ["SPHB200_DUO"] = "1" if duo else "0"
    try:
        return Engine(config_from_setup(setup, **tuning), len(setup.state["r"]))
    finally:
        if old is None:
            os.environ.pop("SPHB200_DUO")
        else:
            os.environ["SPHB200_DUO"] = old


def _clustered(setup, seed, amp=None):
    """The lattice with a smooth displacement field on top: density varies by tens of per cent
    from place to place (crowded and sparse tiles), neighbours stay distinct."""
    rng = np.random.default_rng(seed)
    if amp is None:
        amp = 0.2 if setup.dim == 3 else 0.6
    st = {k: v.copy() for k, v in setup.state.items()}
    box = np.asarray(setup.box_size, dtype=np.float64)
    r = st["r"].astype(np.float64)
    shift = np.zeros_like(r)
    for a in range(setup.dim):
        ph = rng.uniform(0, 2 * np.pi, setup.dim)
        arg = sum(2 * np.pi * r[:, b] / box[b] * (1 + (a + b) % 2) + ph[b] for b in range(setup.dim))
        shift[:, a] = amp * 3.0 * setup.dx * np.sin(arg)
    r = r + shift + rng.normal(0.0, 0.1 * setup.dx, r.shape)
    st["r"] = np.mod(r, box).astype(np.float32)
    return st


VARIANTS = [
    (dict(case="tgv", dim=3, dx=2 * np.pi / 32, tvf=1.0, viscosity=0.02), True),
    (dict(case="tgv", dim=3, dx=2 * np.pi / 32, tvf=0.0, viscosity=0.02), True),
    (dict(case="tgv", dim=2, dx=1.0 / 100, tvf=1.0), True),
    (dict(case="tgv", dim=2, dx=1.0 / 100, kernel="WC2K", h_factor=1.3), True),
    (dict(case="tgv", dim=2, dx=1.0 / 100, solver="RIE", density_evolution=True), False),
    (dict(case="db", dim=2, dx=0.02), False),
    (dict(case="tgv", dim=3, dx=2 * np.pi / 8, tvf=1.0, viscosity=0.02), False),  # box too small
]


@pytest.mark.parametrize("kw,want", VARIANTS)
def test_which_variants_take_the_duo_path(kw, want):
    from oracle import cases

    setup = cases.make_case(dtype=np.float32, **kw)
    eng = _engine(setup)
    eng.upload(setup.state)
    eng.step(setup.dt, 1)
    assert eng.error() == 0
    assert eng.counters()["duo"] == want
    assert not _engine(setup, duo=False).counters()["duo"]


CASES = {
    "tgv3d_tvf": dict(case="tgv", dim=3, dx=2 * np.pi / 40, tvf=1.0, viscosity=0.02),
    "tgv3d_plain": dict(case="tgv", dim=3, dx=2 * np.pi / 32, tvf=0.0, viscosity=0.02),
    "tgv2d_tvf": dict(case="tgv", dim=2, dx=1.0 / 160, tvf=1.0),
    "tgv2d_wc2k": dict(case="tgv", dim=2, dx=1.0 / 120, kernel="WC2K", h_factor=1.3),
}


@pytest.mark.parametrize("name", list(CASES))
def test_duo_equals_per_particle_sweeps(name):
    from oracle import cases

    setup = cases.make_case(dtype=np.float32, **CASES[name])
    state = _clustered(setup, seed=len(name))
    nsteps = 25
    out = {}
    for duo in (False, True):
        eng = _engine(setup, duo=duo)
        eng.upload(state)
        eng.step(0.0, 1, integrate=False, bc=False)
        fwd = {k: v.numpy().copy() for k, v in eng.download(host=True).items()}
        eng.upload(state)
        eng.step(setup.dt, nsteps)
        adv = {k: v.numpy().copy() for k, v in eng.download(host=True).items()}
        assert eng.error() == 0
        cnt = eng.counters()
        # (a few crowded tiles may overflow their list rows as the particles clump: they are swept
        # by the fall-back kernels, which is part of what is compared here)
        assert cnt["duo"] == duo and cnt["tiles_without_lists"] <= cnt["tiles"] // 4, cnt
        assert 2 <= cnt["searches"] < nsteps, cnt  # searches AND frozen steps
        out[duo] = (fwd, adv)
    for k in FWD_KEYS:
        assert_close(k, out[True][0][k], out[False][0][k], setup, what=f"{name} forward, duo vs classic")
    for k in ADV_KEYS:
        assert_close(k, out[True][1][k], out[False][1][k], setup, factor=4.0,
                     what=f"{name} {nsteps} steps, duo vs classic")


@pytest.mark.parametrize("name", ["tgv3d_tvf", "tgv2d_tvf"])
def test_mixed_tiles_with_and_without_lists(name):
    """A staging capacity between the emptiest and the most crowded stencil: some tiles are swept
    by the duo kernels, the rest by the per-particle kernels launched over the not-ok tiles."""
    from oracle import cases

    setup = cases.make_case(dtype=np.float32, **CASES[name])
    state = _clustered(setup, seed=7, amp=0.25 if setup.dim == 3 else 0.7)
    # mean stencil population of a tile, from the plan
    plan = _engine(setup).plan()
    cells = [t + 2 * s_ for t, s_, n in zip(plan["tile"], plan["sub"], plan["cells"]) if n > 1]
    pop = len(state["r"]) / float(np.prod([n for n in plan["cells"] if n > 1]))
    nominal = pop * float(np.prod(cells))
    ref = _engine(setup, duo=False)
    ref.upload(state)
    ref.step(setup.dt, 6)
    want = {k: v.numpy().copy() for k, v in ref.download(host=True).items()}
    mixed = None
    for frac in (0.95, 1.0, 1.05, 0.9, 1.1, 0.85, 1.2):
        eng = _engine(setup, stage_cap=int(nominal * frac) // 32 * 32)
        eng.upload(state)
        eng.step(setup.dt, 6)
        cnt = eng.counters()
        assert eng.error() == 0 and cnt["duo"]
        if 0 < cnt["tiles_without_lists"] < cnt["tiles"]:
            mixed = (eng, cnt)
            break
    assert mixed is not None, f"no staging capacity near {nominal:.0f} gave a mix of ok and not-ok tiles"
    got = mixed[0].download(host=True)
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), want[k], setup, factor=3.0, what=f"{name} mixed tiles {mixed[1]}")


@pytest.mark.parametrize("name", ["tgv3d_tvf", "tgv2d_tvf"])
def test_duo_vs_oracle_on_clustered_particles(name):
    from oracle import cases, integrator

    kw = dict(CASES[name])
    if kw["dim"] == 3:
        kw["dx"] = 2 * np.pi / 32
    setup = cases.make_case(dtype=np.float32, **kw)
    setup.state = _clustered(setup, seed=3)
    nsteps = 8
    ref = integrator.simulate(setup, nsteps, fast_segment_sum=True)
    eng = _engine(setup)
    eng.upload(setup.state)
    eng.step(setup.dt, nsteps)
    got = eng.download(host=True)
    assert eng.error() == 0 and eng.counters()["duo"]
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, factor=4.0, what=f"{name} clustered vs oracle")


@pytest.mark.parametrize("name", ["tgv3d_tvf", "tgv2d_tvf"])
def test_uniform_viscosity_hint(name):
    """SPHB200_HINT_UNIFORM_ETA (every reference case sets eta = viscosity, case_setup.py:152-181):
    the duo force sweep stages no eta column and keeps eta_ij of solver.py:243 per duo.  Same
    operations on the same values per pair; the smaller record buys a wider tile (384 threads), so
    the pair sums run in another order and the results agree to rounding, not to the bit.  A state that breaks the promise raises SPHB200_ERR_HINT, and a particle-wise viscosity without
    the hint still follows the oracle."""
    from jax_sph_b200 import _lib
    from oracle import cases, integrator

    kw = dict(CASES[name])
    if kw["dim"] == 3:
        kw["dx"] = 2 * np.pi / 32
    setup = cases.make_case(dtype=np.float32, **kw)
    setup.state = _clustered(setup, seed=5)
    out = {}
    for hint in (False, True):
        eng = _engine(setup, uniform_eta=hint)
        eng.upload(setup.state)
        eng.step(setup.dt, 6)
        assert eng.error() == 0 and eng.counters()["duo"]
        out[hint] = {k: v.numpy().copy() for k, v in eng.download(host=True).items()}
    for k in ADV_KEYS:
        assert_close(k, out[True][k], out[False][k], setup, factor=2.0, what=f"{name}: hint vs staged eta")
    # particle-wise viscosity: the promise is broken -> error bit; without the hint -> the oracle
    rng = np.random.default_rng(11)
    eta = setup.state["eta"] * rng.uniform(0.5, 1.5, len(setup.state["eta"])).astype(np.float32)
    setup.state = dict(setup.state, eta=eta)
    eng = _engine(setup, uniform_eta=True)
    eng.upload(setup.state)
    eng.step(setup.dt, 1)
    assert eng.error() & _lib.ERR_HINT
    eng = _engine(setup)  # config_from_setup reads the state: no hint
    eng.upload(setup.state)
    eng.step(setup.dt, 4)
    assert eng.error() == 0 and eng.counters()["duo"]
    got = eng.download(host=True)
    ref = integrator.simulate(setup, 4, fast_segment_sum=True)
    for k in ADV_KEYS:
        assert_close(k, got[k].numpy(), ref[k], setup, factor=3.0, what=f"{name} particle-wise eta vs oracle")


def _brute_force_pairs(r, box, cutoff):
    """Directed pairs (i, j), self pairs included, with the reference's float32 test
    d^2 < cutoff^2 on the periodic displacement of space.py:170-181 (torch on the GPU: float32
    add / sub / mul / fmod with the same roundings, no fused multiply-add across statements)."""
    import torch

    r = torch.as_tensor(np.asarray(r, dtype=np.float32), device="cuda")
    box = torch.as_tensor(np.asarray(box, dtype=np.float32), device="cuda")
    half = box * 0.5
    c2 = torch.tensor(np.float32(cutoff) * np.float32(cutoff), device="cuda")
    total = 0
    for i0 in range(0, len(r), 1024):
        d = r[i0:i0 + 1024, None, :] - r[None, :, :]
        d = torch.remainder(d + half, box) - half
        d2 = d * d
        s = d2[..., 0] + d2[..., 1]
        if r.shape[1] == 3:
            s = s + d2[..., 2]
        total += int((s < c2).sum().item())
    return total


def _engine_rel(setup, rel):
    old = os.environ.get("SPHB200_REL_DRIFT")
    os.environ["SPHB200_REL_DRIFT"] = rel
    try:
        return _engine(setup)
    finally:
        if old is None:
            os.environ.pop("SPHB200_REL_DRIFT")
        else:
            os.environ["SPHB200_REL_DRIFT"] = old


@pytest.mark.parametrize("rel", ["0", "1"])
@pytest.mark.parametrize("name", ["tgv3d_tvf", "tgv2d_tvf"])
def test_frozen_lists_hold_every_pair(name, rel):
    """The exact lists of a step taken long after the last search hold exactly the pairs of a
    brute-force search at that step's positions (membership bits counted on the device,
    sphb200_engine_counters), with the absolute re-sort criterion alone (a particle further than
    half the skin from where it was sorted) and with the relative one beside it (cells.cuh,
    k_drift_box: no block of neighbours has drifted apart by more than the skin), which never
    searches more often."""
    from oracle import cases

    kw = dict(CASES[name])
    kw["dx"] = 2 * np.pi / 28 if kw["dim"] == 3 else 1.0 / 120
    setup = cases.make_case(dtype=np.float32, r0_noise_factor=0.1, **kw)
    eng = _engine_rel(setup, rel)
    eng.upload(setup.state)
    cutoff = 3.0 * setup.dx
    done = 0
    for nsteps in (1, 9, 30, 40):
        eng.step(setup.dt, nsteps)
        done += nsteps
        cnt = eng.counters()
        assert eng.error() == 0 and cnt["duo"] and cnt["tiles_without_lists"] == 0
        r_now = eng.download(host=True)["r"].numpy()
        assert cnt["pairs"] == _brute_force_pairs(r_now, setup.box_size, cutoff), (name, rel, done, cnt)
    assert cnt["searches"] < done
    if rel == "1":
        other = _engine_rel(setup, "0")
        other.upload(setup.state)
        other.step(setup.dt, done)
        assert cnt["searches"] <= other.counters()["searches"], (cnt, other.counters())
