"""GPU parity against output of the UNMODIFIED reference (tests/golden/ref_*.npz).

The vectors come from /root/reference's own code executed through the torch
stand-in for jax (tests/golden/make_reference_golden.py); nothing here reads
/root/reference.  For every case the CUDA engine (through the C ABI) starts
from the reference's state0 and must reproduce

* `partition.neighbor_list(...).allocate(r).idx` as a set of (sender, receiver)
  pairs, bit for bit (jax_md/partition.py:885-909);
* `advance(0.0, ...)` = WCSPH.forward + bc_fn (solver.py:705-949) and the state
  after 20 `advance(dt, ...)` calls (integrator.py:22-56) within relative 1e-5
  (+ the float32 EoS noise floor, tests/_util.py) of the reference's float32
  run, and no further from its float64 run than that float32 run is (drift bound).
"""

import glob
import json
import os

import numpy as np
import pytest

from _util import GOLDEN, assert_close, drift_ok

pytestmark = pytest.mark.gpu

NAMES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
FWD_KEYS = ("rho", "p", "u", "v", "dudt", "dvdt", "drhodt", "T", "dTdt")
ADV_KEYS = ("r", "u", "v", "rho", "p", "T", "dudt", "dvdt")


def _load(name):
    from oracle import cases

    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    kw = json.loads(str(z["make_case_kwargs"]))
    meta = json.loads(str(z["meta_f32"]))
    # the Setup carries the solver switches, EoS scalars and bc/g_ext tables of the case (the
    # CPU suite checks it against the reference's initialize()); the particles are the reference's
    setup = cases.make_case(dtype=np.float32, **kw)
    state0 = {k: np.array(z[f"state0_f32_{k}"]) for k in setup.state}
    assert abs(setup.dt - meta["dt"]) <= 1e-12 * meta["dt"]
    setup.state = state0
    return z, setup, state0, meta


def _engine(setup, **tuning):
    from jax_sph_b200 import Engine, config_from_setup

    return Engine(config_from_setup(setup, **tuning), len(setup.state["r"]))


def _unpack(z, key):
    counts, recv = z[key + "_counts"], z[key + "_recv"]
    n = len(counts)
    return np.repeat(np.arange(n, dtype=np.int64), counts) * n + recv.astype(np.int64)


def _keys(idx, n):
    idx = np.asarray(idx)
    ok = (idx[0] < n) & (idx[1] < n)
    return np.sort(idx[1][ok].astype(np.int64) * n + idx[0][ok].astype(np.int64))


def test_reference_goldens_present():
    assert len(NAMES) >= 10


@pytest.mark.parametrize("name", NAMES)
def test_neighbor_list_is_the_references(name):
    import torch

    from jax_sph_b200 import partition, space

    z, setup, state0, meta = _load(name)
    box = np.asarray(meta["box_size"], dtype=np.float64)
    disp, _ = space.periodic(box)
    fns = partition.neighbor_list(disp, box, meta["cutoff"], mask_self=False)
    n = meta["n"]
    for which, r in (("pairs_f32", state0["r"]), ("pairs_end_f32", z["advance_f32_r"])):
        nbrs = fns.allocate(torch.as_tensor(np.ascontiguousarray(r), device="cuda"))
        assert not nbrs.did_buffer_overflow
        got = _keys(nbrs.idx.cpu().numpy(), n)
        assert np.array_equal(got, _unpack(z, which)), f"{name}: {which} differs"


@pytest.mark.parametrize("name", NAMES)
def test_forward_matches_reference(name):
    z, setup, state0, meta = _load(name)
    eng = _engine(setup)
    eng.upload(state0)
    eng.step(0.0, 1)  # simulate.py:110: advance(0.0, state, neighbors)
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in FWD_KEYS:
        g = got[k].numpy()
        assert np.isfinite(g).all(), k
        assert_close(k, g, z[f"forward_f32_{k}"], setup, what=f"{name} forward")
        drift_ok(k, g, z[f"forward_f32_{k}"], z[f"forward_f64_{k}"], setup)


@pytest.mark.parametrize("name", NAMES)
def test_advance_20_steps_matches_reference(name):
    z, setup, state0, meta = _load(name)
    eng = _engine(setup)
    eng.upload(state0)
    eng.step(meta["dt"], meta["nsteps"])
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in ADV_KEYS:
        g = got[k].numpy()
        # 20 steps let the per-step float32 noise accumulate: allow sqrt(20) ~ 5 units
        assert_close(k, g, z[f"advance_f32_{k}"], setup, factor=5.0, what=f"{name} advance")
        drift_ok(k, g, z[f"advance_f32_{k}"], z[f"advance_f64_{k}"], setup, factor=3.0)


# ---- 200-step trajectories (north_star: "tolerance-matched 200-step trajectories") ----------
LONG_NAMES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "ref200_*.npz")))


def test_long_goldens_present():
    assert len(LONG_NAMES) >= 6


@pytest.mark.parametrize("name", LONG_NAMES)
def test_200_step_trajectory_matches_reference(name):
    """The loop of simulate.py:114-131 for 200 steps from the reference's state0: engine vs the
    reference's float32 run (15 tolerance units: per-step float32 noise accumulates) and the
    drift bound against its float64 run."""
    from oracle import cases

    z = np.load(os.path.join(GOLDEN, f"ref200_{name}.npz"))
    kw = json.loads(str(z["make_case_kwargs"]))
    meta = json.loads(str(z["meta_f32"]))
    setup = cases.make_case(dtype=np.float32, **kw)
    setup.state = {k: np.array(z[f"state0_f32_{k}"]) for k in setup.state}
    eng = _engine(setup)
    eng.upload(setup.state)
    eng.step(meta["dt"], meta["nsteps"])
    got = eng.download(host=True)
    assert eng.error() == 0
    for k in ("r", "u", "v", "rho", "T"):
        g = got[k].numpy()
        assert_close(k, g, z[f"advance_f32_{k}"], setup, factor=15.0, what=f"{name} 200 steps")
        drift_ok(k, g, z[f"advance_f32_{k}"], z[f"advance_f64_{k}"], setup, factor=3.0)
    ek = 0.5 * float((got["mass"].numpy()[:, None] * got["u"].numpy() ** 2).sum())
    ref_u, ref_m = z["advance_f32_u"], z["state0_f32_mass"]
    ek_ref = 0.5 * float((ref_m[:, None] * ref_u ** 2).sum())
    assert abs(ek - ek_ref) <= 2e-5 * max(ek_ref, 1e-30), "kinetic energy after 200 steps"
