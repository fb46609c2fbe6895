"""GPU tests of the slab decomposition on ONE device.

All ranks of the ring are driven from this process (`slab.step_local_ring`): same
phases, kernels and message buffers as the multi-process NCCL path, with the four
messages of each exchange moved by device-to-device copies.  The decomposed run must
reproduce the single-engine run of the same state: identical particle set, fields
within the float32 summation-order tolerance of tests/_util.py (the order of the
particles inside a cell differs once particles have migrated, nothing else does).
The 2-GPU NCCL transport itself is covered by test_gpu_slab_nccl.py.
"""

import numpy as np
import pytest

from _util import assert_close

pytestmark = pytest.mark.gpu

CASES = {
    "tgv3d": (dict(case="tgv", dim=3, dx=2 * np.pi / 24, tvf=1.0, viscosity=0.02), [2, 3]),
    "tgv2d_rie": (dict(case="tgv", dim=2, dx=0.0125, solver="RIE", density_evolution=True), [2, 4]),
    "db2d": (dict(case="db", dim=2, dx=0.02), [2]),
    "ht3d": (dict(case="ht", dim=3, dx=0.02), [2]),
    "cf2d_free_slip": (dict(case="cf", dim=2, dx=0.02, free_slip=True), [3]),
    "db2d_renorm": (dict(case="db", dim=2, dx=0.02, density_renormalize=True), [2]),
    # Delta-SPH density diffusion: two more halo refreshes per step (L rows, gradient terms)
    "tgv3d_delta": (dict(case="tgv", dim=3, dx=2 * np.pi / 24, solver="DELTA", density_evolution=True,
                         viscosity=0.02), [2]),
    "db2d_delta": (dict(case="db", dim=2, dx=0.02, solver="DELTA", gamma=7.0, artificial_alpha=0.0),
                   [3]),
}
KEYS = ("r", "u", "v", "rho", "p", "T", "dudt", "dvdt", "drhodt", "dTdt")


def _run_pair(kw, nranks, nsteps, **tuning):
    from jax_sph_b200 import Engine, SlabEngine, config_from_setup
    from jax_sph_b200.slab import assemble, step_local_ring
    from oracle import cases

    setup = cases.make_case(dtype=np.float32, **kw)
    n = len(setup.state["r"])
    # same plan on both sides: same cells, same skin, and -- the ranks max-reduce their re-sort
    # decisions -- the same steps sort, so the two runs differ by the decomposition only
    cfg = config_from_setup(setup, **tuning)
    single = Engine(cfg, n)
    single.upload(setup.state)
    single.step(setup.dt, nsteps)
    ref = {k: v.numpy() for k, v in single.download(host=True).items()}
    assert single.error() == 0

    ring = [SlabEngine(config_from_setup(setup, **tuning), r, nranks) for r in range(nranks)]
    owners0 = np.full(n, -1)
    for e in ring:
        local, ids = e.select_own(setup.state)
        owners0[ids] = e.rank
        e.upload(local, ids)
    assert (owners0 >= 0).all()
    step_local_ring(ring, setup.dt, nsteps)
    parts, owners1 = [], np.full(n, -1)
    for e in ring:
        local, ids = e.download()
        assert e.error(reduce=False) == 0, f"rank {e.rank} device error"
        owners1[ids.numpy()] = e.rank
        parts.append(({k: v.numpy() for k, v in local.items()}, ids.numpy()))
    got = assemble(parts, n)  # raises if a particle is lost or duplicated
    return setup, ref, got, owners0, owners1, ring


@pytest.mark.parametrize("name,nranks", [(k, p) for k, (_, ps) in CASES.items() for p in ps])
def test_slab_ring_matches_single_engine(name, nranks):
    kw, _ = CASES[name]
    setup, ref, got, owners0, owners1, ring = _run_pair(kw, nranks, nsteps=12)
    for k in KEYS:
        if k in ref:
            assert_close(k, got[k], ref[k], setup, factor=4.0, what=f"{name} P={nranks}")
    for k in ("mass", "eta", "tag"):
        assert np.array_equal(got[k], ref[k]), k
    # every particle sits on the rank that owns its layer
    for e in ring:
        lay = np.clip((got["r"][:, e.axis] * np.float32(e.inv_cell)).astype(np.int64), 0, e.layers - 1)
        mine = owners1 == e.rank
        assert ((lay[mine] >= e.z0) & (lay[mine] < e.z1)).all()


def test_slab_ring_freezes_and_sorts_together():
    """The ranks agree on when to sort: after 12 steps every rank has searched the same number of
    times, fewer than 12, and as often as the single engine."""
    kw, _ = CASES["tgv3d"]
    setup, ref, got, owners0, owners1, ring = _run_pair(kw, 2, nsteps=12)
    searches = [e.counters()["searches"] for e in ring]
    assert len(set(searches)) == 1 and 1 <= searches[0] < 12, searches


def test_slab_migration_happens_and_conserves_particles():
    """2D TGV (velocity along the slab axis) long enough for particles to cross slab faces."""
    kw = dict(case="tgv", dim=2, dx=0.0125, tvf=1.0)
    setup, ref, got, owners0, owners1, ring = _run_pair(kw, 4, nsteps=150)
    moved = owners0 != owners1
    assert moved.sum() > 20, "too few particles changed rank: the test does not exercise migration"
    for k in ("r", "u", "rho"):
        assert_close(k, got[k], ref[k], setup, factor=15.0, what="150 steps")
    ek = sum(e.stats(reduce=False)[0] for e in ring)
    ek_ref = 0.5 * float((ref["mass"][:, None] * ref["u"] ** 2).sum())
    assert abs(ek - ek_ref) <= 1e-5 * ek_ref
    assert sum(e.counts()["own"] for e in ring) == len(ref["r"])


def test_slab_forward_only_is_bitwise_single_engine():
    """No integration, no migration: same cells, same in-cell order, same sums -> same bits."""
    from jax_sph_b200 import Engine, SlabEngine, config_from_setup
    from jax_sph_b200.slab import assemble, step_local_ring
    from oracle import cases

    setup = cases.make_case("tgv", dim=3, dx=2 * np.pi / 24, tvf=1.0, viscosity=0.02, dtype=np.float32)
    rng = np.random.default_rng(5)
    setup.state["r"] = np.mod(setup.state["r"] + rng.uniform(-0.2, 0.2, setup.state["r"].shape)
                              * setup.dx, setup.box_size).astype(np.float32)
    n = len(setup.state["r"])
    single = Engine(config_from_setup(setup), n)
    single.upload(setup.state)
    single.step(0.0, 1, integrate=False, bc=False)
    ref = single.download(host=True)
    ring = [SlabEngine(config_from_setup(setup), r, 2) for r in range(2)]
    for e in ring:
        e.upload(setup.state)
    step_local_ring(ring, 0.0, 1, integrate=False, bc=False)
    parts = []
    for e in ring:
        local, ids = e.download()
        assert e.error(reduce=False) == 0
        parts.append(({k: v.numpy() for k, v in local.items()}, ids.numpy()))
    got = assemble(parts, n)
    for k in ("rho", "p", "dudt", "dvdt"):
        assert np.array_equal(got[k], ref[k].numpy()), f"{k} differs from the single engine"


def test_slab_rejects_thin_slabs():
    from jax_sph_b200 import SlabEngine, _lib, config_from_setup
    from oracle import cases

    setup = cases.make_case("tgv", dim=2, dx=0.05, dtype=np.float32)  # 6 cutoffs: 13 layers
    with pytest.raises(_lib.Sphb200Error):
        SlabEngine(config_from_setup(setup), 0, 4)  # 3 layers per slab < 2 S


@pytest.mark.parametrize("name,nranks", [("tgv3d", 2), ("ht3d", 2), ("tgv2d_rie", 3)])
def test_slab_advance_host_live_fields_only(name, nranks):
    """SlabEngine.advance_host moves only the entries advance() reads (host -> device) and the
    ones it reads or writes (device -> host, the particle set changes by migration): several
    host-resident steps of the ring reproduce the resident single engine."""
    from jax_sph_b200 import Engine, SlabEngine, config_from_setup
    from jax_sph_b200.slab import assemble, step_local_ring
    from oracle import cases

    kw, _ = CASES[name]
    setup = cases.make_case(dtype=np.float32, **kw)
    n, nsteps = len(setup.state["r"]), 4
    single = Engine(config_from_setup(setup), n)
    single.upload(setup.state)
    single.step(setup.dt, nsteps)
    ref = {k: v.numpy() for k, v in single.download(host=True).items()}
    ring = [SlabEngine(config_from_setup(setup), r, nranks) for r in range(nranks)]
    read, written = ring[0].live_fields()
    host = [e.select_own(setup.state) for e in ring]
    for _ in range(nsteps):
        # advance_host of every rank, phase by phase (one process drives the whole ring here)
        for e, (local, ids) in zip(ring, host):
            e.upload({k: local[k] for k in read}, ids)
        step_local_ring(ring, setup.dt, 1)
        keys = [k for k in ref if k in read or k in written]
        host = []
        for e in ring:
            local, ids = e.download(keys)
            host.append(({k: v.numpy().copy() for k, v in local.items()}, ids.numpy().copy()))
    for e in ring:
        assert e.error(reduce=False) == 0
    got = assemble(host, n)
    assert set(got) == set(keys)
    for k in written:
        assert_close(k, got[k], ref[k], setup, factor=4.0, what=f"{name} host-resident ring")
    for k in ("mass", "eta", "tag"):
        assert np.array_equal(got[k], ref[k]), k

