"""On-device case initialisation (jax_sph_b200/case_setup.py,
csrc/init.cuh) against the oracle's restatement of the reference generators
(oracle/cases.py: pos_init_cartesian_2d/3d <- jax_sph/utils.py:35-54, pinned by
tests/test_reference_pins.py) and against bench.py's NumPy lattice states.

CPU part: the host logic (lattice shape rounding, spec validation, the C-side argument
checks -- no compute).  GPU part: bit-exact positions / tags / uniform fields / ids, TGV
velocities to float32 rounding of sin / cos, and an engine step from the device-made state.
"""

import ctypes as C

import numpy as np
import pytest

from jax_sph_b200 import _lib, case_setup


# ---------------------------------------------------------------- host logic (no GPU)
def test_lattice_shape_rounds_like_the_reference():
    from oracle import cases

    for box, dx in (([1.0, 1.0], 0.02), ([2 * np.pi] * 3, 2 * np.pi / 20), ([1.0, 0.23, 0.5], 1 / 40),
                    ([0.06, 1.12], 0.02), ([1.0, 2.5], 1.0)):  # 2.5 rounds half to even
        n = case_setup.lattice_shape(box, dx)
        gen = cases.pos_init_cartesian_2d if len(box) == 2 else cases.pos_init_cartesian_3d
        assert int(np.prod(n)) == len(gen(np.array(box), dx, np.float32))


def test_lattice_spec_fields_and_checks():
    lat = case_setup.lattice_spec([1.0, 0.26, 0.5], 0.02, wall_axis=1, n_walls=3, hot=(0.25, 0.75),
                                  T_hot=1.23, eta=0.01, kappa=7.313, Cp=305.27, p=5.0,
                                  planes=(5, 9))
    assert lat.struct_size == C.sizeof(_lib.Lattice) and lat.dim == 3
    assert list(lat.n) == [50, 13, 25] and (lat.k_lo, lat.k_hi) == (5, 9)
    assert lat.mass == np.float32(0.02**3) and lat.T_hot == np.float32(1.23)
    assert case_setup.lattice_rows(lat) == 50 * 13 * 4
    # nx = 855 of the heated channel: 0.5 / dx = 427.5 rounds to 428 planes, the last outside
    with pytest.raises(_lib.Sphb200Error, match="does not fit"):
        case_setup.lattice_spec([1.0, 0.2 + 6 / 855, 0.5], 1 / 855)
    with pytest.raises(_lib.Sphb200Error, match="velocity"):
        case_setup.lattice_spec([1.0, 1.0], 0.1, velocity="vortex")


def test_c_side_argument_checks():
    lib = _lib.load()
    lat = case_setup.lattice_spec([1.0, 1.0], 0.1)
    assert lib.sphb200_lattice_rows(C.byref(lat)) == 100
    for field, bad in (("struct_size", 4), ("dim", 4), ("k_hi", 11), ("k_lo", -1), ("wall_axis", 2),
                       ("velocity", _lib.VEL_TGV3D), ("velocity", 7), ("dx", 0.0)):
        cur = getattr(lat, field)
        setattr(lat, field, bad)
        assert lib.sphb200_lattice_rows(C.byref(lat)) == _lib.EINVAL, field
        setattr(lat, field, cur)
    lat.n[0] = 0
    assert lib.sphb200_lattice_rows(C.byref(lat)) == _lib.EINVAL
    lat.n[0], lat.n[1], lat.k_hi = 65536, 65536, 65536  # ids would not fit int32
    assert lib.sphb200_lattice_rows(C.byref(lat)) == _lib.EINVAL
    assert lib.sphb200_init_lattice(C.byref(lat), None, None, None) == _lib.EINVAL


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(_lib.Sphb200Error):
        case_setup.init_lattice(case_setup.lattice_spec([1.0, 1.0], 0.1))


def test_apply_state0_overwrites_fluid_entries_only():
    """case_setup.py:184-194: restart keys come from the snapshot for FLUID particles only."""
    rng = np.random.default_rng(1)
    tag = np.array([0, 1, 0, 0, 3, 1], dtype=np.int32)
    state = {"r": rng.random((6, 2), dtype=np.float32), "u": rng.random((6, 2), dtype=np.float32),
             "rho": np.ones(6, dtype=np.float32), "tag": tag}
    # the snapshot holds the same fluid particles, walls in another place of the arrays
    tag0 = np.array([1, 0, 0, 1, 0, 3], dtype=np.int32)
    snap = {"r": rng.random((6, 2), dtype=np.float32), "u": rng.random((6, 2), dtype=np.float32),
            "rho": np.full(6, 2.0, dtype=np.float32), "tag": tag0}
    before = {k: v.copy() for k, v in state.items()}
    case_setup.apply_state0(state, snap, keys=("r", "rho"))
    assert np.array_equal(state["r"][tag == 0], snap["r"][tag0 == 0])
    assert np.array_equal(state["r"][tag != 0], before["r"][tag != 0])
    assert np.array_equal(state["rho"], np.where(tag == 0, 2.0, 1.0).astype(np.float32))
    assert np.array_equal(state["u"], before["u"])  # not in keys
    with pytest.raises(ValueError, match="not found"):
        case_setup.apply_state0(state, {"tag": tag0}, keys=("r",))
    with pytest.raises(ValueError, match="Shape mismatch"):
        case_setup.apply_state0(state, {"r": snap["r"][:4], "tag": tag0[:4]}, keys=("r",))


@pytest.mark.parametrize("n_last,dx,layers,nranks", [(256, 2 * np.pi / 256, 170, 8), (427, 1 / 854, 142, 8),
                                                     (30, 1 / 30, 6, 4), (50, 0.02, 16, 3), (7, 0.1, 4, 4)])
def test_slab_planes_partition_the_lattice(n_last, dx, layers, nranks):
    """The plane ranges of the ranks of a ring are disjoint, ordered and cover the lattice, and
    agree with the rule select_own applies to a global state (slab.own_rows)."""
    from types import SimpleNamespace

    from jax_sph_b200.slab import own_rows, slab_range

    inv_cell = layers / (n_last * dx)
    ax = ((np.arange(n_last, dtype=np.float32) + np.float32(0.5)) * np.float32(dx)).astype(np.float32)
    nxt = 0
    for rank in range(nranks):
        z0, z1 = slab_range(layers, rank, nranks)
        eng = SimpleNamespace(inv_cell=inv_cell, layers=layers, z0=z0, z1=z1)
        k_lo, k_hi = case_setup.slab_planes(eng, n_last, dx)
        mine = own_rows(ax, inv_cell, layers, rank, nranks)
        if len(mine) == 0:
            assert k_lo == k_hi
            continue
        assert (k_lo, k_hi) == (int(mine[0]), int(mine[-1]) + 1) and k_hi - k_lo == len(mine)
        assert k_lo == nxt
        nxt = k_hi
    assert nxt == n_last


# ---------------------------------------------------------------- device (B200)
def _np(state):
    return {k: v.cpu().numpy() for k, v in state.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("box,dx", [([1.0, 1.0], 0.02), ([0.06, 1.12], 0.02),
                                    ([2 * np.pi] * 3, 2 * np.pi / 20), ([1.0, 0.26, 0.5], 0.02)])
def test_pos_init_cartesian_bit_exact(box, dx):
    from oracle import cases

    if len(box) == 2:
        got = case_setup.pos_init_cartesian_2d(np.array(box), dx)
        ref = cases.pos_init_cartesian_2d(np.array(box), dx, np.float32)
    else:
        got = case_setup.pos_init_cartesian_3d(np.array(box), dx)
        ref = cases.pos_init_cartesian_3d(np.array(box), dx, np.float32)
    got = got.cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.array_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("workload,nx", [("tgv2d", 50), ("tgv3d", 20)])
def test_tgv_state_matches_host_lattice_and_oracle_case(workload, nx):
    import bench
    from oracle import cases

    host, meta = bench.lattice_state(workload, nx)
    dim = meta["dim"]
    lat = case_setup.lattice_spec(meta["box"], meta["dx"], velocity=workload, eta=meta["viscosity"])
    got = _np(case_setup.init_lattice(lat))
    for k in ("r", "rho", "p", "mass", "eta", "T", "tag", "dudt", "dvdt", "drhodt", "dTdt"):
        assert np.array_equal(got[k], host[k]), k
    assert np.array_equal(got["nw"], np.zeros_like(host["r"]))
    # sin / cos: CUDA's sincosf and NumPy's float32 routines are both within 2 ulp of exact
    assert np.abs(got["u"] - host["u"]).max() <= 4e-7 and np.array_equal(got["u"], got["v"])
    # the oracle's case object (== the reference's initialize(), tests/test_reference_pins.py)
    kw = dict(tvf=1.0, viscosity=0.02) if dim == 3 else dict(tvf=1.0)
    setup = cases.make_case("tgv", dim=dim, dx=meta["dx"], dtype=np.float32, **kw)
    assert np.array_equal(got["r"], setup.state["r"])
    assert np.abs(got["u"] - setup.state["u"]).max() <= 4e-7
    for k in ("rho", "mass", "eta", "tag"):
        assert np.array_equal(got[k], setup.state[k]), k


@pytest.mark.gpu
def test_heated_channel_lattice_whole_and_slabs():
    import bench

    nx = 40
    host, meta = bench.ht3d_state(nx)
    kw = dict(wall_axis=1, n_walls=meta["n_walls"], hot=(0.25, 0.75), T_hot=1.23, p=meta["p_bg"],
              eta=meta["viscosity"], kappa=meta["kappa"], Cp=meta["Cp"])
    got = _np(case_setup.init_lattice(case_setup.lattice_spec(meta["box"], meta["dx"], **kw),
                                      with_ids=True))
    for k in host:
        assert np.array_equal(got[k], host[k]), k
    assert np.array_equal(got["ids"], np.arange(len(host["r"]), dtype=np.int32))
    assert set(np.unique(got["tag"])) == {0, 1, 3}
    # a rank's slab: planes [7, 12) of the last axis, ids into the full lattice
    n2 = meta["nxyz"][2]
    planes = np.zeros(n2, dtype=bool)
    planes[7:12] = True
    hslab, _ = bench.ht3d_state(nx, planes)
    gslab = _np(case_setup.init_lattice(
        case_setup.lattice_spec(meta["box"], meta["dx"], planes=(7, 12), **kw), with_ids=True))
    for k in hslab:
        assert np.array_equal(gslab[k], hslab[k]), k
    for k in ("r", "tag", "T"):
        assert np.array_equal(gslab[k], host[k][gslab["ids"]]), k
    # 2D slab of the TGV lattice and an empty slab
    h2, m2 = bench.lattice_state("tgv2d", 30, np.arange(30) >= 21)
    g2 = _np(case_setup.init_lattice(
        case_setup.lattice_spec(m2["box"], m2["dx"], velocity="tgv2d", eta=m2["viscosity"],
                                planes=(21, 30)), with_ids=True))
    assert np.array_equal(g2["r"], h2["r"]) and np.array_equal(g2["ids"], h2["ids"])
    empty = case_setup.init_lattice(case_setup.lattice_spec(m2["box"], m2["dx"], planes=(4, 4)))
    assert empty["r"].shape == (0, 2)


@pytest.mark.gpu
def test_engine_steps_from_the_device_made_state():
    """Same trajectory from the device-made start as from the host-made one (the velocities
    differ by sin / cos rounding only), device tensors straight into Engine.upload."""
    import bench
    from jax_sph_b200 import Engine, make_config

    host, meta = bench.lattice_state("tgv3d", 24)
    cfg = make_config(3, meta["box"], meta["dx"], meta["dt"], tvf=1.0, c_ref=meta["c_ref"],
                      p_ref=meta["p_ref"])
    res = []
    for state in (host, case_setup.init_lattice(
            case_setup.lattice_spec(meta["box"], meta["dx"], velocity="tgv3d", eta=meta["viscosity"]))):
        eng = Engine(cfg, len(host["r"]))
        eng.upload(state)
        eng.step(meta["dt"], 5)
        assert eng.error() == 0
        res.append(_np(eng.download()))
        eng.close()
    for k in ("r", "u", "rho"):
        scale = float(np.abs(res[0][k]).max())
        assert np.abs(res[1][k] - res[0][k]).max() <= 1e-5 * scale, k


@pytest.mark.gpu
@pytest.mark.parametrize("workload,nx,nranks", [("tgv3d", 24, 3), ("ht3d", 40, 2), ("tgv2d", 60, 4)])
def test_slab_ranks_start_from_device_made_planes(workload, nx, nranks):
    """case_setup.init_slab: every rank of a slab ring generates its own planes on the device;
    the ring then holds exactly the particles select_own() cuts out of the host-made global
    state, and steps to the same fields (one device, local ring: tests/test_gpu_slab.py)."""
    import bench
    from jax_sph_b200 import SlabEngine, make_config
    from jax_sph_b200.slab import assemble, step_local_ring

    host, meta = bench.lattice_state(workload, nx)
    n = len(host["r"])
    kw = dict(meta.get("cfg_kwargs", {}))

    def cfg():
        return make_config(meta["dim"], meta["box"], meta["dx"], meta["dt"], tvf=meta["tvf"],
                           c_ref=meta["c_ref"], p_ref=meta["p_ref"], **kw)

    if workload == "ht3d":
        spec = dict(wall_axis=1, n_walls=meta["n_walls"], hot=(0.25, 0.75), T_hot=1.23,
                    p=meta["p_bg"], eta=meta["viscosity"], kappa=meta["kappa"], Cp=meta["Cp"])
    else:
        spec = dict(velocity=workload, eta=meta["viscosity"])
    dev_ring = [SlabEngine(cfg(), r, nranks) for r in range(nranks)]
    host_ring = [SlabEngine(cfg(), r, nranks) for r in range(nranks)]
    total = 0
    for e, h in zip(dev_ring, host_ring):
        rows = case_setup.init_slab(e, meta["box"], meta["dx"], **spec)
        local, ids = h.select_own(host)
        h.upload(local, ids)
        assert rows == len(ids)
        got_local, got_ids = e.download()
        assert np.array_equal(np.sort(got_ids.numpy()), np.sort(ids))
        total += rows
    assert total == n
    for ring in (dev_ring, host_ring):
        step_local_ring(ring, meta["dt"], 6)
    res = []
    for ring in (dev_ring, host_ring):
        parts = []
        for e in ring:
            assert e.error(reduce=False) == 0
            local, ids = e.download()
            parts.append(({k: v.numpy() for k, v in local.items()}, ids.numpy()))
        res.append(assemble(parts, n))
    for k in ("r", "u", "rho", "T"):
        scale = float(np.abs(res[1][k]).max())
        assert np.abs(res[0][k] - res[1][k]).max() <= 1e-5 * scale, k
    assert np.array_equal(res[0]["tag"], res[1]["tag"])

