"""CPU tests of the host-side logic: config marshalling and the table forms of
the case callables (bc_fn / g_ext_fn) that the fused CUDA epilogue consumes."""

import numpy as np
import pytest

from oracle import cases
from oracle.solver import DIRICHLET_WALL, FLUID, MOVING_WALL, SOLID_WALL


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g

    g.build()


def apply_bc_table(table, state, dim):
    """NumPy statement of what k_sweep<PhysForce>::finish + k_bc do with the table."""
    s = {k: v.copy() for k, v in state.items()}
    for tag, rule in table["tags"].items():
        m = s["tag"] == tag
        if "u" in rule:
            s["u"][m] = np.asarray(rule["u"][:dim], dtype=s["u"].dtype)
        if "v" in rule:
            s["v"][m] = np.asarray(rule["v"][:dim], dtype=s["v"].dtype)
        if rule.get("zero_dudt"):
            s["dudt"][m] = 0
        if rule.get("zero_dvdt"):
            s["dvdt"][m] = 0
        if "p" in rule:
            s["p"][m] = rule["p"]
        if "T" in rule:
            s["T"][m] = rule["T"]
        if rule.get("zero_dTdt"):
            s["dTdt"][m] = 0
    fluid = s["tag"] == FLUID
    if table.get("inflow_x"):
        m = fluid & (s["r"][:, 0] < np.float32(table["inflow_x"]["x"]))
        s["T"][m] = table["inflow_x"]["T"]
        s["dTdt"][m] = 0
    if table.get("outflow_x"):
        m = fluid & (s["r"][:, 0] > np.float32(table["outflow_x"]["x"]))
        s["dTdt"][m] = 0
    return s


@pytest.mark.parametrize("case,dim", [("tgv", 2), ("db", 2), ("pf", 2), ("cf", 2), ("ht", 2), ("ht", 3)])
def test_bc_table_equals_case_callable(case, dim):
    setup = cases.make_case(case, dim=dim, dx=0.05, dtype=np.float32)
    rng = np.random.default_rng(1)
    state = {k: (rng.standard_normal(v.shape).astype(v.dtype) if v.dtype == np.float32 else v.copy())
             for k, v in setup.state.items()}
    state["r"] = setup.state["r"].copy()
    want = setup.bc_fn({k: v.copy() for k, v in state.items()})
    got = apply_bc_table(setup.bc_table, state, dim)
    for k in want:
        assert np.array_equal(want[k], got[k]), f"{case}: table != bc_fn for {k}"


@pytest.mark.parametrize("case,dim", [("tgv", 3), ("db", 2), ("pf", 2), ("ht", 3)])
def test_g_ext_spec_equals_case_callable(case, dim):
    setup = cases.make_case(case, dim=dim, dx=0.05, dtype=np.float32)
    r = setup.state["r"]
    want = setup.g_ext_fn(r)
    spec = setup.g_ext_spec
    got = np.zeros_like(r)
    if spec["mode"] == "const":
        got[:] = np.asarray(spec["g"][:dim], dtype=np.float32)
    elif spec["mode"] == "band":
        x = r[:, spec["axis"]]
        m = (x < np.float32(spec["hi"])) & (x > np.float32(spec["lo"]))
        got[m] = np.asarray(spec["g"][:dim], dtype=np.float32)
    assert np.array_equal(want.astype(np.float32), got)


def test_config_from_setup_fields():
    from jax_sph_b200 import _lib, config_from_setup

    setup = cases.make_case("ht", dim=3, dx=0.05, dtype=np.float32)
    cfg = config_from_setup(setup)
    assert cfg.dim == 3 and cfg.solver == 0 and cfg.kernel == 0 and cfg.eos == _lib.EOS_TAIT
    assert cfg.flags == _lib.F_BC_TRICK | _lib.F_HEAT
    assert list(cfg.box) == pytest.approx(list(setup.box_size))
    assert cfg.h == setup.dx and cfg.dt == setup.dt and cfg.p_bg == pytest.approx(0.05 * setup.p_ref)
    assert cfg.g_mode == _lib.G_BAND and cfg.g_axis == 1 and cfg.g[0] == pytest.approx(2.3)
    assert cfg.bc[SOLID_WALL].flags & _lib.BC_SET_T and cfg.bc[DIRICHLET_WALL].T == pytest.approx(1.23)
    assert cfg.bc[MOVING_WALL].flags == 0 and cfg.bc_inflow_on == 1 and cfg.bc_outflow_on == 1
    rie = cases.make_case("tgv", dim=2, dx=0.05, dtype=np.float32, solver="RIE", density_evolution=True)
    cfg = config_from_setup(rie, cell_sub=[1, 1], tile=[4, 1], threads=128)
    assert cfg.solver == 1 and cfg.eos == _lib.EOS_RIEMANN and cfg.flags == _lib.F_RHO_EVOL
    assert list(cfg.cell_sub) == [1, 1, 0] and list(cfg.tile) == [4, 1, 0] and cfg.threads == 128


def test_engine_bytes_scale_with_features():
    import ctypes as C

    from jax_sph_b200 import _lib, config_from_setup

    lib = _lib.load()

    def nbytes(setup, n=100000, **tuning):
        b = C.c_size_t()
        _lib.check(lib.sphb200_engine_bytes(C.byref(config_from_setup(setup, **tuning)), n, C.byref(b)))
        return b.value

    tgv = cases.make_case("tgv", dim=3, dx=0.2, dtype=np.float32)
    plain = nbytes(tgv, nl_cap=-1)
    heat = nbytes(cases.make_case("ht", dim=3, dx=0.05, dtype=np.float32), nl_cap=-1)
    assert 200 * 100000 < plain < 260 * 100000  # ~212 B / particle + cell tables
    assert heat > plain  # kappa / Cp carried only when heat conduction is on
    # neighbour lists: the skin list and the exact list of the step, rows of 1.3 x 159 + 8 -> 216
    # uint16 entries (159 = particles within 1.12 cutoffs: default skin of the 3D duo sweeps) + a
    # count per particle each, + the tile descriptors, compact force records and duo rows of
    # sweep2.cuh (a few tens of bytes per particle)
    lists = nbytes(tgv) - plain
    assert 2 * 216 * 2 * 100000 - 65536 <= lists < 2 * (216 * 2 + 40) * 100000 + 4096
    # without a skin the rows hold the 113 neighbours of the cutoff sphere: 160 entries
    lists = nbytes(tgv, skin=-1.0) - plain
    assert 2 * 160 * 2 * 100000 <= lists < 2 * (160 * 2 + 40) * 100000 + 4096


def test_bench_ht3d_lattice_is_the_reference_case():
    """bench.py's rank-local generator of BASELINE configs[4] (3D heated channel) produces the
    particles, tags, fields, dt and bc / g_ext tables of the case setup (oracle.cases.make_case,
    itself pinned against SimulationSetup.initialize by tests/test_reference_pins.py)."""
    import bench
    from oracle import cases

    nx = 25
    setup = cases.make_case("ht", dim=3, dx=1.0 / nx, dtype=np.float32, r0_noise_factor=0.0)
    state, meta = bench.ht3d_state(nx)
    assert len(state["r"]) == len(setup.state["r"]) == int(np.prod(meta["nxyz"]))
    assert abs(meta["dt"] - setup.dt) <= 1e-12 * setup.dt
    assert np.allclose(meta["box"], setup.box_size, rtol=1e-7)
    assert meta["cfg_kwargs"]["bc_table"] == setup.bc_table
    g1, g2 = meta["cfg_kwargs"]["g_ext_spec"], setup.g_ext_spec
    assert g1["mode"] == g2["mode"] and g1["axis"] == g2["axis"] and np.allclose(g1["g"], g2["g"])
    assert abs(g1["lo"] - g2["lo"]) < 1e-9 and abs(g1["hi"] - g2["hi"]) < 1e-9
    assert abs(meta["p_bg"] - setup.p_bg) < 1e-12 and abs(meta["p_ref"] - setup.p_ref) < 1e-12

    def order(r):  # same lattice, different enumeration order
        q = np.rint(r * nx - 0.5).astype(np.int64)
        return np.lexsort((q[:, 0], q[:, 1], q[:, 2]))

    a, b = order(state["r"]), order(setup.state["r"])
    for k in ("r", "tag", "u", "v", "rho", "p", "mass", "eta", "T", "kappa", "Cp", "dudt"):
        x, y = state[k][a], setup.state[k][b]
        assert np.allclose(x, y, rtol=0, atol=2e-7 * max(1.0, float(np.abs(y).max()))), k
    # a slab's planes carry the ids of the full enumeration
    sub, _ = bench.ht3d_state(nx, planes=np.arange(meta["nxyz"][2]) % 3 == 1)
    assert len(np.unique(sub["ids"])) == len(sub["ids"]) < len(state["r"])


def test_eos_records_keep_the_reference_forms_and_feed_the_engine_config():
    """jax_sph_b200.eos: closed forms of jax_sph/eos.py:20-57 (value-identical to the oracle's
    restatement in float64), reference attribute names, and the make_config keywords."""
    from jax_sph_b200 import eos, make_config
    from oracle import eos as oeos

    rho = np.linspace(0.9, 1.1, 21)
    t, r = eos.TaitEoS(100.0, 1.0, 5.0, 7.0), eos.RIEMANNEoS(1.0, 0.5, 1.25)
    ot, orr = oeos.TaitEoS(100.0, 1.0, 5.0, 7.0), oeos.RIEMANNEoS(1.0, 0.5, 1.25)
    assert np.array_equal(t.p_fn(rho), ot.p_fn(rho)) and np.array_equal(r.p_fn(rho), orr.p_fn(rho))
    assert np.allclose(t.rho_fn(t.p_fn(rho)), rho, rtol=1e-13)
    assert np.allclose(r.rho_fn(r.p_fn(rho)), rho, rtol=1e-13)
    assert (t.p_ref, t.rho_ref, t.p_bg, t.gamma) == (100.0, 1.0, 5.0, 7.0) and not hasattr(t, "u_ref")
    assert (r.rho_ref, r.p_bg, r.u_ref) == (1.0, 0.5, 1.25)
    cfg = make_config(2, [1.0, 1.0], 0.02, 1e-4, **t.engine_fields())
    assert cfg.eos == 0 and cfg.p_ref == 100.0 and cfg.gamma == 7.0 and cfg.p_bg == 5.0
    cfg = make_config(2, [1.0, 1.0], 0.02, 1e-4, solver="RIE", **r.engine_fields())
    assert cfg.eos == 1 and cfg.u_ref == 1.25 and cfg.p_bg == 0.5
    with pytest.raises(ValueError):
        eos.TaitEoS(100.0, 0.0, 0.0, 1.0)


def test_cpu_arm_worker_and_aggregation(monkeypatch):
    """bench.py's CPU arm: one oracle worker per core (oracle/cpu_arm.py), aggregate over the
    common window.  Tiny sample, two workers."""
    import argparse
    import json
    import subprocess
    import sys

    import bench

    out = subprocess.run([sys.executable, "-m", "oracle.cpu_arm", "--workload", "tgv2d", "--nx", "12",
                          "--warmup", "0", "--steps", "1"], cwd=bench.ROOT, capture_output=True,
                         text=True, check=True).stdout
    rec = json.loads(out.strip().splitlines()[-1])
    assert rec["n"] == 144 and rec["steps"] == 1 and rec["t1"] > rec["t0"]
    monkeypatch.setattr(bench, "cpu_workers", lambda: 2)
    args = argparse.Namespace(workload="tgv2d", cpu_nx=12, cpu_steps=2)
    r = bench.run_cpu_arm(args, 2, warmup=0)
    assert r["workers"] == 2 and r["n"] == 144 and r["window_s"] > 0
    # the aggregate cannot exceed the sum of the workers' own rates
    assert 0 < r["value"] <= 2 * r["per_core"] * 1.0001
    base = bench.cpu_baseline(args)
    assert base["cores"] == 2 and base["kind"] == "port" and "2 single-core workers" in base["sample"]


def test_bench_db2d_generator_is_the_reference_case():
    """bench.py's generator of BASELINE configs[2] (2D dam break) lays out the particles, tags,
    fields, dt, box and bc / g_ext tables of the case setup (oracle.cases.make_case, itself pinned
    against SimulationSetup.initialize by tests/test_reference_pins.py), bit for bit."""
    import bench
    from oracle import cases

    dx = 0.02
    state, meta = bench.db2d_state(dx)
    setup = cases.make_case("db", dim=2, dx=dx, dtype=np.float32)
    assert len(state["r"]) == len(setup.state["r"])
    for k, v in state.items():
        assert np.array_equal(v, setup.state[k]), k
    assert meta["dt"] == setup.dt and np.allclose(meta["box"], setup.box_size, rtol=0, atol=1e-12)
    assert meta["cfg_kwargs"]["bc_table"] == setup.bc_table
    assert meta["cfg_kwargs"]["g_ext_spec"] == setup.g_ext_spec
    assert setup.is_bc_trick and setup.density_evolution and setup.artificial_alpha == 0.1
    # 4 028 622 particles at dx = 0.00071 (counted, not generated here)
    n_f = round(2.0 / 0.00071) * round(1.0 / 0.00071)
    dxn = 3 * 0.00071
    n_w = 2 * round(dxn / 0.00071) * round((2.0 + 2 * dxn) / 0.00071) + 2 * round(5.366 / 0.00071) * 3
    assert n_f + n_w == 4028622


@pytest.mark.parametrize("n_cells,S", [((11, 12, 13), 2), ((12, 13, 1), 2), ((16, 17, 19), 3)])
def test_relative_drift_window_bounds_every_pair(n_cells, S):
    """NumPy model of the relative re-sort criterion (csrc/cells.cuh: k_drift_box / k_drift_join):
    particles are binned into the cells of the frozen table, cells into blocks of S^dim cells (the
    last block of an axis may be partial), every block joins -- axis after axis -- the drift boxes
    of the blocks that hold a cell within 2 S cells of its own cells (periodic in cells).  Then for
    EVERY pair of particles either (a) |disp_i - disp_j| is at most the diagonal of the joined box
    of i's block, or (b) the two were at least 2 S cells apart along some axis at the sort -- the
    two cases of the sufficiency argument in cells.cuh."""
    rng = np.random.default_rng(sum(n_cells) + S)
    dim = 3 if n_cells[2] > 1 else 2
    n_cells = np.array(n_cells[:dim])
    box = np.array([1.0, 0.9, 1.1][:dim])
    cell = box / n_cells
    npart = 800
    rb = rng.uniform(0, 1, (npart, dim)) * box
    amp = 0.05 * S * cell.min()
    disp = amp * np.sin(2 * np.pi * rb / box + rng.uniform(0, 6, dim)) + rng.normal(0, 0.2 * amp, (npart, dim))
    cidx = np.minimum((rb / cell).astype(int), n_cells - 1)
    nb = (n_cells + S - 1) // S
    bidx = cidx // S
    lo = np.full(tuple(nb) + (dim,), np.inf)
    hi = np.full(tuple(nb) + (dim,), -np.inf)
    for p in range(npart):
        b = tuple(bidx[p])
        lo[b] = np.minimum(lo[b], disp[p])
        hi[b] = np.maximum(hi[b], disp[p])

    def window_blocks(bc, a):  # the kernel's rule, one axis
        out, last = [], -1
        for off in range(-2 * S, 3 * S):
            blk = int(((S * bc + off) % n_cells[a]) // S)
            if blk != last:
                out.append(blk)
            last = blk
        return out

    for a in range(dim):  # the separable passes of k_drift_join
        lo2, hi2 = np.empty_like(lo), np.empty_like(hi)
        for b in np.ndindex(*nb):
            src = [tuple(b[:a]) + (q,) + tuple(b[a + 1:]) for q in window_blocks(b[a], a)]
            lo2[b] = np.min([lo[s] for s in src], axis=0)
            hi2[b] = np.max([hi[s] for s in src], axis=0)
        lo, hi = lo2, hi2
    spread = np.sqrt((np.maximum(hi - lo, 0.0) ** 2).sum(-1))
    rel = np.sqrt(((disp[:, None, :] - disp[None, :, :]) ** 2).sum(-1))
    # cell gap of j from the cells [S b, S b + S) of i's block, periodic, per axis
    b0 = (bidx * S)[:, None, :]
    cj = cidx[None, :, :]
    fwd = np.mod(cj - b0, n_cells)              # 0 .. S-1 inside the block
    gap = np.minimum(np.maximum(fwd - (S - 1), 0), np.mod(b0 - cj, n_cells))
    in_window = (gap <= 2 * S).all(-1)
    bound = spread[tuple(bidx.T)][:, None]
    assert (rel[in_window] <= np.broadcast_to(bound, rel.shape)[in_window] + 1e-12).all()
    # pairs outside the window were 2 S cells apart along some axis: distance >= 2 S cells
    d = rb[:, None, :] - rb[None, :, :]
    d = np.abs(np.mod(d + box / 2, box) - box / 2)
    far = ~in_window
    assert far.any() and ((d / cell)[far].max(-1) >= 2 * S - 1e-9).all()
    # and the bound is not vacuous
    assert bound.max() < 8 * rel[in_window].max()
