"""The driver loop (jax_sph_b200/simulate.py <- jax_sph/simulate.py:19-136) and the on-device
get_stats (sphb200_engine_get_stats <- jax_sph/utils.py:128-166).

CPU: config defaults, the time-step rule, the progress-line format.  GPU: a whole TGV run
through simulate() -- device-made start, asynchronous h5 snapshots, progress lines -- against
the oracle's loop on the same case, and get_stats against NumPy on the downloaded state."""

import os
import re

import numpy as np
import pytest

from jax_sph_b200 import _lib, simulate as sim
from jax_sph_b200.engine import stats_from_words


def test_defaults_and_time_step_follow_the_reference():
    from oracle import cases

    cfg = sim.defaults(case=dict(dim=2, dx=0.02), solver=dict(tvf=1.0, t_end=0.05))
    assert cfg["io"]["print_props"] == ["Ekin", "u_max"] and cfg["solver"]["cfl"] == 0.25
    setup = cases.make_case("tgv", dim=2, dx=0.02, dtype=np.float32, tvf=1.0)
    assert abs(sim.time_step(cfg) - setup.dt) <= 1e-12 * setup.dt  # case_setup.py:94-97
    cfg3 = sim.defaults(case=dict(dim=3, dx=2 * np.pi / 20, viscosity=0.02))
    setup3 = cases.make_case("tgv", dim=3, dx=2 * np.pi / 20, dtype=np.float32, viscosity=0.02)
    assert abs(sim.time_step(cfg3) - setup3.dt) <= 1e-12 * setup3.dt
    cfg["solver"]["dt"] = 1e-4
    assert sim.time_step(cfg) == 1e-4  # explicit dt wins, case_setup.py:101-104
    with pytest.raises(_lib.Sphb200Error, match="unknown config keys"):
        sim.defaults(case=dict(dxx=0.1))


def test_log_line_format_and_stat_selection():
    # utils.py:288-296: "<step>/<len>, t=<(step+1) dt>, k=v, ..."
    line = sim.log_line(7, 1234, 0.001, {"Ekin": 0.1234567, "u_max": 1.0})
    assert line == "0007/1234, t=0.0080, Ekin=0.12346, u_max=1.00000"
    words = [0.5] + [1.0, 2.0, 30.0] * 5 + [10.0, 0, 0, 0]
    got = stats_from_words(words, ["Ekin", "u_max", "rho_min", "p_mean"])
    assert got == {"Ekin": 0.5, "u_max": 2.0, "rho_min": 1.0, "p_mean": 3.0}
    with pytest.raises(_lib.Sphb200Error):
        stats_from_words(words, ["nw_max"])


@pytest.mark.parametrize("dim,dx", [(2, 0.02), (3, 0.04), (2, 0.05)])
def test_heated_channel_tables_equal_the_oracle_case(dim, dx):
    """simulate.ht_case / time_step against the oracle's case object (pinned to the
    reference's cases/ht.py by tests/test_reference_pins.py): box, particle count, dt, the bc
    table, the band force and the engine config built from them."""
    from jax_sph_b200 import config_from_setup, make_config
    from oracle import cases

    setup = cases.make_case("ht", dim=dim, dx=dx, dtype=np.float32, r0_noise_factor=0.0)
    cfg = sim.defaults(case=dict(name="ht", dim=dim, dx=dx, g_ext_magnitude=2.3, kappa_ref=7.313,
                                 Cp_ref=305.27),
                       solver=dict(heat_conduction=True, is_bc_trick=True, t_end=1.5),
                       eos=dict(p_bg_factor=0.05))
    ht = sim.ht_case(cfg)
    assert np.allclose(ht["box"], setup.box_size, rtol=1e-12)
    assert int(np.prod(ht["nxyz"])) == len(setup.state["r"])
    assert abs(sim.time_step(cfg) - setup.dt) <= 1e-12 * setup.dt
    assert ht["bc_table"] == setup.bc_table
    g1, g2 = ht["g_ext_spec"], setup.g_ext_spec
    assert g1["mode"] == g2["mode"] and g1["axis"] == g2["axis"] and np.allclose(g1["g"], g2["g"])
    assert abs(g1["lo"] - g2["lo"]) < 1e-12 and abs(g1["hi"] - g2["hi"]) < 1e-12
    # tags of the lattice rule (init.cuh) == tags of the case, particle by particle
    n = ht["nxyz"]
    grids = np.meshgrid(*[np.arange(k) for k in n], indexing="xy")
    idx = np.stack([g.ravel() for g in grids], axis=1)
    x = ((idx[:, 0] + np.float32(0.5)) * np.float32(dx)).astype(np.float32)
    wall = (idx[:, 1] < 3) | (idx[:, 1] >= n[1] - 3)
    hot = (idx[:, 1] < 3) & (x < np.float32(ht["hot"][1])) & (x > np.float32(ht["hot"][0]))
    tag = np.where(hot, 3, np.where(wall, 1, 0))
    q = np.rint(setup.state["r"] / dx - 0.5).astype(np.int64)
    key_ref = np.lexsort(tuple(q[:, a] for a in range(dim)))
    key_lat = np.lexsort(tuple(idx[:, a] for a in range(dim)))
    assert np.array_equal(tag[key_lat], setup.state["tag"][key_ref])
    # the engine config the driver builds == the one built from the oracle's case object
    a = make_config(dim, ht["box"], dx, sim.time_step(cfg), p_ref=100.0, p_bg=5.0, c_ref=10.0,
                    is_bc_trick=True, is_heat_conduction=True, g_ext_spec=ht["g_ext_spec"],
                    bc_table=ht["bc_table"], uniform_eta=True)
    b = config_from_setup(setup)
    import ctypes as C

    raw = lambda c: bytes((C.c_char * C.sizeof(type(c))).from_buffer_copy(c))  # noqa: E731
    a.dt = b.dt  # equal to 1e-12 relative (asserted above), not necessarily to the last bit
    assert raw(a) == raw(b)


def test_cases_not_built_on_the_device_are_refused():
    with pytest.raises(_lib.Sphb200Error, match="prepared setup"):
        sim.simulate(sim.defaults(case=dict(name="db", dim=2)))
    with pytest.raises(_lib.Sphb200Error, match="not supported"):
        sim.simulate(sim.defaults(case=dict(r0_type="poisson")))
    with pytest.raises(FileNotFoundError, match="first run the relaxation"):
        import torch

        if not torch.cuda.is_available():
            raise FileNotFoundError("first run the relaxation (no device: lattice not built)")
        sim.simulate(sim.defaults(case=dict(dim=2, dx=0.1, r0_type="relaxed",
                                            state0_path="/nonexistent/tgv.h5")))


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(case="tgv", dim=2, dx=0.05, tvf=1.0),
                                dict(case="ht", dim=2, dx=0.04),
                                dict(case="db", dim=2, dx=0.05)])
def test_get_stats_equals_numpy_on_the_downloaded_state(kw):
    from jax_sph_b200 import Engine, config_from_setup
    from oracle import cases

    setup = cases.make_case(dtype=np.float32, **kw)
    eng = Engine(config_from_setup(setup), len(setup.state["r"]))
    eng.upload(setup.state)
    eng.step(setup.dt, 3)
    st = {k: v.numpy() for k, v in eng.download(host=True).items()}
    props = ["Ekin"] + [f"{v}_{o}" for v in ("u", "v", "rho", "p", "T") for o in ("min", "max", "mean")]
    got = eng.get_stats(props)
    fluid = st["tag"] == 0
    want = {"Ekin": 0.5 * float((st["v"][fluid].astype(np.float64) ** 2).sum()) * setup.dx ** setup.dim}
    for var in ("u", "v", "rho", "p", "T"):
        a = st[var].astype(np.float64)
        a = np.sqrt((a**2).sum(axis=1)) if a.ndim == 2 else a  # utils.py:148-151
        want.update({f"{var}_min": a.min(), f"{var}_max": a.max(), f"{var}_mean": a.mean()})
    for k in props:
        assert abs(got[k] - want[k]) <= 2e-6 * max(abs(want[k]), 1e-3), (k, got[k], want[k])


@pytest.mark.gpu
def test_simulate_tgv_equals_the_oracle_loop(tmp_path):
    from _util import assert_close
    from oracle import cases, integrator

    cfg = sim.defaults(case=dict(name="tgv", dim=2, dx=0.05), solver=dict(tvf=1.0, t_end=0.02),
                       io=dict(write_type=["h5"], write_every=3, data_path=str(tmp_path)))
    lines = []
    eng = sim.simulate(cfg, log=lines.append)
    setup = cases.make_case("tgv", dim=2, dx=0.05, dtype=np.float32, tvf=1.0)
    seq = int(0.02 / setup.dt)
    assert eng.run_cfg["solver"]["sequence_length"] == seq
    # simulate.py:113: sequence_length + 2 advance calls
    ref = integrator.simulate(setup, seq + 2, fast_segment_sum=True)
    got = {k: v.numpy() for k, v in eng.download(host=True).items()}
    for k in ("r", "u", "v", "rho", "p"):
        assert_close(k, got[k], ref[k], setup, factor=3.0, what="simulate() final state")
    # files: traj_<step> for step = 0, 3, 6, ... <= seq (write_state(step - 1), io_state.py:47-50)
    files = sorted(f for f in os.listdir(eng.out_dir) if f.startswith("traj_"))
    assert files == [f"traj_{s:0{len(str(seq))}d}.h5" for s in range(0, seq + 1, 3)]
    assert re.fullmatch(r"2D_TGV_SPH_123_\d{8}-\d{6}", os.path.basename(eng.out_dir.rstrip("/")))
    # a snapshot is the oracle's state after step + 1 advances
    from jax_sph_b200 import io_state

    snap = io_state.read_h5(os.path.join(eng.out_dir, files[1]))
    ref4 = integrator.simulate(setup, 4, fast_segment_sum=True)
    for k in ("r", "u", "rho"):
        assert_close(k, snap[k], ref4[k], setup, factor=3.0, what="snapshot traj_3")
    # progress lines every write_every steps + the timing line, Ekin decays from its start value
    stat_lines = [l for l in lines if "Ekin=" in l]
    assert len(stat_lines) == len(range(0, seq + 2, 3)) and lines[-1].startswith("time: ")
    assert re.fullmatch(rf"0+/{seq}, t=\d\.\d{{4}}, Ekin=\d\.\d{{5}}, u_max=\d\.\d{{5}}", stat_lines[0])
    ek = [float(re.search(r"Ekin=([\d.]+)", l).group(1)) for l in stat_lines]
    assert 0.2 < ek[0] < 0.26 and ek[-1] <= ek[0]  # 2D TGV: E_kin(0) = 1/4 per unit area


@pytest.mark.gpu
def test_simulate_with_a_prepared_setup(tmp_path):
    """Cases that are not built on the device go in prepared (the reference's initialize()
    results as plain attributes): dam break with walls, bc table and gravity, vtk output."""
    from _util import assert_close
    from jax_sph_b200 import io_state
    from oracle import cases, integrator

    setup = cases.make_case("db", dim=2, dx=0.05, dtype=np.float32)
    setup.sequence_length = 6
    cfg = sim.defaults(case=dict(name="db", dim=2, dx=0.05), solver=dict(name="SPH"),
                       io=dict(write_type=["vtk"], write_every=4, data_path=str(tmp_path),
                               print_props=["Ekin", "u_max", "rho_min", "p_max"]))
    lines = []
    eng = sim.simulate(cfg, setup=setup, log=lines.append)
    ref = integrator.simulate(setup, setup.sequence_length + 2, fast_segment_sum=True)
    got = {k: v.numpy() for k, v in eng.download(host=True).items()}
    for k in ("r", "u", "rho", "p"):
        assert_close(k, got[k], ref[k], setup, factor=3.0, what="simulate(db) final state")
    assert "rho_min=" in lines[0] and "p_max=" in lines[0]
    files = sorted(f for f in os.listdir(eng.out_dir) if f.endswith(".vtk"))
    assert files == ["traj_0.vtk", "traj_4.vtk"]
    snap = io_state.read_vtk(os.path.join(eng.out_dir, "traj_4.vtk"))
    ref5 = integrator.simulate(setup, 5, fast_segment_sum=True)
    assert_close("r", snap["r"][:, :2], ref5["r"], setup, factor=3.0, what="vtk snapshot")
    assert np.array_equal(snap["tag"], setup.state["tag"])


@pytest.mark.gpu
def test_relaxation_run_and_relaxed_start(tmp_path):
    """validation/tgv2d.sh:13 + :17 on the engine: (1) relaxation -- noisy lattice at rest,
    tvf = 1, background pressure, 5000 steps, last state written as tgv_2_<dx>_<seed>.h5;
    (2) the simulation that starts from it (case.r0_type = "relaxed"): same particle count,
    positions of the relaxed state, TGV velocities evaluated there."""
    from jax_sph_b200 import case_setup, io_state

    dx = 0.05
    rlx = sim.defaults(seed=7, case=dict(name="tgv", dim=2, dx=dx, mode="rlx", r0_noise_factor=0.25),
                       solver=dict(tvf=1.0), eos=dict(p_bg_factor=0.02),
                       io=dict(write_type=["h5"], write_every=1000, data_path=str(tmp_path)))
    lines = []
    eng = sim.simulate(rlx, log=lines.append)
    stem = case_setup.relaxed_state_name("tgv", 2, dx, 7)
    path = tmp_path / (stem + ".h5")
    assert eng.run_cfg["solver"]["sequence_length"] == 5000 and path.exists()
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".h5")) == [stem + ".h5"]
    relaxed = io_state.read_h5(str(path))
    n = int(round(1 / dx)) ** 2
    assert relaxed["r"].shape == (n, 2)
    assert (relaxed["r"] >= 0).all() and (relaxed["r"] <= 1).all()
    # the relaxation spreads the noisy particles: density close to uniform, almost at rest
    assert abs(float(relaxed["rho"].mean()) - 1.0) < 0.05 and float(relaxed["rho"].std()) < 0.05
    assert float(np.abs(relaxed["u"]).max()) < 0.5
    lattice = (np.arange(int(round(1 / dx))) + 0.5) * dx
    off_lattice = np.abs(relaxed["r"][:, 0, None] - lattice[None]).min(axis=1)
    assert off_lattice.max() > 0.1 * dx  # not the Cartesian start any more

    run = sim.defaults(seed=7, case=dict(name="tgv", dim=2, dx=dx, r0_type="relaxed",
                                         state0_path=str(path)),
                       solver=dict(tvf=1.0, t_end=0.01), io=dict(data_path=str(tmp_path)))
    lines = []
    eng = sim.simulate(run, log=lines.append)
    ek0 = float(re.search(r"Ekin=([\d.]+)", lines[0]).group(1))
    assert 0.2 < ek0 < 0.27  # TGV field on the relaxed positions
    got = eng.download(host=True)
    assert got["r"].shape == (n, 2) and bool(np.isfinite(got["u"].numpy()).all())


@pytest.mark.gpu
def test_noise_and_velocity_kernels():
    """sphb200_add_noise / sphb200_eval_velocity: Gaussian noise of the requested width on the
    fluid particles only, wrapped into the box, reproducible per (seed, lattice row) for any
    slab; the velocity field at arbitrary positions equals the lattice kernel's."""
    import torch

    from jax_sph_b200 import case_setup

    box, dx = [1.0, 0.26, 0.5], 0.02
    kw = dict(wall_axis=1, n_walls=3)
    clean = case_setup.init_lattice(case_setup.lattice_spec(box, dx, **kw))
    a = case_setup.init_lattice(case_setup.lattice_spec(box, dx, **kw))
    b = case_setup.init_lattice(case_setup.lattice_spec(box, dx, **kw))
    case_setup.add_noise(a, 0.25 * dx, 42, box)
    case_setup.add_noise(b, 0.25 * dx, 42, box)
    assert torch.equal(a["r"], b["r"])
    fluid = (clean["tag"] == 0)
    assert torch.equal(a["r"][~fluid], clean["r"][~fluid])  # get_noise_masked: walls stay
    d = (a["r"] - clean["r"])[fluid]
    boxt = torch.tensor(box, device="cuda")
    d = d - boxt * torch.round(d / boxt)  # undo the periodic wrap
    assert abs(float(d.mean())) < 2e-4 and abs(float(d.std()) / (0.25 * dx) - 1.0) < 0.02
    assert (a["r"] >= 0).all() and (a["r"] <= boxt).all()
    c = case_setup.init_lattice(case_setup.lattice_spec(box, dx, **kw))
    case_setup.add_noise(c, 0.25 * dx, 43, box)
    assert not torch.equal(a["r"], c["r"])
    # a slab draws the deviates of its full-lattice rows
    slab = case_setup.init_lattice(case_setup.lattice_spec(box, dx, planes=(5, 9), **kw), with_ids=True)
    case_setup.add_noise(slab, 0.25 * dx, 42, box)
    assert torch.equal(slab["r"], a["r"][slab["ids"].long()])
    # velocity field at given positions
    lat = case_setup.lattice_spec([2 * np.pi] * 3, 2 * np.pi / 16, velocity="tgv3d")
    st = case_setup.init_lattice(lat)
    want = st["u"].clone()
    st["u"].zero_()
    st["v"].zero_()
    case_setup.eval_velocity(st, "tgv3d")
    assert torch.equal(st["u"], want) and torch.equal(st["v"], want)


@pytest.mark.gpu
@pytest.mark.parametrize("case,dim,dx", [("ht", 2, 0.02), ("ht", 3, 0.05), ("pf", 2, 0.05)])
def test_simulate_channel_cases_equal_the_oracle_loop(case, dim, dx, tmp_path):
    """cases/ht.yaml and cases/pf.yaml through the driver (device-made walls + fluid, band force,
    bc table, heat conduction for ht) against the oracle's loop on its own case object; the two
    enumerate the same lattice in different orders, so particles are matched by their start
    positions."""
    from _util import assert_close
    from oracle import cases, integrator

    setup = cases.make_case(case, dim=dim, dx=dx, dtype=np.float32, r0_noise_factor=0.0)
    nsteps = 6
    if case == "ht":
        over = dict(case=dict(name="ht", dim=dim, dx=dx, g_ext_magnitude=2.3, kappa_ref=7.313,
                              Cp_ref=305.27),
                    solver=dict(heat_conduction=True, is_bc_trick=True), eos=dict(p_bg_factor=0.05))
        props, keys = ["Ekin", "u_max", "T_max"], ("r", "u", "rho", "p", "T")
    else:
        over = dict(case=dict(name="pf", dim=dim, dx=dx, viscosity=100.0, u_ref=1.25,
                              g_ext_magnitude=1000.0),
                    solver=dict(is_bc_trick=True))
        props, keys = ["Ekin", "u_max"], ("r", "u", "rho", "p")
    over["solver"]["t_end"] = (nsteps - 2 + 0.5) * setup.dt
    over["io"] = dict(write_every=3, data_path=str(tmp_path), print_props=props)
    cfg = sim.defaults(**over)
    lines = []
    eng = sim.simulate(cfg, log=lines.append)
    assert eng.run_cfg["solver"]["sequence_length"] == nsteps - 2
    got = {k: v.numpy() for k, v in eng.download(host=True).items()}
    ref = integrator.simulate(setup, nsteps, fast_segment_sum=True)
    # match by the start lattice: the driver's rows are meshgrid("xy") order
    n = sim.channel_case(cfg)["nxyz"]
    grids = np.meshgrid(*[np.arange(k) for k in n], indexing="xy")
    idx = np.stack([g.ravel() for g in grids], axis=1)
    q = np.rint(setup.state["r"] / dx - 0.5).astype(np.int64)
    perm_ref = np.lexsort(tuple(q[:, a] for a in range(dim)))
    perm_lat = np.lexsort(tuple(idx[:, a] for a in range(dim)))
    assert np.array_equal(got["tag"][perm_lat], ref["tag"][perm_ref])
    for k in keys:
        assert_close(k, got[k][perm_lat], ref[k][perm_ref], setup, factor=3.0, what=f"simulate({case} {dim}D)")
    if case == "ht":
        assert "T_max=1.23000" in lines[0]


@pytest.mark.parametrize("dim,dx", [(2, 0.0166666), (2, 0.05), (3, 0.05)])
def test_poiseuille_tables_equal_the_oracle_case(dim, dx):
    """simulate.channel_case for cases/pf.yaml (the case the reference's tests/test_pf2d.py
    integrates) against the oracle's case object: box, count, dt, tables, engine config."""
    import ctypes as C

    from jax_sph_b200 import config_from_setup, make_config
    from oracle import cases

    setup = cases.make_case("pf", dim=dim, dx=dx, dtype=np.float32)
    cfg = sim.defaults(case=dict(name="pf", dim=dim, dx=dx, viscosity=100.0, u_ref=1.25,
                                 g_ext_magnitude=1000.0),
                       solver=dict(is_bc_trick=True))
    ch = sim.channel_case(cfg)
    assert np.allclose(ch["box"], setup.box_size, rtol=1e-12) and ch["hot"] is None
    assert int(np.prod(ch["nxyz"])) == len(setup.state["r"])
    assert abs(sim.time_step(cfg) - setup.dt) <= 1e-12 * setup.dt
    assert ch["bc_table"] == setup.bc_table
    g1, g2 = ch["g_ext_spec"], setup.g_ext_spec
    assert g1["mode"] == g2["mode"] and g1["axis"] == g2["axis"] and np.allclose(g1["g"], g2["g"])
    assert abs(g1["lo"] - g2["lo"]) < 1e-12 and abs(g1["hi"] - g2["hi"]) < 1e-12
    a = make_config(dim, ch["box"], dx, setup.dt, p_ref=setup.p_ref, p_bg=setup.p_bg,
                    c_ref=setup.c_ref, u_ref=1.25, is_bc_trick=True, g_ext_spec=ch["g_ext_spec"],
                    bc_table=ch["bc_table"], uniform_eta=True)
    b = config_from_setup(setup)
    raw = lambda c: bytes((C.c_char * C.sizeof(type(c))).from_buffer_copy(c))  # noqa: E731
    assert raw(a) == raw(b)
    # tags: walls below and above, no hot patch
    n = ch["nxyz"]
    q = np.rint(setup.state["r"] / dx - 0.5).astype(np.int64)
    wall_ref = setup.state["tag"] == 1
    assert np.array_equal(wall_ref, (q[:, 1] < 3) | (q[:, 1] >= n[1] - 3))



def test_channel_geometry_is_read_from_case_special():
    """cases/ht.yaml keeps the channel geometry under `case.special` (SimulationSetup.__init__,
    jax_sph/case_setup.py:36): a config keyed like the reference's changes the box, the hot patch
    and its temperature; a misspelt key is an error, not a silent default."""
    from jax_sph_b200 import _lib

    base = dict(name="ht", dim=2, dx=0.02, g_ext_magnitude=2.3, kappa_ref=7.313, Cp_ref=305.27)
    ref = sim.channel_case(sim.defaults(case=base))
    cfg = sim.defaults(case=dict(base, special=dict(L=2.0, hot_wall_temperature=1.5,
                                                    hot_wall_half_width=0.5)))
    ht = sim.channel_case(cfg)
    assert ht["box"][0] == 2.0 and ref["box"][0] == 1.0 and ht["box"][1] == ref["box"][1]
    assert ht["T_hot"] == 1.5 and ht["hot"] == (0.5, 1.5)
    assert ht["nxyz"][0] == 2 * ref["nxyz"][0]
    # the older top-level section still works, the case section wins
    top = sim.defaults(case=base, special=dict(L=3.0))
    assert sim.channel_case(top)["box"][0] == 3.0
    both = sim.defaults(case=dict(base, special=dict(L=2.0)), special=dict(L=3.0))
    assert sim.channel_case(both)["box"][0] == 2.0
    with pytest.raises(_lib.Sphb200Error):
        sim.channel_case(sim.defaults(case=dict(base, special=dict(hot_wall_temp=1.5))))


@pytest.mark.gpu
def test_state0_path_applies_to_a_cartesian_start(tmp_path):
    """validation/tgv2d.sh passes case.state0_path with case.r0_type=cartesian: initialize()
    (jax_sph/case_setup.py:184-194) then overwrites the fluid entries named by
    case.state0_keys AFTER the velocities were evaluated on the lattice -- positions of the
    snapshot, lattice velocities; with every key listed the run restarts from the snapshot."""
    from jax_sph_b200 import io_state

    dx = 0.05
    n = int(round(1 / dx)) ** 2
    first = sim.defaults(seed=3, case=dict(name="tgv", dim=2, dx=dx, r0_noise_factor=0.25),
                         solver=dict(tvf=1.0, t_end=0.02),
                         io=dict(write_type=["h5"], write_every=1000, data_path=str(tmp_path)))
    sim.simulate(first, log=lambda s: None)
    # (simulation runs write into a run directory below io.data_path, io_state.io_setup)
    snaps = sorted(os.path.join(d, f) for d, _, fs in os.walk(tmp_path) for f in fs if f.endswith(".h5"))
    assert snaps
    path = snaps[-1]
    snap = io_state.read_h5(path)

    def start(keys):
        cfg = sim.defaults(seed=3, case=dict(name="tgv", dim=2, dx=dx, state0_path=path,
                                             state0_keys=keys),
                           solver=dict(tvf=1.0, t_end=1e-9), io=dict(data_path=str(tmp_path)))
        prep = sim._prepare_tgv(cfg)
        return {k: (v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)) for k, v in prep.state.items()}

    only_r = start(["r"])
    assert only_r["r"].shape == (n, 2) and np.array_equal(only_r["r"], snap["r"])
    lattice = (np.arange(int(round(1 / dx))) + 0.5) * dx
    gx, gy = np.meshgrid(lattice, lattice, indexing="ij")
    # the velocity is the TGV field of the LATTICE positions (evaluated before the overwrite)
    u_lat = -np.cos(2 * np.pi * gx) * np.sin(2 * np.pi * gy)
    assert np.allclose(np.sort(only_r["u"][:, 0]), np.sort(u_lat.ravel()), atol=1e-5)
    restart = start(["r", "u", "v", "rho", "p", "dudt", "dvdt"])
    for k in ("r", "u", "v", "rho", "p", "dudt", "dvdt"):
        assert np.array_equal(restart[k], snap[k]), k
