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


def test_cases_not_built_on_the_device_are_refused():
    with pytest.raises(_lib.Sphb200Error, match="prepared setup"):
        sim.simulate(sim.defaults(case=dict(name="db", dim=2)))
    with pytest.raises(_lib.Sphb200Error, match="without noise"):
        sim.simulate(sim.defaults(case=dict(r0_noise_factor=0.25)))


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

