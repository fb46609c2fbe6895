"""State writer / reader (jax_sph_b200/io_state.py) against the
rules of jax_sph/io_state.py:14-85 and, on the GPU, the asynchronous TrajectoryWriter against
synchronous downloads of the same steps."""

import os
import re

import numpy as np
import pytest

from jax_sph_b200 import io_state


def _cfg(mode="sim", write_every=2, seq=10, write_type=("h5",), path="x"):
    return {"case": {"mode": mode, "name": "tgv", "dim": 2, "dx": 0.02}, "seed": 42,
            "solver": {"name": "SPH", "sequence_length": seq},
            "io": {"write_every": write_every, "write_type": list(write_type), "data_path": path}}


def _state(n=7, dim=2, seed=0):
    rng = np.random.default_rng(seed)
    st = {k: rng.random((n, dim), dtype=np.float32) for k in ("r", "u", "v", "dudt", "dvdt", "nw")}
    st.update({k: rng.random(n, dtype=np.float32) for k in ("rho", "p", "mass", "eta", "T")})
    st["tag"] = rng.integers(0, 4, n).astype(np.int32)
    return st


def test_snapshot_names_follow_the_reference_rules():
    cfg = _cfg(seq=100, write_every=10)
    # io_state.py:47-50: sim mode writes steps >= 0 that are multiples of write_every
    assert io_state.snapshot_name(-1, cfg) is None
    assert io_state.snapshot_name(0, cfg) == "traj_000"  # zfill(len(str(step_max)))
    assert io_state.snapshot_name(5, cfg) is None
    assert io_state.snapshot_name(30, cfg) == "traj_030"
    rlx = _cfg(mode="rlx", seq=100)
    assert io_state.snapshot_name(50, rlx) is None and io_state.snapshot_name(100, rlx) is None
    assert io_state.snapshot_name(99, rlx) == "tgv_2_0.02_42"  # io_state.py:57-59

    class NS:  # attribute-style config (OmegaConf-like)
        def __init__(self, d):
            for k, v in d.items():
                setattr(self, k, NS(v) if isinstance(v, dict) else v)

    assert io_state.snapshot_name(30, NS(cfg)) == "traj_030"


def test_io_setup_directory_rule(tmp_path):
    d = io_state.io_setup(_cfg(path=str(tmp_path)))
    assert re.fullmatch(r"2D_TGV_SPH_42_\d{8}-\d{6}", os.path.basename(d)) and os.path.isdir(d)
    assert os.path.exists(os.path.join(d, "config.yaml"))
    d = io_state.io_setup(_cfg(path=str(tmp_path), write_type=()))
    assert d == str(tmp_path) + "/"  # nothing written: the bare data_path
    d = io_state.io_setup(_cfg(path=str(tmp_path), mode="rlx"))
    assert d == str(tmp_path) + "/"


@pytest.mark.parametrize("dim", [2, 3])
def test_write_state_round_trips(tmp_path, dim):
    cfg = _cfg(write_type=("h5", "vtk"))
    st = _state(dim=dim)
    for step in range(-1, 5):
        io_state.write_state(step, st, str(tmp_path), cfg)
    assert sorted(os.listdir(tmp_path)) == [f"traj_{s:02d}.{e}" for s in (0, 2, 4) for e in ("h5", "vtk")]
    back = io_state.read_h5(str(tmp_path / "traj_02.h5"))
    assert sorted(back) == sorted(st)
    for k in st:
        assert back[k].dtype == st[k].dtype and np.array_equal(back[k], st[k]), k
    with pytest.raises(ValueError):
        io_state.read_h5(str(tmp_path / "traj_02.h5"), array_type="jax")
    vtk = io_state.read_vtk(str(tmp_path / "traj_04.vtk"))
    assert sorted(vtk) == sorted(st)
    for k in st:
        v = vtk[k]
        if st[k].ndim == 2 and dim == 2:  # dict2pyvista pads 2D vectors with a zero column
            assert v.shape == (7, 3) and np.all(v[:, 2] == 0)
            v = v[:, :2]
        assert np.array_equal(v, st[k]), k
    head = open(tmp_path / "traj_04.vtk", "rb").read(80)
    assert head.startswith(b"# vtk DataFile Version 3.0\n") and b"BINARY\nDATASET POLYDATA" in head


def test_snapshots_hold_all_sixteen_reference_keys(tmp_path):
    """An engine without wall normals / heat conduction downloads no `nw`, `kappa`, `Cp`; the
    snapshot still has the reference's sixteen entries (zeros / the case constants), and entries
    that are present pass through untouched."""
    cfg = _cfg(write_type=("h5",))
    full = _state(dim=3)
    n = len(full["r"])
    full.update(drhodt=np.zeros(n, np.float32), dTdt=np.zeros(n, np.float32),
                kappa=np.full(n, 7.0, np.float32), Cp=np.full(n, 3.0, np.float32))
    part = {k: v for k, v in full.items() if k not in ("nw", "kappa", "Cp")}
    io_state.write_state(0, part, str(tmp_path), cfg)
    back = io_state.read_h5(str(tmp_path / "traj_00.h5"))
    assert sorted(back) == sorted(io_state.REFERENCE_KEYS)
    for k in part:
        assert np.array_equal(back[k], part[k]), k
    assert back["nw"].shape == full["r"].shape and not back["nw"].any()
    assert back["kappa"].shape == (len(full["r"]),) and back["Cp"].dtype == np.float32
    assert io_state.complete_state(full) is full


@pytest.mark.gpu
def test_trajectory_writer_equals_synchronous_downloads(tmp_path):
    """The loop of jax_sph/simulate.py:113-134 with the asynchronous writer: every snapshot
    equals the engine state downloaded synchronously at the same step of a second run."""
    import torch

    from jax_sph_b200 import Engine, config_from_setup
    from oracle import cases

    setup = cases.make_case("tgv", dim=2, dx=0.02, dtype=np.float32, tvf=1.0)
    cfg = _cfg(write_every=3, seq=10, write_type=("h5", "vtk"))
    n = len(setup.state["r"])

    eng = Engine(config_from_setup(setup), n)
    eng.upload(setup.state)
    eng.step(0.0, 1)
    writer = io_state.TrajectoryWriter(eng, str(tmp_path), cfg)
    for step in range(cfg["solver"]["sequence_length"] + 2):
        writer.write(step - 1)
        eng.step(setup.dt, 1)
    writer.close()
    assert eng.error() == 0
    assert writer.written == ["traj_00", "traj_03", "traj_06", "traj_09"]

    ref = Engine(config_from_setup(setup), n)
    ref.upload(setup.state)
    ref.step(0.0, 1)
    for step in range(cfg["solver"]["sequence_length"] + 2):
        if (step - 1) >= 0 and (step - 1) % 3 == 0:
            want = {k: v.numpy() for k, v in ref.download(host=True).items()}
            got = io_state.read_h5(str(tmp_path / f"traj_{step - 1:02d}.h5"))
            assert sorted(got) == sorted(io_state.REFERENCE_KEYS)  # nw / kappa / Cp filled in
            for k in want:
                assert np.array_equal(got[k], want[k]), (step, k)
            vtk = io_state.read_vtk(str(tmp_path / f"traj_{step - 1:02d}.vtk"))
            assert np.array_equal(vtk["u"][:, :2], want["u"]) and np.array_equal(vtk["tag"], want["tag"])
        ref.step(setup.dt, 1)
    # resume: a snapshot read back as CUDA tensors goes straight into an engine
    snap = io_state.read_h5(str(tmp_path / "traj_09.h5"), array_type="torch")
    assert all(v.is_cuda for v in snap.values())
    res = Engine(config_from_setup(setup), n)
    res.upload(snap)
    res.step(setup.dt, 1)
    assert res.error() == 0 and torch.isfinite(res.download()["u"]).all()
