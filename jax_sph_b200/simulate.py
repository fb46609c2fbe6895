"""The driver loop around the hot path: a mirror of ``simulate(cfg)`` (jax_sph/simulate.py:19-136)
with the state resident in a B200 engine.

    from jax_sph_b200.simulate import defaults, simulate
    cfg = defaults(case=dict(name="tgv", dim=3, dx=6.2832 / 64, viscosity=0.02),
                   solver=dict(tvf=1.0, t_end=1.0), io=dict(write_type=["h5"], write_every=100))
    engine = simulate(cfg)

What is kept: the config keys (jax_sph/defaults.py), the time-step rule (case_setup.py:94-112),
the order of the loop -- ``write_state(step - 1)``, ``advance(dt)``, overflow check, progress
line every ``write_every`` steps (simulate.py:113-134, utils.py:278-296) -- the file names and
the printed line format.  What differs, by construction: there is no edge list to re-allocate
(the per-step ``did_buffer_overflow`` test becomes the engine's device error word, read at
the logging cadence instead of every step, so the loop has no per-step host sync), and the
reference's untimed ``advance(0.0, ...)`` compile call (simulate.py:108-109, result discarded)
has nothing to compile.

Cases.  ``cfg.case.name == "tgv"`` is built on the device (case_setup.lattice_spec /
init_lattice / add_noise / eval_velocity): Cartesian starts with or without position noise, the
relaxation run of the reference's validation scripts (``case.mode="rlx"``: noisy lattice at
rest, 5000 steps, last state written as ``tgv_<dim>_<dx>_<seed>.h5``, validation/tgv3d.sh:19)
and the relaxed start that reads it back (``case.r0_type="relaxed"``); so is the heated
channel ``cfg.case.name == "ht"`` (cases/ht.py, 2D and 3D: BASELINE configs[4]) and the
Poiseuille flow ``"pf"`` (cases/pf.py) on one regular lattice of walls and fluid.  Every other case is passed in prepared:
``simulate(cfg, setup=obj)`` where ``obj`` carries the reference's ``initialize()`` results as
plain attributes (``state``, ``box_size``, ``dt``, ... -- exactly what ``config_from_setup``
reads); the case classes themselves (cases/*.py) are outside the hot-path scope.
"""

import copy
import time
from typing import Dict, Optional

import numpy as np

from . import _lib, case_setup, io_state
from .engine import Engine, config_from_setup, make_config

EPS = float(np.finfo(np.float32).eps)  # jnp.finfo(float).eps with x64 off (case_setup.py:26)


def defaults(**overrides) -> Dict:
    """The entries of jax_sph/defaults.py:11-148 the path reads, as a plain nested dict;
    ``overrides`` are merged per section (``defaults(case=dict(dx=0.02))``)."""
    cfg = {
        "seed": 123, "dtype": "float32",
        "case": dict(name="tgv", mode="sim", dim=3, dx=0.05, r0_type="cartesian", state0_path=None,
                     state0_keys=["r"], special=dict(), r0_noise_factor=0.0, g_ext_magnitude=0.0, viscosity=0.01, u_ref=1.0,
                     c_ref_factor=10.0, rho_ref=1.0, T_ref=1.0, kappa_ref=0.0, Cp_ref=0.0),
        "solver": dict(name="SPH", tvf=0.0, cfl=0.25, density_evolution=False,
                       density_renormalize=False, dt=None, t_end=0.2, artificial_alpha=0.0,
                       free_slip=False, eta_limiter=3, n_walls=3, heat_conduction=False,
                       is_bc_trick=False, diff_delta=0.1, diff_alpha=0.01),
        "kernel": dict(name="QSK", h_factor=1.0),
        "eos": dict(name="Tait", gamma=1.0, p_bg_factor=0.0),
        "io": dict(write_type=[], write_every=1, data_path="./", print_props=["Ekin", "u_max"]),
        # `case.special` of cases/ht.yaml / pf.yaml: per-case defaults apply where a key is absent
        "special": dict(),
    }
    for section, vals in overrides.items():
        if isinstance(vals, dict):
            unknown = set(vals) - set(cfg.get(section, vals))
            if section in cfg and unknown and section != "special":
                raise _lib.Sphb200Error(f"unknown config keys in {section}: {sorted(unknown)}")
            cfg.setdefault(section, {}).update(vals)
        else:
            cfg[section] = vals
    return cfg


def time_step(cfg) -> float:
    """case_setup.py:94-112: CFL minimum of the convective, viscous and body-force limits,
    overridden by an explicit ``solver.dt``."""
    g = io_state._get
    h = g(cfg, "case.dx")
    u_ref, rho_ref = g(cfg, "case.u_ref"), g(cfg, "case.rho_ref")
    c_ref = g(cfg, "case.c_ref_factor") * u_ref
    cfl = g(cfg, "solver.cfl")
    dt_convective = cfl * h / (c_ref + u_ref)
    dt_viscous = cfl * h**2 * rho_ref / (g(cfg, "case.viscosity") + EPS)
    dt_body_force = cfl * (h / (g(cfg, "case.g_ext_magnitude") + EPS)) ** 0.5
    dt = float(np.amin([dt_convective, dt_viscous, dt_body_force]))
    explicit = g(cfg, "solver.dt")
    return float(explicit) if explicit is not None else dt


class _Prepared:
    """What the loop needs of a case: engine config, state, dt, sequence length."""

    def __init__(self, engine_cfg, state, dt, sequence_length, dx, n):
        self.engine_cfg, self.state, self.dt = engine_cfg, state, dt
        self.sequence_length, self.dx, self.n = sequence_length, dx, n


def _prepare_tgv(cfg) -> _Prepared:
    g = io_state._get
    r0_type, mode = g(cfg, "case.r0_type"), g(cfg, "case.mode")
    if r0_type not in ("cartesian", "relaxed") or mode not in ("sim", "rlx"):
        raise _lib.Sphb200Error(f"case.r0_type {r0_type!r} / case.mode {mode!r} not supported")
    if str(g(cfg, "dtype")) != "float32":
        raise _lib.Sphb200Error("the engine is float32 only (cfg.dtype)")
    dim, dx = g(cfg, "case.dim"), g(cfg, "case.dx")
    box = [1.0, 1.0] if dim == 2 else [2 * np.pi] * 3  # cases/tgv.py:25-29
    rho_ref, u_ref = g(cfg, "case.rho_ref"), g(cfg, "case.u_ref")
    c_ref = g(cfg, "case.c_ref_factor") * u_ref
    gamma = g(cfg, "eos.gamma")
    p_ref = rho_ref * c_ref**2 / gamma       # case_setup.py:82
    p_bg = g(cfg, "eos.p_bg_factor") * p_ref  # :84
    dt = time_step(cfg)
    seq = 5000 if g(cfg, "case.mode") == "rlx" else int(g(cfg, "solver.t_end") / dt)  # :106-112
    name = g(cfg, "solver.name")
    field = "tgv2d" if dim == 2 else "tgv3d"
    if mode == "rlx":
        field = "rest"  # set_relaxation: zero velocity, no external force (case_setup.py:371-389)
    lat = case_setup.lattice_spec(
        box, dx, velocity=field, rho=rho_ref, p=p_bg,
        eta=g(cfg, "case.viscosity"), T=g(cfg, "case.T_ref"), kappa=g(cfg, "case.kappa_ref"),
        Cp=g(cfg, "case.Cp_ref"))
    ecfg = make_config(
        dim, box, dx, dt, solver=name, kernel=g(cfg, "kernel.name"),
        h_fac=g(cfg, "kernel.h_factor"), tvf=g(cfg, "solver.tvf"), p_ref=p_ref, rho_ref=rho_ref,
        p_bg=p_bg, gamma=gamma, u_ref=u_ref, c_ref=c_ref, eta_limiter=g(cfg, "solver.eta_limiter"),
        is_bc_trick=g(cfg, "solver.is_bc_trick"), is_rho_evol=g(cfg, "solver.density_evolution"),
        is_rho_renorm=g(cfg, "solver.density_renormalize"), is_free_slip=g(cfg, "solver.free_slip"),
        is_heat_conduction=g(cfg, "solver.heat_conduction"),
        artificial_alpha=g(cfg, "solver.artificial_alpha"),
        diff_delta=g(cfg, "solver.diff_delta"), diff_alpha=g(cfg, "solver.diff_alpha"),
        uniform_eta=_uniform_eta(cfg))
    state = case_setup.init_lattice(lat)
    # SimulationSetup.initialize() (jax_sph/case_setup.py:126-194), in its order: positions (the
    # lattice, or for r0_type == "relaxed" those of data_relaxed/<name>.h5, cases/tgv.py:20-23 +
    # case_setup.py:274-294), noise on the fluid particles, the velocity field AT those positions,
    # and last -- for ANY r0_type -- the fluid entries named by case.state0_keys taken from the
    # snapshot case.state0_path (:184-194; restarts list every key).
    import os

    path0 = g(cfg, "case.state0_path") if _has(cfg, "case.state0_path") else None
    keys0 = list(g(cfg, "case.state0_keys")) if _has(cfg, "case.state0_keys") else ["r"]
    moved = False
    if r0_type == "relaxed" and mode == "sim":
        if path0 is None or "r" not in keys0:  # case_setup.py:38-39
            raise _lib.Sphb200Error("case.r0_type='relaxed' needs case.state0_path and 'r' in "
                                    "case.state0_keys")
        stem = case_setup.relaxed_state_name(g(cfg, "case.name"), dim, dx, g(cfg, "seed"))
        path = os.path.join("data_relaxed", stem + ".h5")
        if not os.path.isfile(path):
            path = path0  # (a driver without the data_relaxed/ tree: the snapshot itself)
        if not os.path.isfile(path):
            raise FileNotFoundError(
                f"{path}: first run the relaxation (case.mode='rlx', solver.tvf=1, "
                "case.r0_noise_factor=0.25, io.write_type=['h5'], io.data_path='data_relaxed/')")
        snap = io_state.read_h5(path, array_type="torch")
        if tuple(snap["r"].shape) != tuple(state["r"].shape):
            raise _lib.Sphb200Error(f"{path}: {tuple(snap['r'].shape)} positions, the case has "
                                    f"{tuple(state['r'].shape)}")
        state["r"] = snap["r"].contiguous()
        moved = True
    if g(cfg, "case.r0_noise_factor") != 0.0:
        case_setup.add_noise(state, g(cfg, "case.r0_noise_factor") * dx, g(cfg, "seed"), box)
        moved = True
    if moved:
        case_setup.eval_velocity(state, field)
    if path0 is not None:
        if not os.path.isfile(path0):
            raise FileNotFoundError(f"case.state0_path: {path0}")
        case_setup.apply_state0(state, io_state.read_h5(path0, array_type="torch"), keys0)
    return _Prepared(ecfg, state, dt, seq, dx, case_setup.lattice_rows(lat))


_CHANNEL_SPECIAL = {
    # cases/ht.yaml / cases/pf.yaml `case.special`, and the depth of the 3D box (ht.py:37, pf.py)
    "ht": (dict(hot_wall_temperature=1.23, hot_wall_half_width=0.25, L=1.0, H=0.2), 0.5),
    "pf": (dict(L=0.4, H=1.0), 0.4),
}


def channel_case(cfg) -> Dict:
    """The two channel cases with walls below and above and a periodic stream-wise axis --
    heated channel (cases/ht.py:29-187, ht.yaml) and Poiseuille flow (cases/pf.py, pf.yaml) --
    in table form: box, lattice counts, the band force `_external_acceleration_fn`
    (ht.py:139-149) and `_boundary_conditions_fn` (ht.py:151-187; pf: walls at rest).  Walls
    and fluid sit on ONE regular lattice (i + 0.5) dx: where H / dx is not an integer the
    reference leaves a sub-dx gap under the top wall, and it enumerates walls before fluid --
    the particles are the same, their order in the arrays is the lattice's."""
    g = io_state._get
    name = str(g(cfg, "case.name")).lower()
    special, depth = _CHANNEL_SPECIAL[name]
    dim, dx, n_walls = g(cfg, "case.dim"), g(cfg, "case.dx"), g(cfg, "solver.n_walls")
    # `cfg.case.special` as in the reference's YAML files (cases/ht.yaml, pf.yaml; read by
    # SimulationSetup.__init__, case_setup.py:36); a top-level `special` section is accepted too
    sp = {}
    for k, v in special.items():
        if _has(cfg, "case.special." + k):
            sp[k] = g(cfg, "case.special." + k)
        elif _has(cfg, "special." + k):
            sp[k] = g(cfg, "special." + k)
        else:
            sp[k] = v
    for section in ("case.special", "special"):
        if _has(cfg, section):
            given = g(cfg, section)
            unknown = [k for k in (given.keys() if hasattr(given, "keys") else []) if k not in special]
            if unknown:
                raise _lib.Sphb200Error(f"{section}: unknown key(s) {unknown} for case '{name}' "
                                        f"(known: {sorted(special)})")
    box = [sp["L"], sp["H"] + 2 * n_walls * dx] + ([depth] if dim == 3 else [])
    nxyz = [int(round(sp["L"] / dx)), int(round(sp["H"] / dx)) + 2 * n_walls] + (
        [int(round(depth / dx))] if dim == 3 else [])
    zero = [0.0, 0.0, 0.0]
    g_ext_spec = {"mode": "band", "g": [g(cfg, "case.g_ext_magnitude"), 0.0, 0.0], "axis": 1,
                  "lo": float(n_walls * dx), "hi": float(box[1] - n_walls * dx)}
    if name == "pf":
        bc_table = {"tags": {1: dict(u=zero, v=zero, zero_dudt=True, zero_dvdt=True)},
                    "inflow_x": None, "outflow_x": None}
        return dict(box=box, nxyz=nxyz, bc_table=bc_table, g_ext_spec=g_ext_spec, hot=None,
                    T_hot=g(cfg, "case.T_ref"), n_walls=n_walls)
    T_ref = g(cfg, "case.T_ref")
    st = dict(u=zero, v=zero, zero_dudt=True, zero_dvdt=True, zero_dTdt=True)
    bc_table = {"tags": {1: dict(st, T=T_ref), 3: dict(st, T=sp["hot_wall_temperature"])},
                "inflow_x": dict(x=float(n_walls * dx), T=T_ref),
                "outflow_x": dict(x=float(box[0]) - n_walls * dx)}
    hot = (box[0] / 2 - sp["hot_wall_half_width"], box[0] / 2 + sp["hot_wall_half_width"])
    return dict(box=box, nxyz=nxyz, bc_table=bc_table, g_ext_spec=g_ext_spec, hot=hot,
                T_hot=sp["hot_wall_temperature"], n_walls=n_walls)


ht_case = channel_case  # the heated channel was the first of the two


def _prepare_channel(cfg) -> _Prepared:
    g = io_state._get
    if g(cfg, "case.r0_type") != "cartesian" or g(cfg, "case.mode") != "sim":
        raise _lib.Sphb200Error("the channel cases are built from the Cartesian lattice in "
                                "simulation mode; pass other starts as a prepared setup")
    if str(g(cfg, "dtype")) != "float32":
        raise _lib.Sphb200Error("the engine is float32 only (cfg.dtype)")
    dim, dx = g(cfg, "case.dim"), g(cfg, "case.dx")
    ht = channel_case(cfg)
    rho_ref, u_ref = g(cfg, "case.rho_ref"), g(cfg, "case.u_ref")
    c_ref = g(cfg, "case.c_ref_factor") * u_ref
    gamma = g(cfg, "eos.gamma")
    p_ref = rho_ref * c_ref**2 / gamma
    p_bg = g(cfg, "eos.p_bg_factor") * p_ref
    dt = time_step(cfg)
    seq = int(g(cfg, "solver.t_end") / dt)
    lat = case_setup.lattice_spec(
        ht["box"], dx, wall_axis=1, n_walls=ht["n_walls"], hot=ht["hot"], T_hot=ht["T_hot"],
        rho=rho_ref, p=p_bg, eta=g(cfg, "case.viscosity"), T=g(cfg, "case.T_ref"),
        kappa=g(cfg, "case.kappa_ref"), Cp=g(cfg, "case.Cp_ref"))
    if [lat.n[a] for a in range(dim)] != ht["nxyz"]:
        raise _lib.Sphb200Error(f"lattice {list(lat.n)[:dim]} != walls + fluid {ht['nxyz']}: pick dx "
                                "with H / dx, L / dx (and 0.5 / dx) away from half-integers")
    ecfg = make_config(
        dim, ht["box"], dx, dt, solver=g(cfg, "solver.name"), kernel=g(cfg, "kernel.name"),
        h_fac=g(cfg, "kernel.h_factor"), tvf=g(cfg, "solver.tvf"), p_ref=p_ref, rho_ref=rho_ref,
        p_bg=p_bg, gamma=gamma, u_ref=u_ref, c_ref=c_ref, eta_limiter=g(cfg, "solver.eta_limiter"),
        is_bc_trick=g(cfg, "solver.is_bc_trick"), is_rho_evol=g(cfg, "solver.density_evolution"),
        is_rho_renorm=g(cfg, "solver.density_renormalize"), is_free_slip=g(cfg, "solver.free_slip"),
        is_heat_conduction=g(cfg, "solver.heat_conduction"),
        artificial_alpha=g(cfg, "solver.artificial_alpha"),
        diff_delta=g(cfg, "solver.diff_delta"), diff_alpha=g(cfg, "solver.diff_alpha"),
        uniform_eta=_uniform_eta(cfg),
        g_ext_spec=ht["g_ext_spec"], bc_table=ht["bc_table"])
    state = case_setup.init_lattice(lat)
    if g(cfg, "case.r0_noise_factor") != 0.0:  # case_setup.py:138-144 (velocities stay zero)
        case_setup.add_noise(state, g(cfg, "case.r0_noise_factor") * dx, g(cfg, "seed"), ht["box"])
    return _Prepared(ecfg, state, dt, seq, dx, case_setup.lattice_rows(lat))


def _has(cfg, path) -> bool:
    try:
        return io_state._get(cfg, path) is not None
    except (KeyError, AttributeError):
        return False


def _prepare_setup(cfg, setup, tuning) -> _Prepared:
    dt = float(setup.dt)
    seq = getattr(setup, "sequence_length", None)
    if seq is None:
        seq = int(io_state._get(cfg, "solver.t_end") / dt)
    return _Prepared(config_from_setup(setup, **tuning), setup.state, dt, int(seq), setup.dx,
                     len(setup.state["r"]))


def _uniform_eta(cfg) -> bool:
    """The state the driver builds carries eta = case.viscosity for every particle
    (case_setup.py:152-181) unless a restart file overwrites it."""
    keys = io_state._get(cfg, "case.state0_keys") if _has(cfg, "case.state0_keys") else ()
    path = io_state._get(cfg, "case.state0_path") if _has(cfg, "case.state0_path") else None
    return not (path and "eta" in (keys or ()))


def log_line(step: int, sequence_length: int, dt: float, stats: Dict) -> str:
    """Logger.print_stats (utils.py:288-296)."""
    digits = len(str(sequence_length))
    stats_str = ", ".join(f"{k}={v:.5f}" for k, v in stats.items())
    return f"{str(step).zfill(digits)}/{sequence_length}, t={(step + 1) * dt:.4f}, {stats_str}"


def simulate(cfg, setup=None, out_dir: Optional[str] = None, log=print, **tuning):
    """Run the loop of jax_sph/simulate.py:95-136 on a resident engine and return it (state in
    HBM, ``engine.download()`` for the final fields).  ``cfg``: a nested dict / namespace with
    the reference's keys (see ``defaults``); it is not modified -- the derived entries the
    reference writes back (solver.dt, solver.sequence_length, case.c_ref ..., case_setup.py:
    196-199) are on the returned engine as ``engine.run_cfg``."""
    if isinstance(cfg, dict):
        cfg = copy.deepcopy(cfg)
    g = io_state._get
    if setup is not None:
        prep = _prepare_setup(cfg, setup, tuning)
    elif str(g(cfg, "case.name")).lower() == "tgv":
        prep = _prepare_tgv(cfg)
    elif str(g(cfg, "case.name")).lower() in _CHANNEL_SPECIAL:
        prep = _prepare_channel(cfg)
    else:
        raise _lib.Sphb200Error(
            f"case {g(cfg, 'case.name')!r} is not built on the device: pass its initialize() "
            "results as a prepared setup")
    if isinstance(cfg, dict):
        cfg["solver"]["dt"], cfg["solver"]["sequence_length"] = prep.dt, prep.sequence_length
        cfg["case"]["num_particles_max"] = prep.n
    write_cfg = cfg if isinstance(cfg, dict) else {
        "case": dict(mode=g(cfg, "case.mode"), name=g(cfg, "case.name"), dim=g(cfg, "case.dim"),
                     dx=g(cfg, "case.dx")),
        "seed": g(cfg, "seed"), "solver": dict(name=g(cfg, "solver.name"),
                                               sequence_length=prep.sequence_length),
        "io": dict(write_every=g(cfg, "io.write_every"), write_type=list(g(cfg, "io.write_type")),
                   data_path=g(cfg, "io.data_path"))}
    write_every = g(cfg, "io.write_every")
    props = list(g(cfg, "io.print_props"))
    directory = out_dir if out_dir is not None else io_state.io_setup(write_cfg)

    engine = Engine(prep.engine_cfg, prep.n)
    engine.upload(prep.state)
    engine.run_cfg = write_cfg
    writer = io_state.TrajectoryWriter(engine, directory, write_cfg) if g(cfg, "io.write_type") else None

    start = time.time()
    for step in range(prep.sequence_length + 2):  # simulate.py:113
        if writer is not None:
            writer.write(step - 1)
        engine.step(prep.dt, 1)
        if step % write_every == 0:
            # simulate.py:120-131: the overflow check (no list to re-allocate: any device error
            # is fatal), then the progress line (:133-134)
            err = engine.error()
            if err:
                if writer is not None:
                    writer.close()
                raise _lib.Sphb200Error(f"device error word {err:#x} at step {step}")
            if log is not None:
                log(log_line(step, prep.sequence_length, prep.dt, engine.get_stats(props)))
    if writer is not None:
        writer.close()
    err = engine.error()
    if err:
        raise _lib.Sphb200Error(f"device error word {err:#x} at the end of the run")
    if log is not None:
        log(f"time: {time.time() - start:.2f} s")
    engine.out_dir = directory
    return engine
