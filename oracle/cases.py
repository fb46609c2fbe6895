"""Case setups (oracle; test infrastructure only).

Restates what ``SimulationSetup.initialize`` (jax_sph/case_setup.py:44-231) does
for the cases the benchmark configs name, without OmegaConf / jax:

* lattices                 jax_sph/utils.py:35-118 (pos_init_cartesian_2d/3d, pos_box_2d)
* dt from CFL              jax_sph/case_setup.py:94-110
* EoS selection            jax_sph/case_setup.py:121-124
* field initialisation     jax_sph/case_setup.py:299-306, :164-181
* wall normals (scipy)     jax_sph/utils.py:169-194, case_setup.py:157-161
* TGV                      cases/tgv.py:25-57
* dam break                cases/db.py:47-142, cases/db.yaml
* Poiseuille               cases/pf.py:32-147, cases/pf.yaml
* Couette                  cases/cf.py:39-160, cases/cf.yaml
* heated channel           cases/ht.py:29-187, cases/ht.yaml

Noise uses ``numpy.random.default_rng(seed)`` because ``jax.random`` cannot be
reproduced without jax (documented deviation; inputs are synthetic anyway).
Every case also exports the table form of its ``bc_fn`` / ``g_ext_fn``
(``bc_table``, ``g_ext_spec``) that the CUDA engine consumes, so that tests can
check table == callable.
"""

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import numpy as np

from . import space
from .eos import RIEMANNEoS, TaitEoS
from .solver import DIRICHLET_WALL, FLUID, MOVING_WALL, SOLID_WALL, WALL_TAGS


def pos_init_cartesian_2d(box_size, dx, dtype):
    """utils.py:35-46."""
    n = np.array((np.asarray(box_size) / dx).round(), dtype=int)
    grid = np.meshgrid(range(n[0]), range(n[1]), indexing="xy")
    t = np.dtype(dtype).type
    r = (np.vstack([g.ravel() for g in grid]).T.astype(dtype) + t(0.5)) * t(dx)
    return r.astype(dtype)


def pos_init_cartesian_3d(box_size, dx, dtype):
    """utils.py:49-54."""
    n = np.array((np.asarray(box_size) / dx).round(), dtype=int)
    grid = np.meshgrid(range(n[0]), range(n[1]), range(n[2]), indexing="xy")
    t = np.dtype(dtype).type
    r = (np.vstack([g.ravel() for g in grid]).T.astype(dtype) + t(0.5)) * t(dx)
    return r.astype(dtype)


def pos_box_2d(fluid_box, dx, n_walls, dtype):
    """utils.py:57-78."""
    dxn = n_walls * dx
    vertical = pos_init_cartesian_2d(np.array([dxn, fluid_box[1] + 2 * dxn]), dx, dtype)
    horiz = pos_init_cartesian_2d(np.array([fluid_box[0], dxn]), dx, dtype)
    wall_l = vertical.copy()
    wall_b = horiz.copy() + np.array([dxn, 0.0])
    wall_r = vertical.copy() + np.array([fluid_box[0] + dxn, 0.0])
    wall_t = horiz.copy() + np.array([dxn, fluid_box[1] + dxn])
    return np.concatenate([wall_l, wall_b, wall_r, wall_t]).astype(dtype)


@dataclass
class Setup:
    """Everything ``simulate`` needs (the 10-tuple of case_setup.py:220-231 + solver cfg)."""

    name: str
    dim: int
    dx: float
    dt: float
    dtype: np.dtype
    box_size: np.ndarray
    state: Dict[str, np.ndarray]
    eos: object
    c_ref: float
    p_ref: float
    p_bg: float
    g_ext_fn: Callable
    bc_fn: Callable
    displacement_fn: Callable
    shift_fn: Callable
    solver: str = "SPH"
    kernel: str = "QSK"
    h_factor: float = 1.0
    tvf: float = 0.0
    eta_limiter: float = 3
    is_bc_trick: bool = False
    density_evolution: bool = False
    density_renormalize: bool = False
    artificial_alpha: float = 0.0
    free_slip: bool = False
    heat_conduction: bool = False
    viscosity: float = 0.01
    u_ref: float = 1.0
    rho_ref: float = 1.0
    gamma: float = 1.0
    g_ext_magnitude: float = 0.0
    n_walls: int = 3
    bc_table: dict = field(default_factory=dict)
    g_ext_spec: dict = field(default_factory=dict)
    nw_fn: Optional[Callable] = None  # case_setup.py:218
    diff_delta: float = 0.1  # defaults.py:97
    diff_alpha: float = 0.01  # defaults.py:99
    nw_spec: Optional[dict] = None


def _dt_cfl(dx, u_ref, c_ref, viscosity, g_mag, rho_ref, cfl, eps):
    """case_setup.py:94-97."""
    h = dx
    dt_convective = cfl * h / (c_ref + u_ref)
    dt_viscous = cfl * h**2 * rho_ref / (viscosity + eps)
    dt_body_force = cfl * (h / (g_mag + eps)) ** 0.5
    return float(np.amin([dt_convective, dt_viscous, dt_body_force]))


def _compute_nws_scipy(r, tag, dx, n_walls, offset_vec, wall_part_fn, eps):
    """utils.py:169-194."""
    from scipy.spatial import KDTree

    dx_fac = 5
    is_w = np.isin(tag, WALL_TAGS)
    r_walls = r[is_w].astype(np.float64)
    r_aligned = r_walls - offset_vec
    layer = wall_part_fn(dx / dx_fac, 1).astype(np.float64) - offset_vec / n_walls / dx_fac
    tree = KDTree(layer)
    dist, match_idx = tree.query(r_aligned, k=1)
    dr = layer[match_idx] - r_aligned
    nw_walls = dr / (dist[:, None] + eps)
    nw = np.zeros_like(r)
    nw[is_w] = nw_walls.astype(r.dtype)
    return nw


def make_nw_fn(state0, dx, n_walls, offset_vec, box_size, wall_part_fn, dtype):
    """compute_nws_jax_wrapper (utils.py:197-277): wall normals recomputed every step when a
    MOVING_WALL exists (case_setup.py:209-218, integrator.py:33-34).  For every wall particle
    the closest particle of a 5x finer one-layer discretisation of the wall surface, among
    those within dx * n_walls * sqrt(2) * 1.01 (the Dense neighbour list of :227-238), gives
    nw = disp(layer, wall) / (dist + EPS) in the position dtype.  Reference quirk kept:
    ``mask_to_layer = idx > len(r_walls)`` (:252) never matches the FIRST layer particle."""
    dtype = np.dtype(dtype)
    t = dtype.type
    eps = t(np.finfo(dtype).eps)
    tag = state0["tag"]
    is_w = np.isin(tag, WALL_TAGS)
    off = np.asarray(offset_vec).astype(dtype)
    layer = (wall_part_fn(dx / 5, 1).astype(dtype) - (np.asarray(offset_vec) / n_walls / 5).astype(dtype))
    cutoff = dx * n_walls * 2.0**0.5 * 1.01
    c2 = t(t(cutoff) * t(cutoff))  # weak-typed scalar squared in the position dtype
    side = np.asarray(box_size).astype(dtype)

    def nw_fn(r):
        r_walls = (r[is_w] - off).astype(dtype)
        dr = space.periodic_displacement(side, (layer[None, :, :] - r_walls[:, None, :]).astype(dtype))
        d2 = space.square_distance(dr)
        dist = space.distance(dr)
        ok = d2 < c2
        ok[:, 0] = False  # :252, idx > len(r_walls)
        dist = np.where(ok, dist, np.inf).astype(dtype)
        k = np.argmin(dist, axis=1)
        rows = np.arange(len(r_walls))
        with np.errstate(invalid="ignore"):
            nw_walls = dr[rows, k] / (dist[rows, k] + eps)[:, None]
        nw = np.zeros_like(r)
        nw[is_w] = nw_walls.astype(dtype)
        return nw

    return nw_fn


def make_case(
    case: str,
    dim: int = 2,
    dx: float = 0.02,
    dtype=np.float32,
    solver: str = "SPH",
    tvf: float = 0.0,
    kernel: str = "QSK",
    h_factor: float = 1.0,
    density_evolution: Optional[bool] = None,
    density_renormalize: bool = False,
    is_bc_trick: Optional[bool] = None,
    artificial_alpha: Optional[float] = None,
    free_slip: bool = False,
    heat_conduction: Optional[bool] = None,
    viscosity: Optional[float] = None,
    u_ref: Optional[float] = None,
    g_ext_magnitude: Optional[float] = None,
    dt: Optional[float] = None,
    r0_noise_factor: Optional[float] = None,
    p_bg_factor: Optional[float] = None,
    eta_limiter: float = 3,
    seed: int = 123,
    n_walls: int = 3,
    cfl: float = 0.25,
    special: Optional[dict] = None,
    kappa_ref: Optional[float] = None,
    Cp_ref: Optional[float] = None,
    T_ref: float = 1.0,
    box_override=None,
    gamma: float = 1.0,
    diff_delta: float = 0.1,
    diff_alpha: float = 0.01,
) -> Setup:
    """Build a Setup the way ``SimulationSetup.initialize`` does for ``cases/<case>.yaml``."""
    dtype = np.dtype(dtype)
    t = dtype.type
    case = case.lower()
    # defaults.py:6-148 merged with cases/<case>.yaml
    yaml = {
        "tgv": dict(),
        "db": dict(g_ext_magnitude=1.0, u_ref=2.0**0.5, viscosity=0.00005, is_bc_trick=True,
                   density_evolution=True, artificial_alpha=0.1,
                   special=dict(L_wall=5.366, H_wall=2.0, L=2.0, H=1.0, W=0.2, box_offset=0.1)),
        "pf": dict(viscosity=100.0, u_ref=1.25, g_ext_magnitude=1000.0, is_bc_trick=True,
                   special=dict(L=0.4, H=1.0)),
        "cf": dict(viscosity=100.0, u_ref=1.25, is_bc_trick=True,
                   special=dict(L=0.4, H=1.0, u_x_wall=1.25)),
        "ht": dict(r0_noise_factor=0.05, g_ext_magnitude=2.3, kappa_ref=7.313, Cp_ref=305.27,
                   heat_conduction=True, is_bc_trick=True, p_bg_factor=0.05,
                   special=dict(hot_wall_temperature=1.23, hot_wall_half_width=0.25, L=1.0, H=0.2)),
    }[case]

    def pick(val, key, default):
        return val if val is not None else yaml.get(key, default)

    viscosity = pick(viscosity, "viscosity", 0.01)
    u_ref = pick(u_ref, "u_ref", 1.0)
    g_mag = pick(g_ext_magnitude, "g_ext_magnitude", 0.0)
    is_bc_trick = pick(is_bc_trick, "is_bc_trick", False)
    density_evolution = pick(density_evolution, "density_evolution", False)
    artificial_alpha = pick(artificial_alpha, "artificial_alpha", 0.0)
    heat_conduction = pick(heat_conduction, "heat_conduction", False)
    r0_noise_factor = pick(r0_noise_factor, "r0_noise_factor", 0.0)
    p_bg_factor = pick(p_bg_factor, "p_bg_factor", 0.0)
    kappa_ref = pick(kappa_ref, "kappa_ref", 0.0)
    Cp_ref = pick(Cp_ref, "Cp_ref", 0.0)
    sp = dict(yaml.get("special", {}))
    sp.update(special or {})
    rho_ref, c_ref_factor = 1.0, 10.0  # defaults.py:119-127 (gamma: eos.gamma)
    eps = float(np.finfo(dtype).eps)

    c_ref = c_ref_factor * u_ref  # case_setup.py:83
    p_ref = rho_ref * c_ref**2 / gamma
    p_bg = p_bg_factor * p_ref
    mass_ref = dx**dim * rho_ref
    dt_cfl = _dt_cfl(dx, u_ref, c_ref, viscosity, g_mag, rho_ref, cfl, eps)
    dt = dt_cfl if dt is None else dt
    eos = RIEMANNEoS(rho_ref, p_bg, u_ref) if solver == "RIE" else TaitEoS(p_ref, rho_ref, p_bg, gamma)

    init2 = lambda b, d=dx: pos_init_cartesian_2d(b, d, dtype)  # noqa: E731
    init3 = lambda b, d=dx: pos_init_cartesian_3d(b, d, dtype)  # noqa: E731
    dxn = dx * n_walls
    offset_vec = np.zeros(dim)
    wall_part_fn = None
    u_wall = None

    if case == "tgv":
        box = np.array([1.0, 1.0]) if dim == 2 else 2 * np.pi * np.array([1.0, 1.0, 1.0])
        if box_override is not None:
            box = np.asarray(box_override, dtype=np.float64)
        r = init2(box) if dim == 2 else init3(box)
        tag = np.full(len(r), FLUID, dtype=np.int32)
    elif case == "db":
        assert dim == 2, "3D dam break is marked 'not validated' upstream (db.py:92)"
        box = np.array([sp["L_wall"] + 2 * dxn + sp["box_offset"],
                        sp["H_wall"] + 2 * dxn + sp["box_offset"]])
        r_f = (t(n_walls * dx) + init2(np.array([sp["L"], sp["H"]]))).astype(dtype)
        wall_part_fn = lambda d, nwl: pos_box_2d(  # noqa: E731
            np.array([sp["L_wall"], sp["H_wall"]]), d, nwl, dtype)
        r_w = wall_part_fn(dx, n_walls)
        r = np.concatenate([r_w, r_f]).astype(dtype)
        tag = np.concatenate([np.full(len(r_w), SOLID_WALL), np.full(len(r_f), FLUID)]).astype(np.int32)
        offset_vec = np.ones(dim) * dxn
    elif case in ("pf", "cf", "ht"):
        depth = 0.5 if case == "ht" else 0.4
        if dim == 2:
            box = np.array([sp["L"], sp["H"] + 2 * dxn])
            ey = np.array([0.0, 1.0])
            r_f = (ey * dxn + init2(np.array([sp["L"], sp["H"]]))).astype(dtype)

            def wall_part_fn(d, nwl):
                horiz = pos_init_cartesian_2d(np.array([sp["L"], d * nwl]), d, dtype)
                return np.concatenate([horiz, horiz + np.array([0.0, sp["H"] + d * nwl])]).astype(dtype)
        else:
            box = np.array([sp["L"], sp["H"] + 2 * dxn, depth])
            ey = np.array([0.0, 1.0, 0.0])
            r_f = (ey * dxn + init3(np.array([sp["L"], sp["H"], depth]))).astype(dtype)

            def wall_part_fn(d, nwl):
                horiz = pos_init_cartesian_3d(np.array([sp["L"], d * nwl, depth]), d, dtype)
                return np.concatenate(
                    [horiz, horiz + np.array([0.0, sp["H"] + d * nwl, 0.0])]).astype(dtype)

        r_w = wall_part_fn(dx, n_walls)
        r = np.concatenate([r_w, r_f]).astype(dtype)
        tag = np.concatenate([np.full(len(r_w), SOLID_WALL), np.full(len(r_f), FLUID)]).astype(np.int32)
        offset_vec = ey * dxn
        if case == "cf":  # cf.py:100-103
            tag = np.where(r[:, 1] > (box[1] - n_walls * dx), MOVING_WALL, tag).astype(np.int32)
            u_wall = np.array([sp["u_x_wall"]] + [0.0] * (dim - 1))
        if case == "ht":  # ht.py:90-97
            hw = sp["hot_wall_half_width"]
            mask_hot = (r[:, 1] < dx * n_walls) & (r[:, 0] < box[0] / 2 + hw) & (r[:, 0] > box[0] / 2 - hw)
            tag = np.where(mask_hot, DIRICHLET_WALL, tag).astype(np.int32)
    else:
        raise ValueError(case)

    displacement_fn, shift_fn = space.periodic(side=box)
    N = len(r)

    if r0_noise_factor != 0.0:  # case_setup.py:139-144
        rng = np.random.default_rng(seed)
        noise = (r0_noise_factor * dx) * rng.standard_normal(r.shape)
        noise = np.where((tag == FLUID)[:, None], noise, 0.0).astype(dtype)
        r = shift_fn(r, noise)

    # initial velocity
    if case == "tgv" and dim == 2:  # tgv.py:37-43
        x, y = r[:, 0], r[:, 1]
        two_pi = t(2.0 * np.pi)
        vel = np.stack([-np.cos(two_pi * x) * np.sin(two_pi * y),
                        np.sin(two_pi * x) * np.cos(two_pi * y)], axis=1).astype(dtype)
    elif case == "tgv":  # tgv.py:45-51
        x, y, z = r[:, 0], r[:, 1], r[:, 2]
        vel = np.stack([np.sin(x) * np.cos(y) * np.cos(z),
                        -np.cos(x) * np.sin(y) * np.cos(z), np.zeros_like(x)], axis=1).astype(dtype)
    else:
        vel = np.zeros_like(r)

    ones = np.ones(N, dtype=dtype)
    rho = ones * t(rho_ref)
    is_nw = (free_slip or solver == "RIE") and is_bc_trick
    if is_nw:
        nw = _compute_nws_scipy(r, tag, dx, n_walls, offset_vec, wall_part_fn, eps)
    else:
        nw = np.zeros_like(r)

    state = {
        "r": r, "tag": tag, "u": vel.copy(), "v": vel.copy(),
        "dudt": np.zeros_like(vel), "dvdt": np.zeros_like(vel),
        "drhodt": np.zeros_like(rho), "rho": rho, "p": eos.p_fn(rho),
        "mass": ones * t(mass_ref), "eta": ones * t(viscosity),
        "dTdt": np.zeros_like(rho), "T": ones * t(T_ref),
        "kappa": ones * t(kappa_ref), "Cp": ones * t(Cp_ref), "nw": nw,
    }

    # --- g_ext_fn / bc_fn per case, plus their table forms --------------------
    zero_vec = [0.0, 0.0, 0.0]
    bc_table = {"tags": {}, "inflow_x": None, "outflow_x": None}
    g_ext_spec = {"mode": "none"}

    if case == "tgv":
        g_ext_fn = lambda r_: np.zeros_like(r_)  # noqa: E731
        bc_fn = lambda s: s  # noqa: E731
    elif case == "db":
        def g_ext_fn(r_):
            res = np.zeros_like(r_)
            res[:, 1] = -g_mag
            return res

        g_ext_spec = {"mode": "const", "g": [0.0, -g_mag, 0.0]}

        def bc_fn(s):
            s = dict(s)
            m = s["tag"] == SOLID_WALL
            for k in ("u", "v", "dudt", "dvdt"):
                s[k] = np.where(m[:, None], t(0.0), s[k])
            s["p"] = np.where(m, t(0.0), s["p"])
            return s

        bc_table["tags"][SOLID_WALL] = dict(u=zero_vec, v=zero_vec, zero_dudt=True, zero_dvdt=True, p=0.0)
    elif case in ("pf", "ht"):
        box2 = np.array([sp["L"], sp["H"] + 2 * dxn])  # both cases use _box_size2D here

        def g_ext_fn(r_):
            res = np.zeros_like(r_)
            fm = (r_[:, 1] < box2[1] - dxn) * (r_[:, 1] > dxn)
            res[:, 0] = np.where(fm, 1.0, 0.0)
            return (res * g_mag).astype(r_.dtype)

        g_ext_spec = {"mode": "band", "g": [g_mag, 0.0, 0.0], "axis": 1,
                      "lo": float(dxn), "hi": float(box2[1] - dxn)}
        if case == "pf":
            def bc_fn(s):
                s = dict(s)
                m = (s["tag"] == SOLID_WALL)[:, None]
                for k in ("u", "v", "dudt", "dvdt"):
                    s[k] = np.where(m, t(0), s[k])
                return s

            bc_table["tags"][SOLID_WALL] = dict(u=zero_vec, v=zero_vec, zero_dudt=True, zero_dvdt=True)
        else:
            T_hot = sp["hot_wall_temperature"]
            x_out = float(box[0]) - n_walls * dx  # bounds[0][1] - n_walls*dx, ht.py:182-184

            def bc_fn(s):
                s = dict(s)
                mask_fluid = s["tag"] == FLUID
                mask_inflow = mask_fluid * (s["r"][:, 0] < n_walls * dx)
                s["T"] = np.where(mask_inflow, t(T_ref), s["T"])
                s["dTdt"] = np.where(mask_inflow, t(0.0), s["dTdt"])
                mask_hot = s["tag"] == DIRICHLET_WALL
                s["T"] = np.where(mask_hot, t(T_hot), s["T"])
                s["dTdt"] = np.where(mask_hot, t(0.0), s["dTdt"])
                mask_solid = s["tag"] == SOLID_WALL
                s["T"] = np.where(mask_solid, t(T_ref), s["T"])
                s["dTdt"] = np.where(mask_solid, t(0), s["dTdt"])
                ms = (mask_hot + mask_solid)[:, None]
                for k in ("u", "v", "dudt", "dvdt"):
                    s[k] = np.where(ms, t(0.0), s[k])
                mask_outflow = mask_fluid * (s["r"][:, 0] > x_out)
                s["dTdt"] = np.where(mask_outflow, t(0.0), s["dTdt"])
                return s

            st = dict(u=zero_vec, v=zero_vec, zero_dudt=True, zero_dvdt=True, zero_dTdt=True)
            bc_table["tags"][SOLID_WALL] = dict(st, T=T_ref)
            bc_table["tags"][DIRICHLET_WALL] = dict(st, T=T_hot)
            bc_table["inflow_x"] = dict(x=float(n_walls * dx), T=T_ref)
            bc_table["outflow_x"] = dict(x=x_out)
    elif case == "cf":
        g_ext_fn = lambda r_: np.zeros_like(r_)  # noqa: E731
        uw = u_wall.astype(dtype)

        def bc_fn(s):
            s = dict(s)
            m1 = (s["tag"] == SOLID_WALL)[:, None]
            m2 = (s["tag"] == MOVING_WALL)[:, None]
            for k in ("u", "v"):
                s[k] = np.where(m1, t(0.0), s[k])
                s[k] = np.where(m2, uw, s[k])
            for k in ("dudt", "dvdt"):
                s[k] = np.where(m1, t(0.0), s[k])
                s[k] = np.where(m2, t(0.0), s[k])
            return s

        uw3 = list(u_wall) + [0.0] * (3 - dim)
        bc_table["tags"][SOLID_WALL] = dict(u=zero_vec, v=zero_vec, zero_dudt=True, zero_dvdt=True)
        bc_table["tags"][MOVING_WALL] = dict(u=uw3, v=uw3, zero_dudt=True, zero_dvdt=True)

    state = bc_fn(state)  # case_setup.py:202

    # case_setup.py:209-218: recompute the wall normals every step when a wall moves
    nw_fn, nw_spec = None, None
    if is_nw and (tag == MOVING_WALL).any():
        nw_fn = make_nw_fn(state, dx, n_walls, offset_vec, box, wall_part_fn, dtype)
        # the same three inputs in the form the CUDA engine takes (sphb200_engine_set_wall_layer)
        nw_spec = dict(
            layer=(wall_part_fn(dx / 5, 1).astype(dtype)
                   - (np.asarray(offset_vec) / n_walls / 5).astype(dtype)).astype(np.float32),
            offset=np.asarray(offset_vec, dtype=np.float32),
            cutoff=dx * n_walls * 2.0**0.5 * 1.01)

    return Setup(
        nw_fn=nw_fn, nw_spec=nw_spec, diff_delta=diff_delta, diff_alpha=diff_alpha,
        name=case, dim=dim, dx=dx, dt=dt, dtype=dtype, box_size=box, state=state, eos=eos,
        c_ref=c_ref, p_ref=p_ref, p_bg=p_bg, g_ext_fn=g_ext_fn, bc_fn=bc_fn,
        displacement_fn=displacement_fn, shift_fn=shift_fn, solver=solver, kernel=kernel,
        h_factor=h_factor, tvf=tvf, eta_limiter=eta_limiter, is_bc_trick=is_bc_trick,
        density_evolution=density_evolution, density_renormalize=density_renormalize,
        artificial_alpha=artificial_alpha, free_slip=free_slip, heat_conduction=heat_conduction,
        viscosity=viscosity, u_ref=u_ref, rho_ref=rho_ref, gamma=gamma, g_ext_magnitude=g_mag,
        n_walls=n_walls, bc_table=bc_table, g_ext_spec=g_ext_spec,
    )
