"""On-device case initialisation for the regular-lattice cases.

Mirrors, with the output resident in HBM (torch CUDA tensors in the reference's layouts):

* ``pos_init_cartesian_2d`` / ``pos_init_cartesian_3d``  <- jax_sph/utils.py:35-54
* the field setup of ``SimulationSetup.initialize()``     <- jax_sph/case_setup.py:127-181
  for lattice starts: TGV velocity fields (cases/tgv.py:37-51), channel walls along one
  axis with the Dirichlet hot patch of cases/ht.py:90-97, uniform rho / p / mass / eta / T /
  kappa / Cp.

The reference builds these with NumPy ``meshgrid`` + ``vstack`` on the host and ships the
whole state to the device; here one write-only kernel (csrc/init.cuh) fills the arrays, also
per slab (``planes``) for the multi-GPU engine, so that a 64 M-particle start costs
milliseconds.  Position noise (``case.r0_noise_factor``) is added on the device as well
(``add_noise``: counter-based Philox deviates keyed by seed and lattice row -- the same
distribution as the reference's jax.random stream, not the same numbers), and the case's
velocity field can be evaluated at any positions (``eval_velocity``: noisy and relaxed starts,
``case.r0_type == "relaxed"`` with the positions read from a state file).
All compute goes through ``sphb200_init_lattice`` / ``sphb200_add_noise`` /
``sphb200_eval_velocity`` (include/sphb200.h); no CPU fallback.
"""

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib

_VEL = {"rest": _lib.VEL_REST, "tgv2d": _lib.VEL_TGV2D, "tgv3d": _lib.VEL_TGV3D}
_KEYS = _lib.VECTOR_FIELDS + _lib.SCALAR_FIELDS + ("tag",)


def lattice_shape(box_size, dx) -> Tuple[int, ...]:
    """Planes per axis, `np.array((box_size / dx).round(), dtype=int)` (utils.py:40,51)."""
    box = np.asarray(box_size, dtype=np.float64).reshape(-1)
    return tuple(int(v) for v in np.array((box / dx).round(), dtype=int))


def lattice_spec(box_size, dx, *, velocity="rest", wall_axis=-1, n_walls=0,
                 hot: Optional[Sequence[float]] = None, T_hot=1.0, rho=1.0, p=0.0, mass=None,
                 eta=0.0, T=1.0, kappa=0.0, Cp=0.0, planes: Optional[Tuple[int, int]] = None):
    """`sphb200_lattice` for a box of `box_size` at spacing `dx`.  `planes` = (k_lo, k_hi)
    restricts the last axis to one rank's slab; `hot` = (x_lo, x_hi) marks the Dirichlet patch
    of the lower wall; `mass` defaults to rho * dx**dim (case_setup.py:100)."""
    n = lattice_shape(box_size, dx)
    dim = len(n)
    if dim not in (2, 3):
        raise _lib.Sphb200Error("box_size must have 2 or 3 entries")
    if velocity not in _VEL:
        raise _lib.Sphb200Error(f"velocity {velocity!r} is not one of {tuple(_VEL)}")
    # a lattice plane at (n - 0.5) dx outside the periodic box would be reported by the engine as
    # SPHB200_ERR_OUTSIDE_BOX on the first step; refuse it here, where the cause is visible
    box = np.asarray(box_size, dtype=np.float64).reshape(-1)
    for a in range(dim):
        if (n[a] - 0.5) * dx >= box[a] * (1 - 1e-6):
            raise _lib.Sphb200Error(
                f"lattice {n} at dx={dx} does not fit the box {box.tolist()} (axis {a})")
    lat = _lib.Lattice()
    lat.struct_size = C.sizeof(_lib.Lattice)
    lat.dim = dim
    for a in range(3):
        lat.n[a] = n[a] if a < dim else 1
    lat.k_lo, lat.k_hi = (0, n[dim - 1]) if planes is None else (int(planes[0]), int(planes[1]))
    lat.velocity = _VEL[velocity]
    lat.wall_axis, lat.n_walls = int(wall_axis), int(n_walls)
    lat.hot_lo, lat.hot_hi = (1.0, 0.0) if hot is None else (float(hot[0]), float(hot[1]))
    lat.T_hot = T_hot
    lat.dx, lat.rho, lat.p = dx, rho, p
    lat.mass = rho * dx**dim if mass is None else mass
    lat.eta, lat.T, lat.kappa, lat.Cp = eta, T, kappa, Cp
    return lat


def lattice_rows(lat) -> int:
    rows = _lib.load().sphb200_lattice_rows(C.byref(lat))
    if rows < 0:
        _lib.check(int(rows))
    return int(rows)


def init_lattice(lat, keys: Sequence[str] = _KEYS, with_ids: bool = False) -> Dict:
    """Fill a fresh state dict (torch CUDA tensors, reference layouts and row order) on the
    current device and stream.  With `with_ids` the dict also carries `ids`, the row of each
    particle in the full lattice (what SlabEngine.upload takes)."""
    import torch

    from .engine import _stream_ptr

    lib = _lib.load()
    rows = lattice_rows(lat)
    if not torch.cuda.is_available():
        raise _lib.Sphb200Error("init_lattice needs a CUDA device (there is no CPU fallback)")
    st = _lib.State()
    out = {}
    for k in keys:
        if k not in _KEYS:
            raise _lib.Sphb200Error(f"unknown state key {k!r}")
        shape = (rows, lat.dim) if k in _lib.VECTOR_FIELDS else (rows,)
        out[k] = torch.empty(shape, dtype=torch.int32 if k == "tag" else torch.float32,
                             device="cuda")
        setattr(st, k, out[k].data_ptr() if rows else None)
    ids = torch.empty(rows, dtype=torch.int32, device="cuda") if with_ids else None
    _lib.check(lib.sphb200_init_lattice(
        C.byref(lat), C.byref(st), C.c_void_p(ids.data_ptr() if (with_ids and rows) else None),
        _stream_ptr()))
    if with_ids:
        out["ids"] = ids
    return out


def pos_init_cartesian_2d(box_size, dx):
    """jax_sph/utils.py:35-46: particles at the centres of the Cartesian grid cells, (N, 2)."""
    if len(np.asarray(box_size).reshape(-1)) != 2:
        raise _lib.Sphb200Error("pos_init_cartesian_2d needs a 2-entry box_size")
    return init_lattice(_bare_spec(box_size, dx), keys=("r",))["r"]


def pos_init_cartesian_3d(box_size, dx):
    """jax_sph/utils.py:49-54, (N, 3)."""
    if len(np.asarray(box_size).reshape(-1)) != 3:
        raise _lib.Sphb200Error("pos_init_cartesian_3d needs a 3-entry box_size")
    return init_lattice(_bare_spec(box_size, dx), keys=("r",))["r"]


def _bare_spec(box_size, dx):
    # the reference's generators take any block size (wall blocks, utils.py:57-118): no box check
    n = lattice_shape(box_size, dx)
    lat = _lib.Lattice()
    lat.struct_size = C.sizeof(_lib.Lattice)
    lat.dim = len(n)
    for a in range(3):
        lat.n[a] = n[a] if a < len(n) else 1
    lat.k_lo, lat.k_hi = 0, n[-1]
    lat.wall_axis = -1
    lat.dx = dx
    return lat


def slab_planes(engine, n_last: int, dx: float) -> Tuple[int, int]:
    """Lattice planes [k_lo, k_hi) of the last axis whose particles a SlabEngine rank owns at
    the start: plane k sits at (k + 0.5) dx (float32, as the kernel writes it) and belongs to
    the rank whose cell layers [z0, z1) contain it (slab.layer_of, the rule select_own applies
    to a global state).  Layers grow with k, so the planes of a rank are one range."""
    from .slab import layer_of

    ax = ((np.arange(n_last, dtype=np.float32) + np.float32(0.5)) * np.float32(dx)).astype(np.float32)
    lay = layer_of(ax, engine.inv_cell, engine.layers)
    mine = np.nonzero((lay >= engine.z0) & (lay < engine.z1))[0]
    if len(mine) == 0:
        return 0, 0
    return int(mine[0]), int(mine[-1]) + 1


def init_slab(engine, box_size, dx, **spec):
    """Device-made start of ONE rank of a slab-decomposed run: the rank's lattice planes are
    generated where they will live (no global state on the host, no host-to-device copy) and
    uploaded with their full-lattice ids.  Returns the number of particles the rank holds."""
    n = lattice_shape(box_size, dx)
    lat = lattice_spec(box_size, dx, planes=slab_planes(engine, n[-1], dx), **spec)
    state = init_lattice(lat, with_ids=True)
    ids = state.pop("ids")
    engine.upload(state, ids)
    return int(ids.numel())


def apply_state0(state: Dict, state0: Dict, keys: Sequence[str] = ("r",)) -> Dict:
    """The restart / relaxed-start rule of SimulationSetup.initialize() (jax_sph/case_setup.py:
    184-194): for every key in `keys`, the FLUID entries of `state` are overwritten by the
    fluid entries of `state0` (a snapshot read with io_state.read_h5), walls keep the freshly
    generated values.  Works on NumPy arrays and on torch tensors (host or CUDA); `state` is
    modified in place and returned."""
    FLUID = 0
    mask, mask0 = state["tag"] == FLUID, state0["tag"] == FLUID
    for k in state:
        if k not in keys:
            continue
        if k not in state0:
            raise ValueError(f"Key {k} not found in state0 file.")
        src = state0[k][mask0]
        if tuple(state[k][mask].shape) != tuple(src.shape):
            raise ValueError(f"Shape mismatch for key {k} in state0 file.")
        if hasattr(state[k], "is_cuda") and not hasattr(src, "is_cuda"):
            import torch

            src = torch.as_tensor(np.ascontiguousarray(src), device=state[k].device)
        state[k][mask] = src
    return state


def add_noise(state: Dict, std: float, seed: int, box_size, ids=None) -> Dict:
    """`r = shift_fn(r, get_noise_masked(r.shape, tag == FLUID, key, std))` of
    SimulationSetup.initialize() (jax_sph/case_setup.py:138-144) on CUDA tensors, in place.
    `ids` (or state["ids"]): full-lattice rows of a slab's particles, so that every
    decomposition draws the same deviates."""
    from .engine import _stream_ptr

    r = state["r"]
    if not r.is_cuda:
        raise _lib.Sphb200Error("add_noise works on CUDA tensors (there is no CPU fallback)")
    n, dim = r.shape
    box = (C.c_double * 3)(*([float(b) for b in np.asarray(box_size, dtype=np.float64).reshape(-1)]
                             + [1.0])[:3])
    tag = state.get("tag")
    ids = ids if ids is not None else state.get("ids")
    _lib.check(_lib.load().sphb200_add_noise(
        dim, n, C.c_void_p(r.data_ptr()), C.c_void_p(tag.data_ptr() if tag is not None else None),
        C.c_void_p(ids.data_ptr() if ids is not None else None), float(std),
        int(seed) & 0xFFFFFFFFFFFFFFFF, C.byref(box), _stream_ptr()))
    return state


def eval_velocity(state: Dict, velocity: str) -> Dict:
    """`u = v = vmap(self._init_velocity{2,3}D)(r)` (case_setup.py:146-150) at the CURRENT
    positions of `state` (CUDA tensors), in place."""
    import torch

    from .engine import _stream_ptr

    if velocity not in _VEL:
        raise _lib.Sphb200Error(f"velocity {velocity!r} is not one of {tuple(_VEL)}")
    r = state["r"]
    if not r.is_cuda:
        raise _lib.Sphb200Error("eval_velocity works on CUDA tensors (there is no CPU fallback)")
    n, dim = r.shape
    for k in ("u", "v"):
        if k not in state:
            state[k] = torch.empty_like(r)
    _lib.check(_lib.load().sphb200_eval_velocity(
        dim, n, _VEL[velocity], C.c_void_p(r.data_ptr()), C.c_void_p(state["u"].data_ptr()),
        C.c_void_p(state["v"].data_ptr()), _stream_ptr()))
    return state


def relaxed_state_name(case_name: str, dim: int, dx: float, seed: int) -> str:
    """File stem of a relaxed start, `_get_relaxed_r0` (case_setup.py:277-280) and the name
    write_state gives the last state of a relaxation run (io_state.py:57-59): tgv_3_0.02_42."""
    return "_".join([str(case_name), str(dim), str(dx), str(seed)])

