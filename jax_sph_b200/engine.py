"""Resident B200 engine: the step loop of jax_sph/simulate.py:110-134 with the
particle state kept cell-sorted in HBM.

`Engine` is a thin, typed wrapper over the C ABI (include/sphb200.h); torch is
used only for device memory and streams.  States are dicts with the reference's
keys (jax_sph/solver.py:930-947): torch CUDA tensors or NumPy arrays.
"""

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib

STATE_KEYS = _lib.VECTOR_FIELDS + _lib.SCALAR_FIELDS + ("tag",)


def make_config(
    dim, box, dx, dt, *, solver="SPH", kernel="QSK", h_fac=1.0, tvf=0.0, eos=None,
    p_ref=None, rho_ref=1.0, p_bg=0.0, gamma=1.0, u_ref=1.0, c_ref=10.0, eta_limiter=3.0,
    is_bc_trick=False, is_rho_evol=False, is_rho_renorm=False, is_free_slip=False,
    is_heat_conduction=False, artificial_alpha=0.0, g_ext_spec=None, bc_table=None,
    cell_sub=None, tile=None, threads=0, list_cap=0, stage_cap=0, nl_cap=0, g_ext_array=False,
    r_cutoff=0.0, wall_layer=None, diff_delta=0.1, diff_alpha=0.01, skin=0.0, uniform_eta=False,
):
    """Build a `sphb200_config` from the WCSPH constructor arguments
    (jax_sph/solver.py:616-637) plus the table forms of the case callables."""
    if solver not in _lib.SOLVER:
        raise _lib.Sphb200Error(f"solver {solver!r} is not supported (SPH, RIE, DELTA)")
    if kernel not in _lib.KERNEL:
        raise _lib.Sphb200Error(f"kernel {kernel!r} is not supported {tuple(_lib.KERNEL)}")
    cfg = _lib.default_config()
    cfg.dim = dim
    cfg.solver = _lib.SOLVER[solver]
    cfg.kernel = _lib.KERNEL[kernel]
    if eos is None:
        eos = "RIEMANN" if solver == "RIE" else "TAIT"
    cfg.eos = _lib.EOS_RIEMANN if eos == "RIEMANN" else _lib.EOS_TAIT
    box = np.asarray(box, dtype=np.float64).reshape(-1)
    if box.size == 1:
        box = np.repeat(box, dim)
    for a in range(dim):
        cfg.box[a] = float(box[a])
    cfg.dx = dx
    cfg.h = h_fac * dx
    # promises about the state (verified on the device every step: SPHB200_ERR_HINT)
    cfg.hints = _lib.HINT_UNIFORM_ETA if uniform_eta else 0
    cfg.dt = dt
    cfg.tvf = tvf
    cfg.c_ref = c_ref
    cfg.eta_limiter = eta_limiter
    cfg.artificial_alpha = artificial_alpha
    cfg.p_ref = p_ref if p_ref is not None else rho_ref * c_ref**2 / gamma
    cfg.rho_ref, cfg.p_bg, cfg.gamma, cfg.u_ref = rho_ref, p_bg, gamma, u_ref
    cfg.flags = ((_lib.F_BC_TRICK if is_bc_trick else 0) | (_lib.F_RHO_EVOL if is_rho_evol else 0)
                 | (_lib.F_RHO_RENORM if is_rho_renorm else 0)
                 | (_lib.F_FREE_SLIP if is_free_slip else 0)
                 | (_lib.F_HEAT if is_heat_conduction else 0))
    spec = g_ext_spec or {"mode": "none"}
    if g_ext_array:
        cfg.g_mode = _lib.G_ARRAY
    elif spec["mode"] == "const":
        cfg.g_mode = _lib.G_CONST
        for a in range(3):
            cfg.g[a] = spec["g"][a]
    elif spec["mode"] == "band":
        cfg.g_mode = _lib.G_BAND
        cfg.g_axis = spec["axis"]
        cfg.g_lo, cfg.g_hi = spec["lo"], spec["hi"]
        for a in range(3):
            cfg.g[a] = spec["g"][a]
    else:
        cfg.g_mode = _lib.G_NONE
    if bc_table:
        for tag, rule in bc_table.get("tags", {}).items():
            r = cfg.bc[int(tag)]
            fl = 0
            if "u" in rule:
                fl |= _lib.BC_SET_U
                for a in range(3):
                    r.u[a] = rule["u"][a]
            if "v" in rule:
                fl |= _lib.BC_SET_V
                for a in range(3):
                    r.v[a] = rule["v"][a]
            if rule.get("zero_dudt"):
                fl |= _lib.BC_ZERO_DUDT
            if rule.get("zero_dvdt"):
                fl |= _lib.BC_ZERO_DVDT
            if "p" in rule:
                fl |= _lib.BC_SET_P
                r.p = rule["p"]
            if "T" in rule:
                fl |= _lib.BC_SET_T
                r.T = rule["T"]
            if rule.get("zero_dTdt"):
                fl |= _lib.BC_ZERO_DTDT
            r.flags = fl
        if bc_table.get("inflow_x"):
            cfg.bc_inflow_on = 1
            cfg.bc_inflow_x = bc_table["inflow_x"]["x"]
            cfg.bc_inflow_T = bc_table["inflow_x"]["T"]
        if bc_table.get("outflow_x"):
            cfg.bc_outflow_on = 1
            cfg.bc_outflow_x = bc_table["outflow_x"]["x"]
    for name, val in (("cell_sub", cell_sub), ("tile", tile)):
        if val is not None:
            arr = getattr(cfg, name)
            for a in range(3):
                arr[a] = int(val[a]) if a < len(val) else 0
    cfg.threads, cfg.list_cap, cfg.stage_cap = threads, list_cap, stage_cap
    cfg.nl_cap = nl_cap
    cfg.r_cutoff = float(r_cutoff)
    cfg.diff_delta, cfg.diff_alpha = float(diff_delta), float(diff_alpha)
    # neighbour-list skin / cutoff: 0 = automatic, < 0 = sort and search every step
    cfg.skin = float(skin)
    # not part of the C struct: Engine / SlabEngine hand it to sphb200_engine_set_wall_layer
    cfg.wall_layer = wall_layer
    return cfg


STAT_VARS = ("u", "v", "rho", "p", "T")


def stats_from_words(words, props):
    """Pick `props` out of the SPHB200_NSTATS words of sphb200_engine_get_stats (summed /
    min-max-reduced over the ranks first on a slab engine)."""
    res = {}
    for prop in props:
        if prop == "Ekin":
            res[prop] = words[0]
            continue
        var, operation = prop.split("_")  # e.g. "u_max", utils.py:164
        if var not in STAT_VARS or operation not in ("min", "max", "mean"):
            raise _lib.Sphb200Error(f"get_stats: {prop!r} is not offered on the device "
                                    f"(Ekin, <{'|'.join(STAT_VARS)}>_<min|max|mean>)")
        k = 1 + 3 * STAT_VARS.index(var)
        res[prop] = {"min": words[k], "max": words[k + 1],
                     "mean": words[k + 2] / max(words[16], 1.0)}[operation]
    return res


def live_fields(cfg):
    """(read, written): the state entries one `advance(dt, state, neighbors)`
    (jax_sph/integrator.py:22-56 + jax_sph/solver.py:705-949) of THIS solver variant reads
    and the ones it changes.  Everything else passes through the reference's advance()
    untouched (`v` is overwritten before it is read, integrator.py:27; `p` is recomputed
    from rho, solver.py:801, and only the Riemann continuity equation reads the incoming
    one, :773-791; `drhodt` is an output of the evolution variants only; T / dTdt / kappa /
    Cp matter with heat conduction only, :832-849; `nw` with free-slip or Riemann walls)."""
    f = cfg.flags
    rie = cfg.solver == _lib.SOLVER["RIE"]
    evol, heat = bool(f & _lib.F_RHO_EVOL), bool(f & _lib.F_HEAT)
    has_nw = rie or bool(f & _lib.F_FREE_SLIP)
    read = ["r", "u", "dudt", "dvdt", "tag", "mass", "eta", "rho"]
    read += ["p"] if (rie and evol) else []
    read += ["T", "dTdt", "kappa", "Cp"] if heat else []
    read += ["nw"] if has_nw else []
    written = ["r", "u", "v", "dudt", "dvdt", "rho", "p"]
    written += ["drhodt"] if evol else []
    written += ["T", "dTdt"] if heat else []
    written += ["nw"] if (has_nw and getattr(cfg, "wall_layer", None)) else []
    return tuple(read), tuple(written)


def config_from_setup(setup, **tuning):
    """`setup` is anything exposing the fields of the reference's
    SimulationSetup / WCSPH call (jax_sph/simulate.py:49-69): used by tests and
    bench with the oracle's case objects -- only plain attributes are read.  The uniform-viscosity
    hint is set when the setup's state says so (every reference case: eta = viscosity)."""
    if "uniform_eta" not in tuning:
        eta = getattr(setup, "state", {}).get("eta") if isinstance(getattr(setup, "state", None), dict) else None
        tuning = dict(tuning, uniform_eta=bool(eta is not None and len(eta) and np.ptp(np.asarray(eta)) == 0.0))
    return make_config(
        setup.dim, setup.box_size, setup.dx, setup.dt, solver=setup.solver, kernel=setup.kernel,
        h_fac=setup.h_factor, tvf=setup.tvf, p_ref=setup.p_ref, rho_ref=setup.rho_ref,
        p_bg=setup.p_bg, gamma=setup.gamma, u_ref=setup.u_ref, c_ref=setup.c_ref,
        eta_limiter=setup.eta_limiter, is_bc_trick=setup.is_bc_trick,
        is_rho_evol=setup.density_evolution, is_rho_renorm=setup.density_renormalize,
        is_free_slip=setup.free_slip, is_heat_conduction=setup.heat_conduction,
        artificial_alpha=setup.artificial_alpha, g_ext_spec=setup.g_ext_spec,
        bc_table=setup.bc_table, wall_layer=getattr(setup, "nw_spec", None),
        diff_delta=getattr(setup, "diff_delta", 0.1), diff_alpha=getattr(setup, "diff_alpha", 0.01),
        **tuning)


def set_wall_layer(lib, handle, dim, layer, offset, cutoff):
    layer = np.ascontiguousarray(np.asarray(layer, dtype=np.float32).reshape(-1, dim))
    offset = np.ascontiguousarray(np.asarray(offset, dtype=np.float32).reshape(dim))
    _lib.check(lib.sphb200_engine_set_wall_layer(
        handle, layer.ctypes.data_as(C.c_void_p), int(len(layer)),
        offset.ctypes.data_as(C.c_void_p), float(cutoff)))


def _torch():
    import torch

    return torch


def _stream_ptr():
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One engine = one particle system resident on the current CUDA device."""

    def __init__(self, cfg, n: int):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.Sphb200Error("no CUDA device: the engine has no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.n = int(n)
        self.dim = int(cfg.dim)
        nbytes = C.c_size_t()
        _lib.check(self.lib.sphb200_engine_bytes(C.byref(cfg), self.n, C.byref(nbytes)))
        # torch owns the arena (caching allocator, freed with the object)
        self._arena = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
        self._h = C.c_void_p()
        _lib.check(self.lib.sphb200_engine_create_in(
            C.byref(cfg), self.n, C.c_void_p(self._arena.data_ptr()), nbytes.value,
            C.byref(self._h)))
        self.arena_bytes = nbytes.value
        self._keep = []
        if getattr(cfg, "wall_layer", None):
            self.set_wall_layer(**cfg.wall_layer)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.sphb200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- marshalling -----------------------------------------------------------
    def _state_struct(self, state: Dict, writable: bool):
        """dict -> (sphb200_state, on_host).  Arrays must be float32/int32 and contiguous."""
        torch = _torch()
        st = _lib.State()
        on_host = None
        keep = []
        for k in STATE_KEYS + ("g_ext",):
            a = state.get(k)
            if a is None:
                continue
            want_int = k == "tag"
            if isinstance(a, np.ndarray):
                host = True
                if a.dtype == np.float64:
                    raise _lib.Sphb200Error(f"state[{k!r}] is float64: float32 only (SPHB200_EDTYPE)")
                want = np.int32 if want_int else np.float32
                if a.dtype != want or not a.flags.c_contiguous:
                    if writable:
                        raise _lib.Sphb200Error(f"output state[{k!r}] must be contiguous {want}")
                    a = np.ascontiguousarray(a, dtype=want)
                ptr = a.ctypes.data
            else:
                host = not a.is_cuda
                if a.dtype == torch.float64:
                    raise _lib.Sphb200Error(f"state[{k!r}] is float64: float32 only (SPHB200_EDTYPE)")
                want = torch.int32 if want_int else torch.float32
                if a.dtype != want or not a.is_contiguous():
                    if writable:
                        raise _lib.Sphb200Error(f"output state[{k!r}] must be contiguous {want}")
                    a = a.to(want).contiguous()
                ptr = a.data_ptr()
            n_expected = self.n * (self.dim if k in _lib.VECTOR_FIELDS + ("g_ext",) else 1)
            if a.size if isinstance(a, np.ndarray) else a.numel():
                pass
            count = a.size if isinstance(a, np.ndarray) else a.numel()
            if count != n_expected:
                raise _lib.Sphb200Error(f"state[{k!r}] has {count} elements, expected {n_expected}")
            if on_host is None:
                on_host = host
            elif on_host != host:
                raise _lib.Sphb200Error("state mixes host and device arrays")
            keep.append(a)
            setattr(st, k, ptr)
        return st, bool(on_host), keep

    # -- API -------------------------------------------------------------------
    def upload(self, state: Dict):
        st, on_host, keep = self._state_struct(state, writable=False)
        self._keep = keep
        _lib.check(self.lib.sphb200_engine_upload(self._h, C.byref(st), int(on_host), _stream_ptr()))

    def refresh(self, state: Dict):
        """Upload a state of the SAME particles (row i is still particle i) into the slots they
        occupy: the cell table and the neighbour lists are kept as far as the new positions allow
        (sphb200_engine_refresh).  Before the first step it is an ordinary upload."""
        st, on_host, keep = self._state_struct(state, writable=False)
        self._keep = keep
        _lib.check(self.lib.sphb200_engine_refresh(self._h, C.byref(st), int(on_host), _stream_ptr()))

    def step(self, dt: float, nsteps: int = 1, integrate: bool = True, bc: bool = True):
        flags = (_lib.STEP_INTEGRATE if integrate else 0) | (_lib.STEP_BC if bc else 0)
        _lib.check(self.lib.sphb200_engine_step(self._h, float(dt), int(nsteps), flags, _stream_ptr()))

    def download(self, out: Optional[Dict] = None, keys=None, host: bool = False):
        """State in the original particle order.  `out` may hold preallocated arrays."""
        torch = _torch()
        if out is None:
            out = {}
            keys = keys or STATE_KEYS
            for k in keys:
                if k == "nw" and not (self.cfg.solver == 1 or self.cfg.flags & _lib.F_FREE_SLIP):
                    continue
                if k in ("kappa", "Cp") and not self.cfg.flags & _lib.F_HEAT:
                    continue
                shape = (self.n, self.dim) if k in _lib.VECTOR_FIELDS else (self.n,)
                dt = torch.int32 if k == "tag" else torch.float32
                if host:
                    out[k] = torch.empty(shape, dtype=dt, pin_memory=True)
                else:
                    out[k] = torch.empty(shape, dtype=dt, device="cuda")
        st, on_host, keep = self._state_struct(out, writable=True)
        _lib.check(self.lib.sphb200_engine_download(self._h, C.byref(st), int(on_host), _stream_ptr()))
        if on_host:
            torch.cuda.current_stream().synchronize()
        return out

    def live_fields(self):
        """(read, written) state entries of one advance() of this solver variant: see
        `live_fields` (module level)."""
        return live_fields(self.cfg)

    def advance_host(self, dt: float, state: Dict, out: Optional[Dict] = None):
        """`advance(dt, state, neighbors)` on HOST buffers, the call a reference user makes with
        host-resident state: copies host->device the entries this variant reads, runs one
        step, copies device->host the entries it writes (into `out` when given, else into
        `state`'s own arrays) and returns the full state dict -- the untouched entries are the
        caller's arrays, exactly what the reference's advance() returns for them."""
        read, written = self.live_fields()
        dst = out if out is not None else state
        st_in, host_in, keep_in = self._state_struct({k: state[k] for k in read}, writable=False)
        st_out, host_out, keep_out = self._state_struct({k: dst[k] for k in written}, writable=True)
        if host_in and host_out:
            # one fused call: r (and u, v when nothing rewrites them after the reorder pass) go
            # back to the host while the sweeps run (sphb200_engine_advance_host)
            self._keep = (keep_in, keep_out)
            _lib.check(self.lib.sphb200_engine_advance_host(
                self._h, float(dt), C.byref(st_in), C.byref(st_out),
                _lib.STEP_INTEGRATE | _lib.STEP_BC, _stream_ptr()))
            _torch().cuda.current_stream().synchronize()
        else:
            self.upload({k: state[k] for k in read})
            self.step(dt, 1)
            self.download(out={k: dst[k] for k in written})
        res = dict(state)
        res.update({k: dst[k] for k in written})
        return res

    def vjp(self, dt: float, cot: Dict, integrate: bool = True) -> Dict:
        """Vector-Jacobian product of the step this engine ran LAST (`step(dt, 1)`, or
        `step(0, 1, integrate=False)` for forward() alone) at the state it holds: `cot` maps
        output names (r, u, v, dudt, rho, p) to cotangent CUDA tensors in the caller's particle
        order, the result maps the step's inputs (r, u, v, dudt, dvdt, rho, p) to theirs
        (sphb200_engine_vjp, csrc/adjoint.cuh).  The counterpart of jax.vjp on `advance`
        (jax_sph/integrator.py:22-56); SPH with summation density and tvf = 0."""
        torch = _torch()
        names = ("r", "u", "v", "dudt", "dvdt", "rho", "p")
        cin = {k: cot[k].to(device="cuda", dtype=torch.float32).contiguous()
               for k in names if k in cot and cot[k] is not None}
        out = {k: torch.zeros((self.n, self.dim) if k in _lib.VECTOR_FIELDS else (self.n,),
                              dtype=torch.float32, device="cuda") for k in names}
        st_in, _, keep_in = self._state_struct(cin, writable=False)
        st_out, _, keep_out = self._state_struct(out, writable=True)
        self._keep = (keep_in, keep_out, cin)
        flags = _lib.STEP_INTEGRATE if integrate else 0
        _lib.check(self.lib.sphb200_engine_vjp(self._h, float(dt), flags, C.byref(st_in),
                                               C.byref(st_out), _stream_ptr()))
        return out

    def error(self) -> int:
        code = C.c_uint32()
        _lib.check(self.lib.sphb200_engine_error(self._h, C.byref(code), _stream_ptr()))
        return code.value

    def neighbor_list(self, capacity: int, mask_self: bool = False):
        """(idx[2, capacity] int32 cuda tensor, edge count); capacity 0 = count only."""
        torch = _torch()
        idx = torch.empty((2, capacity), dtype=torch.int32, device="cuda") if capacity else None
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        _lib.check(self.lib.sphb200_engine_neighbor_list(
            self._h, C.c_void_p(idx.data_ptr() if capacity else None), int(capacity),
            int(mask_self), C.c_void_p(cnt.data_ptr()), _stream_ptr()))
        return idx, int(cnt.item())

    def set_wall_layer(self, layer, offset, cutoff):
        """The integrator's `nw_fn` (jax_sph/integrator.py:33-34, utils.py:197-277): from now on
        every integrating step recomputes the wall normals from this one-layer discretisation
        of the wall surface (`layer` (n, dim), `offset` (dim,), list cutoff)."""
        set_wall_layer(self.lib, self._h, self.dim, layer, offset, cutoff)

    def stats(self):
        ek, um = C.c_double(), C.c_double()
        _lib.check(self.lib.sphb200_engine_stats(self._h, C.byref(ek), C.byref(um), _stream_ptr()))
        return ek.value, um.value

    def get_stats(self, props=("Ekin", "u_max")):
        """`get_stats(state, props, dx)` of jax_sph/utils.py:156-166 on the resident state:
        "Ekin" (get_ekin: transport velocity, fluid particles, times dx**dim) or
        "<var>_<min|max|mean>" for var in u, v, rho, p, T (Euclidean norm for vectors)."""
        out = (C.c_double * 20)()
        _lib.check(self.lib.sphb200_engine_get_stats(self._h, C.byref(out), _stream_ptr()))
        return stats_from_words(list(out), props)

    def launches(self) -> int:
        return int(self.lib.sphb200_engine_launches(self._h))

    def profile(self, on: bool = True):
        _lib.check(self.lib.sphb200_engine_profile(self._h, int(on)))

    def last_times(self):
        ms = (C.c_float * 8)()
        _lib.check(self.lib.sphb200_engine_last_times(self._h, C.byref(ms)))
        return dict(cells=ms[1], density=ms[2], wall=ms[3], force=ms[4], total=ms[5])

    def counters(self):
        """steps run, searches (cell sort + candidate walk) among them, list row length, skin /
        cutoff, tiles, tiles swept without lists, whether the duo sweeps (csrc/sweep2.cuh: two
        particles per thread on one union list) serve this solver variant, and `pairs`: the
        directed pairs (self pairs included) in the exact neighbour lists of the last step,
        counted on the device from the lists' membership bits (duo engines whose every tile has
        lists, else -1) -- what a brute-force search at the step's positions finds
        (sphb200_engine_counters)."""
        out = (C.c_int64 * 8)()
        _lib.check(self.lib.sphb200_engine_counters(self._h, C.byref(out), _stream_ptr()))
        v = list(out)
        return dict(steps=v[0], searches=v[1], list_rows=v[2], skin=v[3] * 1e-6, tiles=v[4],
                    tiles_without_lists=v[5], duo=bool(v[6]), pairs=v[7])

    def plan(self):
        out = (C.c_int32 * 16)()
        _lib.check(self.lib.sphb200_engine_plan(self._h, C.byref(out)))
        v = list(out)
        return dict(cells=v[0:3], sub=v[3:6], tile=v[6:9], threads=v[9], list_cap=v[10],
                    stage_cap=dict(density=v[11], wall=v[12], force=v[13]), exact_all=v[14],
                    ncells=v[15])


def grad_through_steps(cfg, state: Dict, dt: float, nsteps: int, loss_cotangent):
    """Gradient of a scalar function of the state after `nsteps` x advance(dt) with respect to
    the initial state -- what `jax.grad` of the loop in notebooks/iclr24_grads.ipynb (cell 5)
    computes -- by checkpointing every step's input state and calling `Engine.vjp` in reverse.

    loss_cotangent(final_state) -> {name: d loss / d final_state[name]} (CUDA tensors).
    Returns (final_state, {name: d loss / d state[name]}) for r, u, v, dudt, dvdt."""
    torch = _torch()
    n = len(state["r"])
    eng = Engine(cfg, n)
    keys = ("r", "u", "v", "dudt", "dvdt", "rho", "p")
    cur = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() if not hasattr(v, "is_cuda") else v
           for k, v in state.items()}
    saved = []
    for _ in range(nsteps):
        saved.append({k: v.clone() for k, v in cur.items()})
        eng.upload(cur)
        eng.step(dt, 1)
        new = eng.download()
        cur = dict(cur, **new)
    final = cur
    cot = {k: v for k, v in loss_cotangent(final).items() if k in keys}
    for s in reversed(saved):
        eng.upload(s)
        eng.step(dt, 1)  # the linearisation point of this step
        cot = eng.vjp(dt, cot)
    eng.close()
    return final, cot
