"""WCSPH solver, host mirror of jax_sph/solver.py:613-951.

Same constructor arguments as the reference (solver.py:616-637).  ``forward``
runs the CUDA sweeps through the stateless C-ABI entry point
``sphb200_forward``; the ``neighbors`` argument is accepted for signature
compatibility and ignored -- the engine keeps particles cell-sorted and sweeps
neighbour cells without a list (its neighbour set is bit-identical to
``neighbors.idx``, tests/test_gpu_neighbors.py).
"""

import ctypes as C
from typing import Callable, Dict

import numpy as np

from . import _lib
from .engine import STATE_KEYS, Engine, make_config


class _KernelInfo:
    """The only kernel attribute callers read: ``_kernel_fn.cutoff`` (simulate.py:78)."""

    def __init__(self, name, h):
        if name not in _lib.KERNEL:
            raise _lib.Sphb200Error(f"kernel {name!r} is not supported {tuple(_lib.KERNEL)}")
        self.h = h
        self.cutoff = (3.0 if name in ("QSK", "GK", "SGK") else 2.0) * h  # kernel.py: cutoff


class WCSPH:
    def __init__(
        self, displacement_fn: Callable, eos, g_ext_fn: Callable, dx: float, dim: int, dt: float,
        c_ref: float, eta_limiter: float = 3, diff_delta=0.02, diff_alpha=0.1, solver: str = "SPH",
        kernel: str = "QSK", h_fac: float = 1.0, is_bc_trick: bool = False,
        is_rho_evol: bool = False, artificial_alpha: float = 0.0, is_free_slip: bool = False,
        is_rho_renorm: bool = False, is_heat_conduction: bool = False,
        g_ext_spec=None, bc_table=None, tvf: float = 0.0,
    ):
        side = getattr(displacement_fn, "side", None)
        if side is None:
            raise _lib.Sphb200Error("displacement_fn must come from jax_sph_b200.space.periodic")
        self.dim, self.dx, self.dt = dim, dx, dt
        self.g_ext_fn = g_ext_fn
        self._kernel_fn = _KernelInfo(kernel, h_fac * dx)
        is_rie = hasattr(eos, "u_ref")
        self._cfg_kwargs = dict(
            solver=solver, kernel=kernel, h_fac=h_fac, tvf=tvf,
            eos="RIEMANN" if is_rie else "TAIT",
            p_ref=getattr(eos, "p_ref", None), rho_ref=eos.rho_ref, p_bg=eos.p_bg,
            gamma=getattr(eos, "gamma", 1.0), u_ref=getattr(eos, "u_ref", 1.0), c_ref=c_ref,
            eta_limiter=eta_limiter, is_bc_trick=is_bc_trick, is_rho_evol=is_rho_evol,
            is_rho_renorm=is_rho_renorm, is_free_slip=is_free_slip,
            is_heat_conduction=is_heat_conduction, artificial_alpha=artificial_alpha,
            bc_table=bc_table, diff_delta=diff_delta, diff_alpha=diff_alpha)
        self._box = np.asarray(side, dtype=np.float64)
        # g_ext: table form when given (runs inside the kernels), else g_ext_fn(r) per call
        self._g_spec = g_ext_spec
        self._engines = {}

    def config(self, **extra):
        kw = dict(self._cfg_kwargs)
        kw.update(extra)
        if self._g_spec is not None:
            kw["g_ext_spec"] = self._g_spec
        else:
            kw["g_ext_array"] = True
        return make_config(self.dim, self._box, self.dx, self.dt, **kw)

    def _engine(self, n):
        if n not in self._engines:
            self._engines[n] = Engine(self.config(), n)
        return self._engines[n]

    def forward_wrapper(self):
        def forward(state: Dict, neighbors=None) -> Dict:
            n = state["r"].shape[0]
            eng = self._engine(n)
            st = dict(state)
            if self._g_spec is None:
                st["g_ext"] = self.g_ext_fn(state["r"])
            eng.upload(st)
            eng.step(0.0, 1, integrate=False, bc=False)
            out = eng.download()
            res = dict(state)  # fields the engine does not hold pass through (solver.py:930-947)
            res.update(out)
            return res

        forward.__self_solver__ = self
        return forward
