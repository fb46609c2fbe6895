"""Semi-implicit Euler integrator, host mirror of jax_sph/integrator.py:8-58.

``si_euler(tvf, model, shift_fn, bc_fn, nw_fn)`` keeps the reference signature.
``model`` must be ``WCSPH(...).forward_wrapper()`` of this package and ``bc_fn``
a ``BcTable`` (the table form of the shipped cases' boundary callables, which
the fused CUDA epilogue applies); the returned ``advance(dt, state, neighbors)``
is one fused engine step: kick + drift + wrap, cell rebuild, sweeps, bc.
"""

from typing import Callable, Dict

from . import _lib
from .engine import Engine


class BcTable:
    """Table form of a case ``_boundary_conditions_fn`` (cases/db.py:134-142, cf.py:143-160,
    pf.py, ht.py:154-187): {"tags": {tag: {u, v, zero_dudt, zero_dvdt, p, T, zero_dTdt}},
    "inflow_x": {x, T} | None, "outflow_x": {x} | None}."""

    def __init__(self, table=None):
        self.table = table or {"tags": {}, "inflow_x": None, "outflow_x": None}


def si_euler(tvf: float, model: Callable, shift_fn: Callable, bc_fn, nw_fn: Callable = None):
    solver = getattr(model, "__self_solver__", None)
    if solver is None:
        raise _lib.Sphb200Error("model must be jax_sph_b200.solver.WCSPH(...).forward_wrapper()")
    if nw_fn is not None:
        raise NotImplementedError("per-step wall-normal recomputation (moving walls) is not on "
                                  "the fused path")
    table = bc_fn.table if isinstance(bc_fn, BcTable) else None
    if table is None:
        raise _lib.Sphb200Error("bc_fn must be a BcTable (table form of the case boundary fn)")
    engines = {}

    def advance(dt: float, state: Dict, neighbors=None):
        n = state["r"].shape[0]
        if n not in engines:
            engines[n] = Engine(solver.config(tvf=tvf, bc_table=table), n)
        eng = engines[n]
        st = dict(state)
        if solver._g_spec is None:
            raise _lib.Sphb200Error("the fused advance needs g_ext in table form (g_ext_spec)")
        eng.upload(st)
        eng.step(dt, 1, integrate=True, bc=True)
        out = dict(state)
        out.update(eng.download())
        if neighbors is not None and hasattr(neighbors, "update"):
            neighbors = neighbors.update(out["r"])
        return out, neighbors

    advance.engines = engines
    return advance
