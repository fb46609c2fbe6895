"""Semi-implicit Euler integrator, host mirror of jax_sph/integrator.py:8-58.

``si_euler(tvf, model, shift_fn, bc_fn, nw_fn)`` keeps the reference signature.
``model`` must be ``WCSPH(...).forward_wrapper()`` of this package and ``bc_fn``
a ``BcTable`` (the table form of the shipped cases' boundary callables, which
the fused CUDA epilogue applies); the returned ``advance(dt, state, neighbors)``
is one fused engine step: kick + drift + wrap, cell sort + neighbour search when a particle has
moved half the list skin, exact membership test, sweeps, bc.
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
    # nw_fn (integrator.py:33-34): the engine recomputes the wall normals itself from the one-layer
    # discretisation of the wall surface -- pass its table form {"layer", "offset", "cutoff"}
    # (what compute_nws_jax_wrapper closes over, utils.py:197-277), or an object carrying it as .spec
    nw_spec = getattr(nw_fn, "spec", nw_fn)
    if nw_spec is not None and not (isinstance(nw_spec, dict) and {"layer", "offset", "cutoff"} <= set(nw_spec)):
        raise _lib.Sphb200Error("nw_fn must be the table form {'layer', 'offset', 'cutoff'} of the "
                                "wall-normal function (arbitrary callables cannot run inside the "
                                "fused step)")
    table = bc_fn.table if isinstance(bc_fn, BcTable) else None
    if table is None:
        raise _lib.Sphb200Error("bc_fn must be a BcTable (table form of the case boundary fn)")
    engines = {}

    def advance(dt: float, state: Dict, neighbors=None):
        n = state["r"].shape[0]
        if solver._g_spec is None:
            raise _lib.Sphb200Error("the fused advance needs g_ext in table form (g_ext_spec)")
        if n not in engines:
            engines[n] = Engine(solver.config(tvf=tvf, bc_table=table, wall_layer=nw_spec), n)
        eng = engines[n]
        # the state goes into the slots its particles already occupy: cells and neighbour lists
        # survive from call to call as far as the positions allow (Engine.refresh)
        eng.refresh(dict(state))
        eng.step(dt, 1, integrate=True, bc=True)
        out = dict(state)  # entries advance() does not write pass through, as in the reference
        out.update(eng.download(keys=eng.live_fields()[1]))
        if neighbors is not None and hasattr(neighbors, "update"):
            neighbors = neighbors.update(out["r"])
        return out, neighbors

    advance.engines = engines
    return advance
