"""si_euler (oracle; test infrastructure only).  jax_sph/integrator.py:8-58."""

import numpy as np

from . import partition


def si_euler(tvf, model, shift_fn, bc_fn, nw_fn=None):
    """integrator.py:8-58.  ``neighbors`` is a callable r -> idx (2, E)."""

    def advance(dt, state, neighbors_fn):
        state = dict(state)
        t = state["r"].dtype.type
        state["u"] = state["u"] + t(1.0 * dt) * state["dudt"]  # :26
        state["v"] = state["u"] + t(tvf * 0.5 * dt) * state["dvdt"]  # :27
        state["r"] = shift_fn(state["r"], t(1.0 * dt) * state["v"])  # :30
        if nw_fn is not None:  # :33-34
            state["nw"] = nw_fn(state["r"])
        idx = neighbors_fn(state["r"])  # :48
        state = model(state, idx)  # :51
        state = bc_fn(state)  # :54
        return state, idx

    return advance


def make_neighbors_fn(box, cutoff, faithful=False):
    """r -> idx; ``faithful`` uses the jax-md cell-list restatement (small N)."""

    def fn(r):
        if faithful:
            return partition.canonical_pairs(
                partition.neighbor_list_reference(r, box, cutoff).idx, N=len(r)
            )
        return partition.neighbor_pairs(r, box, cutoff)

    return fn


def simulate(case, n_steps, dtype=None, faithful_nl=False, fast_segment_sum=False, callback=None):
    """The loop of jax_sph/simulate.py:110-134 without IO: the discarded compile
    call ``advance(0.0, ...)`` is skipped, then ``n_steps`` calls of advance(dt)."""
    from .solver import WCSPH

    setup = case
    solver = WCSPH(
        setup.displacement_fn, setup.eos, setup.g_ext_fn, setup.dx, setup.dim, setup.dt,
        setup.c_ref, setup.eta_limiter, setup.diff_delta, setup.diff_alpha, setup.solver, setup.kernel, setup.h_factor,
        setup.is_bc_trick, setup.density_evolution, setup.artificial_alpha, setup.free_slip,
        setup.density_renormalize, setup.heat_conduction, dtype=setup.dtype,
        fast_segment_sum=fast_segment_sum,
    )
    nfn = make_neighbors_fn(setup.box_size, solver._kernel_fn.cutoff, faithful=faithful_nl)
    advance = si_euler(setup.tvf, solver.forward, setup.shift_fn, setup.bc_fn,
                       getattr(setup, "nw_fn", None))
    state = {k: np.array(v, copy=True) for k, v in setup.state.items()}
    for step in range(n_steps):
        state, _ = advance(setup.dt, state, nfn)
        if callback is not None:
            callback(step, state)
    return state
