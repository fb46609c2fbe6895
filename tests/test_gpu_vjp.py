"""GPU test of row f4: the vector-Jacobian product of `advance`.

The reference is differentiable because it is written in JAX: notebooks/iclr24_grads.ipynb
(cell 5) takes `jax.grad` of 0.5 sum u^2 after five `advance` steps (si_euler with tvf = 0,
jax_sph/integrator.py:22-56, standard SPH with summation density) with respect to the initial
positions and checks it against finite differences.  The engine's hand-written adjoint sweeps
(csrc/adjoint.cuh, sphb200_engine_vjp) are checked the same way: gradient through K steps
(engine.grad_through_steps) against central finite differences of the float64 oracle's loop, for
positions AND velocities, in 2D and 3D, Quintic and Wendland C2 kernels.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _loss(state):
    return 0.5 * float((np.asarray(state["u"], dtype=np.float64) ** 2).sum())


def _oracle_loss(setup64, state, nsteps):
    from oracle import integrator

    setup64.state = {k: v.copy() for k, v in state.items()}
    return _loss(integrator.simulate(setup64, nsteps, fast_segment_sum=True))


CASES = [
    # the notebook's setting: 2D TGV, N = 20^2, viscosity 0.1, 10 warm-up steps, 5 steps
    (dict(case="tgv", dim=2, dx=1.0 / 20, tvf=0.0, viscosity=0.1), 10, 5),
    (dict(case="tgv", dim=2, dx=1.0 / 24, tvf=0.0, viscosity=0.05, kernel="WC2K", h_factor=1.3), 6, 4),
    (dict(case="tgv", dim=3, dx=2 * np.pi / 12, tvf=0.0, viscosity=0.1), 4, 3),
]


@pytest.mark.parametrize("kw,warmup,nsteps", CASES)
def test_gradient_through_steps_matches_finite_differences(kw, warmup, nsteps):
    import torch

    from jax_sph_b200 import config_from_setup
    from jax_sph_b200.engine import grad_through_steps
    from oracle import cases, integrator

    setup64 = cases.make_case(dtype=np.float64, **kw)
    # warm up as the notebook does (non-trivial accelerations, particles off the lattice)
    state0 = integrator.simulate(setup64, warmup, fast_segment_sum=True)
    state0 = {k: np.array(v) for k, v in state0.items()}
    setup32 = cases.make_case(dtype=np.float32, **kw)
    state32 = {k: (v.astype(np.float32) if v.dtype == np.float64 else v.copy()) for k, v in state0.items()}
    cfg = config_from_setup(setup32)

    final, grad = grad_through_steps(cfg, state32, setup32.dt, nsteps,
                                     lambda s: {"u": s["u"].clone()})  # d(0.5 sum u^2)/du = u
    g_r, g_u = grad["r"].cpu().numpy().astype(np.float64), grad["u"].cpu().numpy().astype(np.float64)
    # the engine's forward result is the oracle's (same loss to float32 accuracy)
    loss64 = _oracle_loss(setup64, state0, nsteps)
    assert abs(_loss({"u": final["u"].cpu().numpy()}) - loss64) <= 2e-5 * loss64

    rng = np.random.default_rng(0)
    n, dim = state0["r"].shape
    eps = 1e-3 * setup64.dx
    picks = [(int(rng.integers(n)), int(rng.integers(dim))) for _ in range(10)]
    scale_r, scale_u = np.abs(g_r).max(), np.abs(g_u).max()
    assert scale_r > 0 and scale_u > 0
    for key, g, scale, h in (("r", g_r, scale_r, eps), ("u", g_u, scale_u, 1e-4)):
        for i, k in picks:
            plus = {a: v.copy() for a, v in state0.items()}
            minus = {a: v.copy() for a, v in state0.items()}
            plus[key][i, k] += h
            minus[key][i, k] -= h
            if key == "u":  # advance() derives v from u; keep the pair consistent
                plus["v"][i, k] += h
                minus["v"][i, k] -= h
            fd = (_oracle_loss(setup64, plus, nsteps) - _oracle_loss(setup64, minus, nsteps)) / (2 * h)
            assert abs(g[i, k] - fd) <= 3e-3 * scale + 1e-9, (key, i, k, g[i, k], fd, scale)


def test_unsupported_variants_are_refused():
    import torch

    from jax_sph_b200 import Engine, _lib, config_from_setup
    from oracle import cases

    for kw in (dict(case="tgv", dim=2, dx=0.05, tvf=1.0),
               dict(case="tgv", dim=2, dx=0.05, solver="RIE", density_evolution=True),
               dict(case="db", dim=2, dx=0.05)):
        setup = cases.make_case(dtype=np.float32, **kw)
        eng = Engine(config_from_setup(setup), len(setup.state["r"]))
        eng.upload(setup.state)
        eng.step(setup.dt, 1)
        with pytest.raises(_lib.Sphb200Error):
            eng.vjp(setup.dt, {"u": torch.ones((eng.n, eng.dim), device="cuda")})
