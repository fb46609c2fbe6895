"""Error of the engine against the float32 and float64 oracle on the smoke() case (dev tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if len(sys.argv) < 2 else sys.argv[1]
sys.path.insert(0, ROOT)
from jax_sph_b200 import Engine, config_from_setup
from oracle import cases, integrator
for nx, nsteps in ((20, 2), (20, 0), (16, 5)):
    kw = dict(dim=3, dx=2 * np.pi / nx, tvf=1.0, viscosity=0.02)
    s32 = cases.make_case("tgv", dtype=np.float32, **kw)
    s64 = cases.make_case("tgv", dtype=np.float64, **kw)
    for k, v in s32.state.items():
        s64.state[k] = v.astype(np.float64) if v.dtype == np.float32 else v.copy()
    eng = Engine(config_from_setup(s32), len(s32.state["r"]))
    eng.upload(s32.state)
    if nsteps == 0:
        eng.step(0.0, 1)
        r32 = integrator.simulate(s32, 0); r64 = integrator.simulate(s64, 0)
        from oracle.solver import WCSPH
        def fwd(s):
            so = WCSPH(s.displacement_fn, s.eos, s.g_ext_fn, s.dx, s.dim, s.dt, s.c_ref, s.eta_limiter, 0.0, 0.0, s.solver, s.kernel, s.h_factor, s.is_bc_trick, s.density_evolution, s.artificial_alpha, s.free_slip, s.density_renormalize, s.heat_conduction, dtype=s.dtype)
            nfn = integrator.make_neighbors_fn(s.box_size, so._kernel_fn.cutoff)
            return so.forward({k: v.copy() for k, v in s.state.items()}, nfn(s.state["r"]))
        r32, r64 = fwd(s32), fwd(s64)
    else:
        eng.step(s32.dt, nsteps)
        r32 = integrator.simulate(s32, nsteps); r64 = integrator.simulate(s64, nsteps)
    got = eng.download(host=True)
    for k in ("rho", "p", "dudt", "dvdt"):
        a = got[k].numpy().astype(np.float64)
        sc = np.abs(r64[k]).max()
        print(f"nx={nx} steps={nsteps} {k:5s} scale {sc:.3e}  |eng-o32| {np.abs(a - r32[k]).max():.3e}  |eng-o64| {np.abs(a - r64[k]).max():.3e}  |o32-o64| {np.abs(r32[k] - r64[k]).max():.3e}")
