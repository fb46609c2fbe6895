"""Validation curves on the GPU (validation/tgv2d.sh, tgv3d.sh, validate.py:181-218).

    python scripts/validate_curves.py [out.json]

Runs the engine on the 2D Taylor-Green vortex (Re = 100) and the 3D Taylor-Green vortex
(Re = 50, the case validation/tgv3d.sh:20 runs: SPH, tvf = 1, viscosity 0.02) from the
Cartesian lattice and records E_kin(t) and u_max(t); tests/test_gpu_validation.py asserts
the same curves against the analytical decay / the JAX-Fluids reference curve.
"""

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_curve(kw, t_end, nsamples):
    from jax_sph_b200 import Engine, config_from_setup
    from oracle import cases

    setup = cases.make_case(dtype=np.float32, **kw)
    n = len(setup.state["r"])
    eng = Engine(config_from_setup(setup), n)
    eng.upload(setup.state)
    eng.step(0.0, 1)  # simulate.py:110: initialise the accelerations
    nsteps = int(round(t_end / setup.dt))
    every = max(1, nsteps // nsamples)
    vol = float(np.prod(setup.box_size))
    t, ek, um = [0.0], [eng.stats()[0] / vol], [eng.stats()[1]]
    t0 = time.time()
    done = 0
    while done < nsteps:
        k = min(every, nsteps - done)
        eng.step(setup.dt, k)
        done += k
        e, u = eng.stats()
        t.append(done * setup.dt)
        ek.append(e / vol)
        um.append(u)
    assert eng.error() == 0
    return dict(n=n, dt=setup.dt, steps=nsteps, wall_s=time.time() - t0, t=t, ekin=ek, umax=um)


def main():
    out = {}
    for dx in (0.02, 0.01):
        for name, kw in (("tvf", dict(tvf=1.0)), ("notvf", dict()),
                         ("rie", dict(solver="RIE", density_evolution=True))):
            out[f"tgv2d_{name}_dx{dx}"] = run_curve(dict(case="tgv", dim=2, dx=dx, **kw), 5.0, 50)
    for nx in (32, 64):
        out[f"tgv3d_tvf_nx{nx}"] = run_curve(
            dict(case="tgv", dim=3, dx=2 * np.pi / nx, tvf=1.0, viscosity=0.02), 10.0, 100)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "validation_curves.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f)
    ref = np.loadtxt(os.path.join(ROOT, "tests", "golden", "validation_tgv3d_re50.csv"), delimiter=",")
    for k, v in out.items():
        t, ek, um = np.array(v["t"]), np.array(v["ekin"]), np.array(v["umax"])
        if k.startswith("tgv2d"):
            th_u = np.exp(-8 * np.pi**2 / 100 * t)
            th_e = 0.25 * th_u**2
            print(f"{k}: N={v['n']} steps={v['steps']} wall={v['wall_s']:.1f}s  "
                  f"max|log(umax/theory)|={np.abs(np.log(um / th_u)).max():.3f}  "
                  f"max|log(Ek/theory)|={np.abs(np.log(ek / th_e)).max():.3f}  "
                  f"at t=2: umax {um[np.argmin(abs(t-2))]:.4f} theory {np.exp(-8*np.pi**2/100*2):.4f}")
        else:
            e_ref = np.interp(t[1:], ref[:, 0], ref[:, 2])
            rel = np.abs(ek[1:] - e_ref) / e_ref[0]
            print(f"{k}: N={v['n']} steps={v['steps']} wall={v['wall_s']:.1f}s  "
                  f"max|Ek - ref|/Ek0 = {rel.max():.4f} at t={t[1:][rel.argmax()]:.2f}; "
                  f"Ek(10) {ek[-1]:.5f} ref {e_ref[-1]:.5f}")


if __name__ == "__main__":
    main()
