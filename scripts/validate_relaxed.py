"""The 3D Taylor-Green validation of the reference AS ITS SCRIPTS RUN IT (validation/tgv3d.sh:
19-20): a relaxation run (case.mode=rlx, noise 0.25 dx, tvf = 1, p_bg_factor 0.02, 5000 steps)
followed by the simulation from the relaxed positions (SPH, tvf = 1, viscosity 0.02, t_end = 10),
E_kin(t) = get_ekin / volume (validate.py:116-117) against the JAX-Fluids Nx = 64 curve the
reference plots (tests/golden/validation_tgv3d_re50.csv) -- next to the Cartesian-lattice start
tests/test_gpu_validation.py bounds.

    python scripts/validate_relaxed.py [nx ...]      (default 32 64)
"""

import os
import re
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jax_sph_b200 import case_setup  # noqa: E402
from jax_sph_b200.simulate import defaults, simulate  # noqa: E402


def curve(cfg, every):
    lines = []
    cfg["io"]["write_every"] = every
    eng = simulate(cfg, log=lines.append)
    t, ek = [], []
    for l in lines:
        m = re.search(r"t=([\d.]+), Ekin=([\d.]+)", l)
        if m:
            t.append(float(m.group(1)))
            ek.append(float(m.group(2)))
    eng.close()
    return np.array(t), np.array(ek) / (2 * np.pi) ** 3


def main():
    ref = np.loadtxt(os.path.join(ROOT, "tests", "golden", "validation_tgv3d_re50.csv"), delimiter=",")
    for nx in [int(a) for a in sys.argv[1:]] or [32, 64]:
        dx = 2 * np.pi / nx
        with tempfile.TemporaryDirectory() as tmp:
            t0 = time.time()
            simulate(defaults(seed=123, case=dict(name="tgv", dim=3, dx=dx, mode="rlx",
                                                  r0_noise_factor=0.25, viscosity=0.02),
                              solver=dict(tvf=1.0), eos=dict(p_bg_factor=0.02),
                              io=dict(write_type=["h5"], write_every=2500, data_path=tmp)), log=None)
            t_rlx = time.time() - t0
            path = os.path.join(tmp, case_setup.relaxed_state_name("tgv", 3, dx, 123) + ".h5")
            base = dict(name="tgv", dim=3, dx=dx, viscosity=0.02)
            out = {}
            for label, case in (("relaxed", dict(base, r0_type="relaxed", state0_path=path)),
                                ("lattice", dict(base))):
                t0 = time.time()
                t, ek = curve(defaults(seed=123, case=case, solver=dict(tvf=1.0, t_end=10.0),
                                       io=dict(data_path=tmp)), every=50)
                e_ref = np.interp(t, ref[:, 0], ref[:, 2])
                rel = np.abs(ek - e_ref) / ref[0, 2]
                out[label] = (rel.max(), t[rel.argmax()], ek[-1], e_ref[-1], time.time() - t0, len(t))
            print(f"tgv3d Re=50 nx={nx} (N={nx**3}), relaxation 5000 steps in {t_rlx:.1f} s")
            for label, (mx, at, e_end, r_end, wall, ns) in out.items():
                print(f"  {label:8s} start: max|Ek - ref|/Ek0 = {mx:.4f} at t = {at:.2f}; "
                      f"Ek(10) = {e_end:.5f} (ref {r_end:.5f}); {ns} samples, {wall:.1f} s")


if __name__ == "__main__":
    main()
