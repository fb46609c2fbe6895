"""Development check (GPU box): the duo sweeps (csrc/sweep2.cuh) against the per-particle sweeps
(csrc/sweep.cuh, SPHB200_DUO=0) of the same library on the same states -- forward and a run of
steps long enough to re-sort several times -- plus the oracle on a small case.  Not a test;
tests/ holds the asserted versions.  Usage: python scripts/duo_check.py [nx3d] [nsteps]"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jax_sph_b200 import Engine, config_from_setup  # noqa: E402
from oracle import cases  # noqa: E402


def engine(setup, duo, **tune):
    os.environ["SPHB200_DUO"] = "1" if duo else "0"
    eng = Engine(config_from_setup(setup, **tune), len(setup.state["r"]))
    os.environ.pop("SPHB200_DUO")
    return eng


def compare(name, setup, nsteps, noise=0.0, seed=0, **tune):
    st = {k: v.copy() for k, v in setup.state.items()}
    if noise > 0:
        rng = np.random.default_rng(seed)
        r = st["r"] + rng.normal(0.0, noise * setup.dx, st["r"].shape).astype(np.float32)
        st["r"] = np.mod(r, np.asarray(setup.box_size, dtype=np.float32)).astype(np.float32)
    out = {}
    for duo in (False, True):
        eng = engine(setup, duo, **tune)
        eng.upload(st)
        eng.step(0.0, 1, integrate=False, bc=False)
        fwd = {k: v.numpy().copy() for k, v in eng.download(host=True).items()}
        eng.upload(st)
        eng.step(setup.dt, nsteps)
        adv = {k: v.numpy().copy() for k, v in eng.download(host=True).items()}
        cnt = eng.counters() if hasattr(eng, "counters") else None
        out[duo] = (fwd, adv, eng.error(), eng.plan(), cnt)
    print(f"[{name}] N={len(st['r'])} noise={noise} steps={nsteps}")
    print("   classic plan", out[False][3], "err", out[False][2], out[False][4])
    print("   duo     plan", out[True][3], "err", out[True][2], out[True][4])
    for what, i in (("forward", 0), ("advance", 1)):
        line = []
        for k in ("r", "u", "v", "rho", "p", "dudt", "dvdt"):
            a, b = out[True][i][k], out[False][i][k]
            scale = float(np.abs(b).max()) + (setup.p_ref if k == "p" else 0.0)
            line.append(f"{k}={float(np.abs(a - b).max()) / max(scale, 1e-30):.2e}")
        print(f"   duo vs classic, {what}: " + " ".join(line))


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    s3 = cases.make_case("tgv", dim=3, dx=2 * np.pi / nx, dtype=np.float32, tvf=1.0, viscosity=0.02)
    compare("tgv3d_tvf lattice", s3, nsteps)
    compare("tgv3d_tvf noisy", s3, nsteps, noise=0.15)
    s3p = cases.make_case("tgv", dim=3, dx=2 * np.pi / nx, dtype=np.float32, tvf=0.0, viscosity=0.02)
    compare("tgv3d_plain noisy", s3p, nsteps, noise=0.15, seed=1)
    s2 = cases.make_case("tgv", dim=2, dx=1.0 / 200, dtype=np.float32, tvf=1.0)
    compare("tgv2d_tvf noisy", s2, nsteps, noise=0.15, seed=2)
    s2w = cases.make_case("tgv", dim=2, dx=1.0 / 100, dtype=np.float32, tvf=1.0, kernel="WC2K", h_factor=1.3)
    compare("tgv2d_wc2k noisy", s2w, nsteps, noise=0.1, seed=3)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
