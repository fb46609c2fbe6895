"""Per-pass times of solver variants on the bench lattices (GPU box):
python scripts/variant_bench.py  ->  2D TGV 1 M and 3D TGV 4 M for SPH+tvf, RIE, DELTA."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import lattice_state  # noqa: E402
from jax_sph_b200 import Engine, make_config  # noqa: E402

VARIANTS = {
    "SPH+tvf": dict(tvf=1.0),
    "SPH": dict(tvf=0.0),
    "RIE+evol": dict(tvf=0.0, solver="RIE", is_rho_evol=True),
    "DELTA": dict(tvf=0.0, solver="DELTA"),
    "DELTA+evol": dict(tvf=0.0, solver="DELTA", is_rho_evol=True),
    # 80-byte force record: the tile is shortened so that the shared neighbour lists survive
    "SPH+tvf+heat": dict(tvf=1.0, is_heat_conduction=True),
}


def main():
    for workload, nx in (("tgv2d", 1000), ("tgv3d", 160)):
        state, meta = lattice_state(workload, nx)
        n, dim = len(state["r"]), meta["dim"]
        state["kappa"] = state["rho"] * 7.313
        state["Cp"] = state["rho"] * 305.27
        pinned = {k: torch.from_numpy(v).pin_memory() for k, v in state.items()}
        for name, kw in VARIANTS.items():
            kw = dict(kw)
            tvf = kw.pop("tvf")
            cfg = make_config(dim, meta["box"], meta["dx"], meta["dt"], tvf=tvf, c_ref=meta["c_ref"],
                              p_ref=meta["p_ref"], **kw)
            eng = Engine(cfg, n)
            eng.upload(pinned)
            eng.step(meta["dt"], 3)
            torch.cuda.synchronize()
            eng.profile(True)
            acc, steps = {}, 5
            for _ in range(steps):
                eng.step(meta["dt"], 1)
                for k, v in eng.last_times().items():
                    acc[k] = acc.get(k, 0.0) + v / steps
            print(f"{workload} N={n} {name:12s} tile={eng.plan()['tile']}: " + " ".join(f"{k}={v:.3f}" for k, v in acc.items())
                  + f" | {n / acc['total'] / 1e3:.1f} M upd/s err={eng.error()}", flush=True)
            eng.close()
            del eng
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
