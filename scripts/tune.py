"""Tuning sweep (GPU box): per-pass times of the resident step for a list of
(cell_sub, tile, threads, list_cap) plans.  Usage:
  python scripts/tune.py [--workload tgv3d] [--nx 256] [--steps 3] "sub,tx,ty,tz,threads,lcap" ..."""

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import lattice_state  # noqa: E402
from jax_sph_b200 import Engine, make_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="tgv3d")
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("plans", nargs="*")
    args = ap.parse_args()
    state, meta = lattice_state(args.workload, args.nx)
    n = len(state["r"])
    dim = meta["dim"]
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in state.items()}
    for plan in args.plans or ["0,0,0,0,0,0"]:
        sub, tx, ty, tz, threads, lcap = [int(x) for x in plan.split(",")][:6]
        stage = int(plan.split(",")[6]) if len(plan.split(",")) > 6 else 0
        try:
            cfg = make_config(dim, meta["box"], meta["dx"], meta["dt"], tvf=meta["tvf"],
                              c_ref=meta["c_ref"], p_ref=meta["p_ref"],
                              cell_sub=[sub] * dim if sub else None,
                              tile=[tx, ty, tz] if tx else None, threads=threads, list_cap=lcap,
                              stage_cap=stage)
            eng = Engine(cfg, n)
            eng.upload(pinned)
            eng.step(meta["dt"], 2)
            torch.cuda.synchronize()
            eng.profile(True)
            acc = {}
            for _ in range(args.steps):
                eng.step(meta["dt"], 1)
                for k, v in eng.last_times().items():
                    acc[k] = acc.get(k, 0.0) + v / args.steps
            err = eng.error()
            ek, um = eng.stats()
            p = eng.plan()
            print(f"plan {plan}: cells={p['cells']} tile={p['tile']} thr={p['threads']} "
                  f"lcap={p['list_cap']} cap={p['stage_cap']} err={err} ekin={ek:.6e} | "
                  + " ".join(f"{k}={v:.3f}" for k, v in acc.items())
                  + f" | {n / acc['total'] / 1e3:.1f} M upd/s", flush=True)
            eng.close()
            del eng
        except Exception as ex:  # noqa: BLE001
            print(f"plan {plan}: FAILED {ex}", flush=True)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
