"""Throughput on a DISORDERED particle distribution (GPU box): the bench lattice after `--pre`
steps of the 3D Taylor-Green flow (particles have moved several dx: ~113 instead of 93
neighbours, ragged cells), per-pass CUDA-event times.  python scripts/disordered_bench.py"""

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
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--pre", type=int, nargs="*", default=[0, 600, 2000])
    ap.add_argument("--steps", type=int, default=60)
    a = ap.parse_args()
    state, meta = lattice_state("tgv3d", a.nx)
    n = len(state["r"])
    cfg = make_config(3, meta["box"], meta["dx"], meta["dt"], tvf=meta["tvf"], c_ref=meta["c_ref"],
                      p_ref=meta["p_ref"], uniform_eta=True)
    eng = Engine(cfg, n)
    eng.upload({k: torch.from_numpy(v).pin_memory() for k, v in state.items()})
    done = 0
    for pre in a.pre:
        eng.step(meta["dt"], pre - done + 3)
        done = pre + 3
        torch.cuda.synchronize()
        c0 = eng.counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.step(meta["dt"], a.steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        c1 = eng.counters()
        done += a.steps
        eng.profile(True)
        acc = {}
        for _ in range(10):
            eng.step(meta["dt"], 1)
            for k, v in eng.last_times().items():
                acc[k] = acc.get(k, 0.0) + v / 10
        eng.profile(False)
        done += 10
        ek, um = eng.stats()
        cnt = eng.neighbor_list(0)[1] if n <= 2**22 else -1
        print(f"after {pre:5d} steps (t = {pre * meta['dt']:.3f}): "
              + " ".join(f"{k}={v:.3f}" for k, v in acc.items())
              + f" | {ms:.3f} ms/step over {a.steps} steps ({c1['searches'] - c0['searches']} searches) = "
              + f"{n / ms / 1e3:.1f} M upd/s  err={eng.error()} ekin={ek:.5e} "
              + (f"edges/particle={cnt / n:.1f}" if cnt >= 0 else ""), flush=True)


if __name__ == "__main__":
    main()
