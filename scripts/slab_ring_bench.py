"""Drive a local ring of slab engines on ONE GPU (device-to-device copies as transport) on the
bench lattice -- for ncu launch lists of the slab pipeline and for comparing the sum of the
slab kernels with the single engine.  python scripts/slab_ring_bench.py --nx 160 --ranks 2"""

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from jax_sph_b200 import Engine, SlabEngine, make_config
    from jax_sph_b200.slab import layer_of, step_local_ring

    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=160)
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--single", action="store_true")
    args = ap.parse_args()
    meta = bench.lattice_meta("tgv3d", args.nx)

    def cfg():
        return make_config(3, meta["box"], meta["dx"], meta["dt"], tvf=1.0, c_ref=meta["c_ref"],
                           p_ref=meta["p_ref"])

    if args.single:
        state, _ = bench.lattice_state("tgv3d", args.nx)
        eng = Engine(cfg(), len(state["r"]))
        eng.upload(state)
        eng.step(meta["dt"], 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.step(meta["dt"], args.steps)
        torch.cuda.synchronize()
        print("single: %.3f ms/step" % ((time.perf_counter() - t0) / args.steps * 1e3))
        return
    ring = [SlabEngine(cfg(), r, args.ranks) for r in range(args.ranks)]
    ax = ((np.arange(args.nx, dtype=np.float32) + np.float32(0.5)) * np.float32(meta["dx"])).astype(np.float32)
    for e in ring:
        lay = layer_of(ax, e.inv_cell, e.layers)
        state, _ = bench.lattice_state("tgv3d", args.nx, planes=(lay >= e.z0) & (lay < e.z1))
        ids = state.pop("ids")
        e.upload(state, ids)
    step_local_ring(ring, meta["dt"], 3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_local_ring(ring, meta["dt"], args.steps)
    torch.cuda.synchronize()
    print("ring of %d on one GPU: %.3f ms/step (all ranks serialised)" %
          (args.ranks, (time.perf_counter() - t0) / args.steps * 1e3))
    for e in ring:
        print(e.rank, e.counts(), e.error(reduce=False))


if __name__ == "__main__":
    main()
