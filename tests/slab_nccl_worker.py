"""Worker of test_gpu_slab_nccl.py: launched with torch.distributed.run, one rank per GPU.
Every rank steps its slab (ring transport chosen by SPHB200_SLAB_TRANSPORT); rank 0 gathers the global state and checks
it against a single-engine run of the same state."""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist

    from _util import assert_close
    from jax_sph_b200 import Engine, SlabEngine, config_from_setup
    from oracle import cases

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    nsteps = 40
    for kw in (dict(case="tgv", dim=2, dx=0.0125, tvf=1.0),
               dict(case="tgv", dim=3, dx=2 * np.pi / 32, tvf=1.0, viscosity=0.02),
               # (a slab must be two cutoffs thick: the channel is 0.5 deep, 15 cell layers at
               # dx = 0.02 with the list skin -- enough for two ranks, not for four)
               dict(case="ht", dim=3, dx=0.02 if world <= 2 else 0.0125),
               # Delta-SPH density diffusion: five exchanges per step
               dict(case="tgv", dim=3, dx=2 * np.pi / 32, solver="DELTA", density_evolution=True,
                    viscosity=0.02)):
        setup = cases.make_case(dtype=np.float32, **kw)
        n = len(setup.state["r"])
        eng = SlabEngine(config_from_setup(setup))
        transport = eng.transport
        assert transport == os.environ.get("SPHB200_SLAB_TRANSPORT", transport)
        eng.upload(setup.state)
        eng.step(setup.dt, nsteps)
        err = eng.error()
        ek, umax = eng.stats()
        got = eng.gather(n, root=0)
        assert err == 0, f"device error word {err}"
        if rank == 0:
            single = Engine(config_from_setup(setup), n)
            single.upload(setup.state)
            single.step(setup.dt, nsteps)
            ref = {k: v.numpy() for k, v in single.download(host=True).items()}
            for k in ("r", "u", "v", "rho", "p", "dudt", "dvdt", "T"):
                assert_close(k, got[k], ref[k], setup, factor=8.0, what=f"{kw} P={world}")
            ek_ref, umax_ref = single.stats()
            assert abs(ek - ek_ref) <= 1e-5 * ek_ref and abs(umax - umax_ref) <= 1e-5 * umax_ref
            print(f"case {kw['case']}{kw['dim']}d N={n} P={world}: ok", flush=True)
        dist.barrier()
    if rank == 0:
        print(f"SLAB_NCCL_OK transport={transport}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
