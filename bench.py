"""bench.py -- particle-updates/s of the JAX-SPH per-step hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Workload at N=1: BASELINE.json configs[3], the 3D Taylor-Green vortex of
validation/tgv3d.sh (SPH, tvf=1, viscosity 0.02, Quintic kernel) on a 256^3
lattice = 16 777 216 particles, float32, synthetic lattice initialisation.
One "step" = one advance(dt) (integrator.py:22-56): kick + drift + wrap,
neighbour-structure rebuild, density sweep + EoS, force sweep.

Prints ONE JSON line (rank 0).  `value` times the resident engine (state in
HBM); `e2e` times the same step through the public host-buffer API with the
full state copied host->device before and device->host after every step.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec"
UNIT = "particle-updates/s"

# algorithmic HBM bytes per particle-step (float32, 3D / 2D), DESIGN.md section 4
# (state quads + the particle's neighbour-list row: 93 / 25 entries in chunks of 8 x 2 B + count)
BYTES = {
    3: dict(cells=276, density=64 + 196, force=112 + 196),
    2: dict(cells=276, density=64 + 68, force=112 + 68),
}
# useful flops per directed in-range edge, SURVEY.md section 8d
FLOPS_EDGE = {3: dict(density=58 + 1, force=58 + 103 + 16), 2: dict(density=50 + 1, force=50 + 63 + 14)}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # nominal non-tensor FP32 FMA peak


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the sweep, from the committed
    ncu capture of this workload (profiles/r01_traffic.json); None when absent."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def roofline_of(acc, dim, n, edges, peaks, which):
    """`roofline` object: the DOMINANT kernel of the step (largest CUDA-event time) against
    the HBM roofline as the contract asks, with its FP32 figure (the sweeps are issue-bound,
    DESIGN.md section 3) and the same numbers for the other sweep under `kernels`."""
    names = {"density": "k_sweep<PhysDensity, LIST_BUILD> (density sweep + neighbour-list builder)",
             "force": "k_sweep<PhysForce, LIST_CONSUME> (force sweep)"}
    per = {}
    for k in ("density", "force"):
        ms = acc.get(k, 0.0)
        if not ms:
            continue
        gbs = BYTES[dim][k] * n / (ms * 1e-3) / 1e9
        tf = FLOPS_EDGE[dim][k] * edges * n / (ms * 1e-3) / 1e12
        per[k] = {"kernel": names[k], "ms": ms, "achieved": gbs, "frac": gbs / peaks["hbm_gbs"],
                  "traffic": ncu_traffic(k),
                  "fp32": {"achieved_tflops": tf, "peak_tflops_nominal": FP32_PEAK_TFLOPS,
                           "frac": tf / FP32_PEAK_TFLOPS}}
    if not per:
        return {"kernel": None, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": None, "traffic": None, "passes_ms": acc}
    top = max(per, key=lambda k: per[k]["ms"])
    c_ms = acc.get("cells", 0.0)
    roof = {
        "kernel": per[top]["kernel"], "bound": "hbm", "achieved": per[top]["achieved"],
        "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": which, "frac": per[top]["frac"],
        "traffic": per[top]["traffic"], "ms": per[top]["ms"],
        "note": "the sweeps are bound by FP32/ALU issue and shared-memory wavefronts, not by HBM "
                "(20 flop per algorithmic byte, ridge at 11): see fp32 and DESIGN.md section 3; "
                "the HBM-bound passes are under cells",
        "fp32": per[top]["fp32"], "kernels": per,
        "cells": {"kernel": "k_hash + scan + k_scatter_src + k_reorder (integrate, sort, reorder)",
                  "ms": c_ms,
                  "achieved": BYTES[dim]["cells"] * n / (c_ms * 1e-3) / 1e9 if c_ms else None,
                  "frac": BYTES[dim]["cells"] * n / (c_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]
                  if c_ms else None},
        "passes_ms": acc,
    }
    return roof


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            self.proc.terminate()
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4)
                          if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def ht3d_meta(nx):
    """BASELINE configs[4]: cases/ht.yaml with case.dim=3 (3D channel with a hot patch in the
    bottom wall, cases/ht.py:29-187): box L x (H + 6 dx) x W = 1.0 x (0.2 + 6 dx) x 0.5,
    dx = 1 / nx (nx = 855 -> 64.6 M particles), SPH + is_bc_trick + heat_conduction,
    g_ext = 2.3 e_x between the walls, p_bg = 0.05 p_ref, kappa 7.313, Cp 305.27; bc_fn and
    g_ext_fn in table form (what oracle/cases.py builds and tests/test_host_logic.py checks)."""
    dx = 1.0 / nx
    n_walls, g_mag, viscosity, u_ref = 3, 2.3, 0.01, 1.0
    c_ref = 10.0 * u_ref
    eps = float(np.finfo(np.float32).eps)
    dt = float(min(0.25 * dx / (c_ref + u_ref), 0.25 * dx * dx / (viscosity + eps),
                   0.25 * (dx / (g_mag + eps)) ** 0.5))  # case_setup.py:94-97
    nxyz = [int(round(1.0 / dx)), int(round(0.2 / dx)) + 2 * n_walls, int(round(0.5 / dx))]
    box = [1.0, 0.2 + 2 * n_walls * dx, 0.5]
    # (box / dx).round() of utils.py:51 must not put a lattice plane outside the periodic box
    # (nx = 855: 0.5 / dx = 427.5 rounds to 428 planes, the last one at z > 0.5): such sizes are
    # refused here instead of being reported later as SPHB200_ERR_OUTSIDE_BOX / _SLAB_OVERFLOW
    if any((n - 0.5) * dx >= b * (1 - 1e-6) for n, b in zip(nxyz, box)):
        raise SystemExit(f"bench.py: ht3d nx={nx}: lattice {nxyz} does not fit the box {box}; "
                         "pick nx with 0.5 * nx and 0.2 * nx away from half-integers (e.g. 854)")
    zero = [0.0, 0.0, 0.0]
    st = dict(u=zero, v=zero, zero_dudt=True, zero_dvdt=True, zero_dTdt=True)
    bc_table = {"tags": {1: dict(st, T=1.0), 3: dict(st, T=1.23)},  # SOLID_WALL, DIRICHLET_WALL
                "inflow_x": dict(x=float(n_walls * dx), T=1.0),
                "outflow_x": dict(x=float(box[0] - n_walls * dx))}
    g_ext_spec = {"mode": "band", "g": [g_mag, 0.0, 0.0], "axis": 1, "lo": float(n_walls * dx),
                  "hi": float(box[1] - n_walls * dx)}
    return dict(dim=3, box=box, dx=dx, dt=dt, viscosity=viscosity, c_ref=c_ref, p_ref=c_ref**2,
                p_bg=0.05 * c_ref**2, tvf=0.0, nxyz=nxyz, n_walls=n_walls, kappa=7.313, Cp=305.27,
                cfg_kwargs=dict(is_bc_trick=True, is_heat_conduction=True, p_bg=0.05 * c_ref**2,
                                g_ext_spec=g_ext_spec, bc_table=bc_table))


def ht3d_state(nx, planes=None):
    """Lattice state of ht3d_meta(nx): walls (3 layers below and above) and fluid on one regular
    lattice (i + 0.5) dx (where H / dx is not an integer the reference leaves a sub-dx gap under
    the top wall; the synthetic input keeps the lattice regular), tags SOLID_WALL / DIRICHLET_WALL (hot patch |x - 0.5| < 0.25 of the
    bottom wall, ht.py:90-97) / FLUID, fluid at rest, T = 1 (no position noise: lattice input)."""
    meta = ht3d_meta(nx)
    dx, (n0, n1, n2), nw = np.float32(meta["dx"]), meta["nxyz"], meta["n_walls"]
    kk = np.arange(n2) if planes is None else np.nonzero(planes)[0]
    IX, IY, IK = np.meshgrid(np.arange(n0), np.arange(n1), kk, indexing="xy")
    ix, iy, ik = IX.ravel(), IY.ravel(), IK.ravel()
    r = np.stack([(ix + np.float32(0.5)) * dx, (iy + np.float32(0.5)) * dx,
                  (ik + np.float32(0.5)) * dx], axis=1).astype(np.float32)
    wall = (iy < nw) | (iy >= n1 - nw)
    hot = (iy < nw) & (r[:, 0] < 0.5 + 0.25) & (r[:, 0] > 0.5 - 0.25)
    tag = np.where(hot, 3, np.where(wall, 1, 0)).astype(np.int32)
    n = len(r)
    ones, zv = np.ones(n, dtype=np.float32), np.zeros((n, 3), dtype=np.float32)
    state = dict(r=r, u=zv, v=zv.copy(), dudt=zv.copy(), dvdt=zv.copy(), rho=ones.copy(),
                 p=ones * np.float32(meta["p_bg"]), drhodt=np.zeros(n, dtype=np.float32),
                 mass=ones * np.float32(meta["dx"] ** 3), eta=ones * np.float32(meta["viscosity"]),
                 T=np.where(hot, np.float32(1.23), np.float32(1.0)).astype(np.float32),  # bc_fn at setup
                 dTdt=np.zeros(n, dtype=np.float32),
                 kappa=ones * np.float32(meta["kappa"]), Cp=ones * np.float32(meta["Cp"]), tag=tag)
    if planes is not None:
        state["ids"] = ((iy * n0 + ix) * n2 + ik).astype(np.int32)
    return state, meta


def lattice_meta(workload, nx):
    if workload == "ht3d":
        return ht3d_meta(nx)
    dim, box = (3, 2 * np.pi) if workload == "tgv3d" else (2, 1.0)
    dx = box / nx
    viscosity, u_ref = (0.02, 1.0) if dim == 3 else (0.01, 1.0)
    c_ref = 10.0 * u_ref
    # case_setup.py:94-97 (CFL 0.25)
    dt = float(min(0.25 * dx / (c_ref + u_ref), 0.25 * dx * dx / viscosity))
    return dict(dim=dim, box=[box] * dim, dx=dx, dt=dt, viscosity=viscosity, c_ref=c_ref,
                p_ref=c_ref**2, tvf=1.0)


def lattice_state(workload, nx, planes=None):
    """Synthetic lattice initial state WITHOUT importing the oracle (product path):
    particles at (i + 0.5) dx, TGV velocity field (cases/tgv.py:37-51), rho = 1.
    `planes`: boolean mask over the lattice planes along the last axis (a rank's slab);
    the state then also carries `ids`, the particles' indices in the full lattice."""
    if workload == "ht3d":
        return ht3d_state(nx, planes)
    if workload == "tgv3d":
        dim, box = 3, 2 * np.pi
    else:
        dim, box = 2, 1.0
    dx = box / nx
    ax = ((np.arange(nx, dtype=np.float32) + np.float32(0.5)) * np.float32(dx)).astype(np.float32)
    last = ax if planes is None else ax[planes]
    kk = np.arange(nx) if planes is None else np.nonzero(planes)[0]
    if dim == 3:
        # utils.py:49-54 meshgrid(indexing="xy") ravel order
        X, Y, Z = np.meshgrid(ax, ax, last, indexing="xy")
        IX, IY, IK = np.meshgrid(np.arange(nx), np.arange(nx), kk, indexing="xy")
        ids = ((IY.ravel() * nx + IX.ravel()) * nx + IK.ravel()).astype(np.int32)
        r = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
        x, y, z = r[:, 0], r[:, 1], r[:, 2]
        u = np.stack([np.sin(x) * np.cos(y) * np.cos(z), -np.cos(x) * np.sin(y) * np.cos(z),
                      np.zeros_like(x)], axis=1).astype(np.float32)
        viscosity, u_ref = 0.02, 1.0
    else:
        X, Y = np.meshgrid(ax, last, indexing="xy")
        IX, IK = np.meshgrid(np.arange(nx), kk, indexing="xy")
        ids = (IK.ravel() * nx + IX.ravel()).astype(np.int32)
        r = np.stack([X.ravel(), Y.ravel()], axis=1)
        x, y = r[:, 0], r[:, 1]
        tp = np.float32(2 * np.pi)
        u = np.stack([-np.cos(tp * x) * np.sin(tp * y), np.sin(tp * x) * np.cos(tp * y)],
                     axis=1).astype(np.float32)
        viscosity, u_ref = 0.01, 1.0
    n = len(r)
    c_ref = 10.0 * u_ref
    # case_setup.py:94-97 (CFL 0.25)
    dt = float(min(0.25 * dx / (c_ref + u_ref), 0.25 * dx * dx / viscosity))
    ones = np.ones(n, dtype=np.float32)
    state = dict(r=np.ascontiguousarray(r, dtype=np.float32), u=u, v=u.copy(),
                 dudt=np.zeros_like(u), dvdt=np.zeros_like(u), rho=ones.copy(),
                 p=np.zeros(n, dtype=np.float32), drhodt=np.zeros(n, dtype=np.float32),
                 mass=ones * np.float32(dx**dim), eta=ones * np.float32(viscosity),
                 T=ones.copy(), dTdt=np.zeros(n, dtype=np.float32),
                 tag=np.zeros(n, dtype=np.int32))
    meta = dict(dim=dim, box=[box] * dim, dx=dx, dt=dt, viscosity=viscosity, c_ref=c_ref,
                p_ref=c_ref**2, tvf=1.0)
    if planes is not None:
        state["ids"] = ids
    return state, meta


def config_of(args, meta):
    from jax_sph_b200 import make_config

    dim = meta["dim"]
    return make_config(dim, meta["box"], meta["dx"], meta["dt"], tvf=meta["tvf"],
                       c_ref=meta["c_ref"], p_ref=meta["p_ref"],
                       cell_sub=[args.sub] * dim if args.sub else None,
                       threads=args.threads, list_cap=args.list_cap,
                       tile=[args.tile_x, args.tile_y, args.tile_z] if args.tile_x else None,
                       skin=args.skin,
                       **meta.get("cfg_kwargs", {}))


def workload_name(args, n):
    if args.workload == "ht3d":
        return (f"ht3d nx={args.nx} N={n} SPH bc_trick heat band-g QSK (BASELINE configs[4], "
                "cases/ht.yaml case.dim=3)")
    return f"{args.workload} nx={args.nx} N={n} SPH tvf=1 QSK (BASELINE configs[3])"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from jax_sph_b200 import Engine, make_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner through C stdio on fd 1 when the communicator is
        # created (eagerly, with device_id): keep stdout for the one JSON line
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        # NCCL's own streams at high priority: the halo exchanges run while the interior tile
        # layers of the next sweep occupy the SMs on the engine's side stream (slab overlap)
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
        return run_slab(args, world, rank, local, saved_stdout)

    state, meta = lattice_state(args.workload, args.nx)
    n = len(state["r"])
    dim = meta["dim"]
    cfg = config_of(args, meta)
    eng = Engine(cfg, n)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in state.items()}
    eng.upload(pinned)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.step(meta["dt"], args.warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    eng.step(meta["dt"], args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches() - l0
    clocks = sampler.finish()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    err = eng.error()
    counters = eng.counters()

    # per-pass CUDA-event times (same stream, separate loop of the same steps)
    eng.profile(True)
    acc = {}
    for _ in range(args.steps):
        eng.step(meta["dt"], 1)
        for k, v in eng.last_times().items():
            acc[k] = acc.get(k, 0.0) + v / args.steps
    eng.profile(False)

    # end to end through the host-buffer API, every step: Engine.advance_host = H2D of the state
    # entries this solver variant reads, advance, D2H of the entries it writes (the others pass
    # through advance() untouched and stay the caller's arrays, as in the reference)
    out = eng.download(host=True)
    host_in = {k: out[k] for k in out}
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    read, written = eng.live_fields()
    bytes_in = sum(host_in[k].numel() * host_in[k].element_size() for k in read)
    bytes_out = sum(host_in[k].numel() * host_in[k].element_size() for k in written)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_in = eng.advance_host(meta["dt"], host_in)  # download synchronises the stream
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        return
    peaks, which = measured_peaks()
    per_step_ms = ms / args.steps
    value = world * n * args.steps / (ms * 1e-3)
    edges = 93 if dim == 3 else 25  # directed in-range edges per lattice particle incl. self
    roof = roofline_of(acc, dim, n, edges, peaks, which)
    cpu = cpu_baseline(args, bounded=True)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n),
                   "particles_per_gpu": n, "parallelism": "1 engine per GPU" if world == 1 else
                   f"{world} independent periodic boxes (replicas, no halo exchange yet)",
                   "l2_policy": "state (>1.8 GB) larger than L2", "plan": eng.plan(),
                   "neighbour_search": dict(counters, timed_steps=args.steps, note=(
                       "steps / searches since engine creation (warm-up included): the cell sort "
                       "and candidate walk run when a particle has moved half the list skin; the "
                       "exact d^2 < cutoff^2 membership test runs for every pair on every step"))},
        "clocks": clocks, "gpu_launches": int(launches), "device_error_word": err,
        "e2e": {"value": world * n * e2e_steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(bytes_in), "d2h_bytes_per_step": int(bytes_out),
                "steps": e2e_steps,
                "what": "Engine.advance_host(dt, pinned host state) every step: H2D of the "
                        f"entries advance() reads ({','.join(read)}), one step, D2H of the "
                        f"entries it writes ({','.join(written)})"},
        "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def run_slab(args, world, rank, local, saved_stdout):
    """N > 1: the SAME box (strong scaling) cut into `world` slabs along the last axis, one
    rank per GPU, halo exchange + particle migration every step over NCCL (slab.py)."""
    import torch
    import torch.distributed as dist

    from jax_sph_b200 import SlabEngine, make_config
    from jax_sph_b200.engine import STATE_KEYS
    from jax_sph_b200.slab import layer_of

    meta = lattice_meta(args.workload, args.nx)
    dim, nx = meta["dim"], args.nx
    n_last = meta["nxyz"][-1] if "nxyz" in meta else nx  # lattice planes along the slab axis
    n_total = int(np.prod(meta["nxyz"])) if "nxyz" in meta else nx**dim
    cfg = config_of(args, meta)
    eng = SlabEngine(cfg)
    ax = ((np.arange(n_last, dtype=np.float32) + np.float32(0.5)) * np.float32(meta["dx"])).astype(np.float32)
    lay = layer_of(ax, eng.inv_cell, eng.layers)
    state, _ = lattice_state(args.workload, nx, planes=(lay >= eng.z0) & (lay < eng.z1))
    ids = state.pop("ids")
    n_own0 = len(ids)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in state.items()}
    ids_t = torch.from_numpy(ids).pin_memory()
    eng.upload(pinned, ids_t)
    torch.cuda.synchronize()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.step(meta["dt"], args.warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0, x0 = eng.launches(), eng.bytes_exchanged
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    eng.step(meta["dt"], args.steps)
    t_host = time.perf_counter() - t_host0
    e1.record()
    barrier()
    ms = reduce_max(e0.elapsed_time(e1))
    launches = eng.launches() - l0
    xbytes = (eng.bytes_exchanged - x0) / args.steps
    clocks = sampler.finish()
    err = eng.error()
    counts = eng.counts()
    tot = torch.tensor([counts["own"]], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)

    eng.profile(True)
    acc = {}
    for _ in range(args.steps):
        eng.step(meta["dt"], 1)
        for k, v in eng.last_times().items():
            acc[k] = acc.get(k, 0.0) + v / args.steps
    eng.profile(False)
    # stream time of each exchange (includes waiting for the neighbour to reach its send)
    eng.time_exchanges = True
    barrier()
    eng.step(meta["dt"], args.steps)
    xt = eng.exchange_times()
    eng.time_exchanges = False

    # end to end: this rank's slab host -> device, one advance (with its exchanges), device -> host
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    read, written = eng.live_fields()
    host_in, ids_h = eng.download([k for k in STATE_KEYS if k in read or k in written])
    bytes_in = sum(host_in[k].numel() * host_in[k].element_size() for k in read) + ids_h.numel() * 4
    bytes_out = sum(v.numel() * v.element_size() for v in host_in.values()) + ids_h.numel() * 4
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_in, ids_h = eng.advance_host(meta["dt"], host_in, ids_h)
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    bytes_all = torch.tensor([bytes_in, bytes_out], device="cuda", dtype=torch.int64)
    dist.all_reduce(bytes_all)
    barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    try:  # flush C stdio (the NCCL banner) while fd 1 still points at stderr
        import ctypes

        ctypes.CDLL(None).fflush(None)
    except Exception:
        pass
    os.dup2(saved_stdout, 1)

    if rank != 0:
        return
    peaks, which = measured_peaks()
    edges = 93 if dim == 3 else 25
    n_loc = counts["own"]
    roof = roofline_of(acc, dim, n_loc, edges, peaks, which)  # rank 0's slab
    roof["kernel"] = str(roof["kernel"]) + ", rank 0's slab"
    roof["traffic"] = None  # the ncu capture is of the single-GPU launch
    line = {
        "metric": METRIC, "value": n_total * args.steps / (ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n_total),
                   "particles_per_gpu": n_own0, "particles_total_after_run": int(tot.item()),
                   "parallelism": f"slab{world}: 1-D slabs of cell layers along axis {eng.axis}, "
                                  f"halo (one cutoff) + migration every step, NCCL send/recv ring",
                   "slab": {"layers": [eng.z0, eng.z1], "of": eng.layers, "own_cap": eng.own_cap,
                            "halo_cap": eng.halo_cap, "mig_cap": eng.mig_cap,
                            "message_bytes_per_step_per_rank": int(xbytes),
                            "host_enqueue_ms_per_step": t_host / args.steps * 1e3,
                            "exchange_ms_bytes_by_phase": {str(k): [round(v[0], 4), v[1]] for k, v in xt.items()},
                            "counts": counts},
                   "l2_policy": "per-rank state larger than L2" if n_own0 * 212 > 126e6 else
                                "per-rank state fits L2 (strong scaling of the named size)"},
        "clocks": clocks, "gpu_launches": int(launches), "device_error_word": err,
        "e2e": {"value": n_total * e2e_steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(bytes_all[0].item()), "d2h_bytes_per_step": int(bytes_all[1].item()),
                "steps": e2e_steps,
                "what": "per rank, every step: SlabEngine.advance_host(dt, pinned host slab) = H2D of the "
                        f"entries advance() reads ({','.join(read)}) + ids, one step with its ring "
                        "exchanges, D2H of the entries it reads or writes (the particle set changes "
                        "by migration) + ids"},
        "roofline": roof,
        "cpu_baseline": {"value": None, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "reported at N=1 only"},
    }
    print(json.dumps(line))


def cpu_workers():
    """How many single-core oracle workers the host takes: one per core, bounded by memory
    (a worker of the default sample peaks near 0.9 GB) and by 32."""
    cores = os.cpu_count() or 1
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable"))
        cores = min(cores, max(1, int(avail_kb / 1024 / 1024 * 0.5 / 1.5)))
    except Exception:
        pass
    return max(1, min(cores, 32))


def run_cpu_arm(args, steps, warmup=1):
    """The CPU implementation of the path on ALL host cores: one oracle worker per core
    (oracle/cpu_arm.py: the NumPy restatement of the reference's edge-list algorithm on an
    independent periodic box each), aggregate particle-updates/s over the common window
    [first worker's start, last worker's end]."""
    workers = cpu_workers()
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1",
               CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "oracle.cpu_arm", "--workload", args.workload, "--nx",
           str(args.cpu_nx), "--warmup", str(warmup), "--steps", str(steps)]
    procs = [subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, text=True)
             for _ in range(workers)]
    res = []
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("oracle.cpu_arm worker failed")
        res.append(json.loads(out.strip().splitlines()[-1]))
    n = res[0]["n"]
    window = max(r["t1"] for r in res) - min(r["t0"] for r in res)
    value = workers * n * steps / window
    per_core = float(np.mean([r["n"] * r["steps"] / (r["t1"] - r["t0"]) for r in res]))
    return dict(value=value, workers=workers, n=n, window_s=window, per_core=per_core)


def cpu_baseline(args, bounded=True):
    """The NumPy oracle (port of the reference algorithm) on the host cores, on a
    bounded sample of the same workload."""
    r = run_cpu_arm(args, args.cpu_steps)
    return {"value": r["value"], "unit": UNIT, "cores": r["workers"], "kind": "port",
            "sample": f"{args.workload} nx={args.cpu_nx} N={r['n']} per worker, {args.cpu_steps} "
                      f"steps, {r['workers']} single-core workers (one periodic box each) of the "
                      "NumPy restatement of the reference edge-list algorithm (jax is not "
                      f"installable here), {r['window_s']:.1f} s wall, "
                      f"{r['per_core']:.0f} particle-updates/s per worker, host has "
                      f"{os.cpu_count()} cores"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sub = max(1, args.cpu_steps // 4)
    r = run_cpu_arm(args, sub * args.steps, warmup=min(args.warmup, 1))
    value = r["value"]
    sample = (f"{args.workload} nx={args.cpu_nx} N={r['n']} per worker: each bench step = {sub} "
              f"advance() calls on each of {r['workers']} single-core workers (one periodic box "
              "each) of the NumPy port of the reference algorithm (jax/jaxlib absent, reference "
              f"not runnable); {r['per_core']:.0f} particle-updates/s per worker")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["window_s"] / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} (bounded sample nx={args.cpu_nx} N={r['n']} x "
                               f"{r['workers']} workers)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["workers"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tgv3d", choices=["tgv3d", "tgv2d", "ht3d"])
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--cpu-nx", type=int, default=32)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--sub", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--list-cap", type=int, default=0)
    ap.add_argument("--tile-x", type=int, default=0)
    ap.add_argument("--tile-y", type=int, default=0)
    ap.add_argument("--tile-z", type=int, default=0)
    ap.add_argument("--skin", type=float, default=0.0,
                    help="neighbour-list skin / cutoff (0 = engine default, < 0 = search every step)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
