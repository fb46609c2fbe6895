"""bench.py -- particle-updates/s of the JAX-SPH per-step hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Workload at N=1: BASELINE.json configs[3], the 3D Taylor-Green vortex of
validation/tgv3d.sh (SPH, tvf=1, viscosity 0.02, Quintic kernel) on a 256^3
lattice = 16 777 216 particles, float32, synthetic lattice initialisation.
One "step" = one advance(dt) (integrator.py:22-56): kick + drift + wrap, cell
sort + neighbour search (on the steps where a particle has moved half the list
skin since the last one), exact membership test + density sweep + EoS, force
sweep.

Prints ONE JSON line (rank 0).  `value` times the resident engine (state in
HBM); `e2e` times the same step through the public host-buffer API with the
state copied host->device before and device->host after every step;
`stateless_advance` times the C-ABI call an XLA custom call would make
(sphb200_advance: device pointers in, device pointers out, caller workspace);
`configs` holds short runs of the other single-GPU BASELINE configurations
(C1, C2a, C2b, C3); `roofline` puts the dominant kernel against the FP32 peak
measured on this device by the library's own FMA kernel.
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

# Algorithmic HBM bytes per particle-step, float32: what the pass must
# move if every array crossed HBM once -- integrate 144 (R pt um vv du dv rb, W pt um vv),
# density 24 (R pt, W rho p), force 76 (R pt um vv st, W dudt dvdt).  `lists` = what the engine
# adds on top by keeping neighbour lists in HBM, per particle: the density
# pass reads the skin row of its duo (about 190 / 45 entries per TWO particles on the 3D / 2D
# lattice, 2 B each) and writes the exact row (about 120 / 32 per two particles), the force pass
# reads the exact row and the compact force records (52 B written + read per particle).
BYTES = {
    3: dict(cells=144, density=24, force=76, lists=dict(density=190 + 2 + 120 + 2, force=120 + 2 + 104)),
    2: dict(cells=144, density=24, force=76, lists=dict(density=45 + 2 + 32 + 2, force=32 + 2 + 104)),
}
# useful flops per directed in-range edge, SURVEY.md section 8d
FLOPS_EDGE = {3: dict(density=58 + 1, force=58 + 103 + 16), 2: dict(density=50 + 1, force=50 + 63 + 14)}
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # nominal non-tensor FP32 FMA peak


def source_sha():
    """Hash of the CUDA sources: ties a committed ncu capture to the build it describes."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "jax_sph_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the sweep from the committed
    ncu capture (profiles/r02_traffic.json: bytes per particle, scaled to this launch).  None
    unless the capture was taken from THESE sources (source_sha) -- a stale file says nothing."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("source_sha") != source_sha():
            return None
        return t[kernel]["dram_bytes_per_particle"] * n
    except Exception:
        return None


def fp32_peak():
    """Measured FP32 FMA throughput of this device (sphb200_fp32_peak: independent FFMA chains,
    no memory traffic), TFLOP/s; the denominator of the sweeps' roofline."""
    import ctypes as C

    from jax_sph_b200 import _lib

    lib = _lib.load()
    out = {}
    for name, packed in (("ffma", 0), ("ffma2", 1)):
        tf, ms = C.c_double(), C.c_double()
        _lib.check(lib.sphb200_fp32_peak(packed, C.byref(tf), C.byref(ms), None))
        out[name] = tf.value
    return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def roofline_of(acc, dim, n, edges, peaks, which, fp32=None, searches_per_step=None, duo=True):
    """`roofline` object: the DOMINANT kernel of the step (largest CUDA-event time) against the
    roofline that bounds it -- FP32 issue: the sweeps do 20 useful flop per algorithmic byte, the
    ridge of this device is at 11 -- with the measured FMA peak as the
    denominator; its HBM figure, the other sweep and the HBM-bound integrate pass beside it."""
    names = {"density": "k_duo<PhysDensity, DUO_FILTER> (exact membership test + density sweep"
                        " + exact list of the step; on sorting steps preceded by the search "
                        "k_duo<PhysNone, DUO_BUILD>, included in ms)",
             "force": "k_duo<PhysForce, DUO_CONSUME, 2> (force sweep; k_force_rec, the compact "
                      "records it stages by bulk copy, included in ms)"}
    if not duo:
        names = {"density": "k_sweep<PhysDensity, LIST_FILTER> (exact membership test + density "
                            "sweep + exact list of the step; on sorting steps preceded by the "
                            "search k_sweep<PhysNone, LIST_BUILD>, included in ms)",
                 "force": "k_sweep<PhysForce, LIST_CONSUME> (force sweep)"}
    peak_tf = fp32["ffma"] if fp32 else FP32_NOMINAL_TFLOPS
    peak_src = "measured: sphb200_fp32_peak (FFMA chains) on this device" if fp32 else "nominal"
    per = {}
    for k in ("density", "force"):
        ms = acc.get(k, 0.0)
        if not ms:
            continue
        gbs = BYTES[dim][k] * n / (ms * 1e-3) / 1e9
        gbs_l = (BYTES[dim][k] + BYTES[dim]["lists"][k]) * n / (ms * 1e-3) / 1e9
        tf = FLOPS_EDGE[dim][k] * edges * n / (ms * 1e-3) / 1e12
        per[k] = {"kernel": names[k], "ms": ms, "bound": "fp32", "achieved": tf, "peak": peak_tf,
                  "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": ncu_traffic(k, n),
                  "hbm": {"algorithmic_bytes_per_particle": BYTES[dim][k], "achieved_gbs": gbs,
                          "with_list_rows_gbs": gbs_l, "peak_gbs": peaks["hbm_gbs"],
                          "frac": gbs / peaks["hbm_gbs"]}}
    if not per:
        return {"kernel": None, "bound": "fp32", "achieved": None, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": None, "traffic": None, "passes_ms": acc}
    top = max(per, key=lambda k: per[k]["ms"])
    c_ms = acc.get("cells", 0.0)
    tot_ms = acc.get("total", 0.0)
    step_tf = sum(FLOPS_EDGE[dim][k] for k in ("density", "force")) * edges * n / (tot_ms * 1e-3) / 1e12 \
        if tot_ms else None
    roof = dict(per[top])
    roof.update({
        "peak_source": peak_src, "fp32_peaks_tflops": dict(fp32 or {}, nominal=FP32_NOMINAL_TFLOPS),
        "hbm_peak_source": which,
        "note": "achieved = useful flops (SURVEY.md 8d: per in-range directed edge, "
                f"{edges} edges per lattice particle) / CUDA-event time of the pass; the sweeps "
                "are bound by the FP32 pipes (the packed f32x2 instructions halve the issue "
                "slots, not the pipe cycles) and by latency at one block per SM (ncu: profiles/), "
                "not by HBM; `cells` is the HBM-bound integrate pass (plus the cell sort on the "
                "steps that sort)",
        "kernels": per,
        "step": {"ms": tot_ms, "useful_tflops": step_tf,
                 "frac_of_fp32_peak": step_tf / peak_tf if step_tf else None,
                 "searches_per_step": searches_per_step},
        "cells": {"kernel": "k_drift (kick + drift + wrap in place, re-sort decision) [+ k_hash, "
                            "scan, k_scatter_src, k_reorder, k_copyback on the steps that sort]",
                  "ms": c_ms, "bound": "hbm",
                  "achieved": BYTES[dim]["cells"] * n / (c_ms * 1e-3) / 1e9 if c_ms else None,
                  "peak": peaks["hbm_gbs"], "unit": "GB/s",
                  "frac": BYTES[dim]["cells"] * n / (c_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]
                  if c_ms else None},
        "passes_ms": acc,
    })
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


def db2d_meta(dx):
    """BASELINE configs[2]: cases/db.yaml (2D dam break, cases/db.py:17-142): fluid column
    L x H = 2 x 1 at rest in a tank L_wall x H_wall = 5.366 x 2 with three wall layers, gravity
    -e_y, SPH + is_bc_trick + density evolution + artificial viscosity 0.1, u_ref = sqrt(2),
    viscosity 5e-5; dx = 0.00071 gives 4 028 622 particles."""
    n_walls, g_mag, viscosity, u_ref = 3, 1.0, 0.00005, 2.0**0.5
    c_ref = 10.0 * u_ref
    eps = float(np.finfo(np.float32).eps)
    dt = float(min(0.25 * dx / (c_ref + u_ref), 0.25 * dx * dx / (viscosity + eps),
                   0.25 * (dx / (g_mag + eps)) ** 0.5))  # case_setup.py:94-97
    dxn = n_walls * dx
    box = [5.366 + 2 * dxn + 0.1, 2.0 + 2 * dxn + 0.1]
    zero = [0.0, 0.0, 0.0]
    bc_table = {"tags": {1: dict(u=zero, v=zero, zero_dudt=True, zero_dvdt=True, p=0.0)},
                "inflow_x": None, "outflow_x": None}
    return dict(dim=2, box=box, dx=dx, dt=dt, viscosity=viscosity, c_ref=c_ref, p_ref=c_ref**2,
                tvf=0.0, n_walls=n_walls,
                cfg_kwargs=dict(is_bc_trick=True, is_rho_evol=True, artificial_alpha=0.1,
                                g_ext_spec={"mode": "const", "g": [0.0, -g_mag, 0.0]},
                                bc_table=bc_table))


def db2d_state(dx):
    """Particles of db2d_meta(dx) as cases/db.py:94-127 lays them out: the four wall blocks of
    pos_box_2d (utils.py:57-78: left, bottom, right, top; lattices (i + 0.5) dx), then the fluid
    lattice shifted by the wall thickness; float32 arithmetic as upstream."""
    meta = db2d_meta(dx)
    t = np.float32

    def init2(bx, by):  # pos_init_cartesian_2d, utils.py:35-46
        n0, n1 = int(np.round(bx / dx)), int(np.round(by / dx))
        gx, gy = np.meshgrid(range(n0), range(n1), indexing="xy")
        return ((np.vstack([gx.ravel(), gy.ravel()]).T.astype(np.float32) + t(0.5)) * t(dx)).astype(np.float32)

    dxn = meta["n_walls"] * dx
    lw, hw = 5.366, 2.0
    vertical, horiz = init2(dxn, hw + 2 * dxn), init2(lw, dxn)
    walls = np.concatenate([vertical, horiz + np.array([dxn, 0.0]),
                            vertical + np.array([lw + dxn, 0.0]),
                            horiz + np.array([dxn, hw + dxn])]).astype(np.float32)
    fluid = (t(dxn) + init2(2.0, 1.0)).astype(np.float32)
    r = np.concatenate([walls, fluid]).astype(np.float32)
    tag = np.concatenate([np.full(len(walls), 1), np.full(len(fluid), 0)]).astype(np.int32)
    n = len(r)
    ones, zv = np.ones(n, dtype=np.float32), np.zeros((n, 2), dtype=np.float32)
    state = dict(r=r, u=zv, v=zv.copy(), dudt=zv.copy(), dvdt=zv.copy(), rho=ones.copy(),
                 p=np.zeros(n, dtype=np.float32), drhodt=np.zeros(n, dtype=np.float32),
                 mass=ones * t(dx**2), eta=ones * t(meta["viscosity"]), T=ones.copy(),
                 dTdt=np.zeros(n, dtype=np.float32), tag=tag)
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
    if workload == "db2d":
        return db2d_state(1.0 / nx if nx else 0.00071)  # nx = 0: BASELINE's dx, 4 028 622 particles
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
    if workload == "tgv2d_sph":  # BASELINE configs[0]: solver.name=SPH solver.tvf=0.0
        meta["tvf"] = 0.0
    if workload == "tgv2d_rie":  # BASELINE configs[1]: solver.name=RIE (+ density evolution)
        meta["tvf"] = 0.0
        meta["cfg_kwargs"] = dict(solver="RIE", is_rho_evol=True)
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
                       # every generated state has one eta (verified on the device every step)
                       uniform_eta=not args.no_uniform_eta,
                       **meta.get("cfg_kwargs", {}))


def workload_name(args, n):
    if args.workload == "ht3d":
        return (f"ht3d nx={args.nx} N={n} SPH bc_trick heat band-g QSK (BASELINE configs[4], "
                "cases/ht.yaml case.dim=3)")
    what = {"tgv3d": "SPH tvf=1 QSK (BASELINE configs[3])",
            "tgv2d": "SPH tvf=1 QSK (BASELINE configs[1], transport-velocity SPH)",
            "tgv2d_sph": "SPH tvf=0 QSK (BASELINE configs[0], cases/tgv.yaml)",
            "tgv2d_rie": "RIE + density evolution QSK (BASELINE configs[1], Riemann SPH)",
            "db2d": "SPH bc_trick density-evolution alpha=0.1 gravity QSK (BASELINE configs[2], "
                    "cases/db.yaml)"}[args.workload]
    return f"{args.workload} nx={args.nx} N={n} {what}"


# the other single-GPU BASELINE configurations, run for a few steps after the headline
OTHER_CONFIGS = (("C1", "tgv2d_sph", 50), ("C2a", "tgv2d", 1000), ("C2b", "tgv2d_rie", 1000),
                 ("C3", "db2d", 0))


def run_config(args, workload, nx, steps=20, warmup=3):
    """particle-updates/s of the resident engine on another BASELINE configuration."""
    import copy

    import torch

    from jax_sph_b200 import Engine

    a = copy.copy(args)
    a.workload, a.nx = workload, nx
    a.sub = a.threads = a.list_cap = a.tile_x = a.tile_y = a.tile_z = 0
    state, meta = lattice_state(workload, nx)
    n = len(state["r"])
    eng = Engine(config_of(a, meta), n)
    eng.upload({k: torch.from_numpy(v) for k, v in state.items()})
    eng.step(meta["dt"], warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(meta["dt"], steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    err = eng.error()
    cnt = eng.counters()
    eng.close()
    return {"workload": workload_name(a, n), "n": n, "steps": steps, "ms_per_step": ms / steps,
            "value": n * steps / (ms * 1e-3), "unit": UNIT, "device_error_word": err,
            "searches": cnt["searches"], "steps_total": cnt["steps"],
            "tiles_without_lists": cnt["tiles_without_lists"]}


def stateless_advance(args, eng_cfg, state, meta, steps):
    """The drop-in call: sphb200_advance(cfg, n, dt, in, out, err, workspace, stream) on DEVICE
    pointers with a caller-owned workspace -- what the jax.ffi custom call of INTEGRATION.md
    executes once per `advance(dt, state, neighbors)` (jax_sph/simulate.py:117).  Outputs of one
    call are the inputs of the next (two sets of arrays, swapped), the workspace is the same."""
    import ctypes as C

    import torch

    from jax_sph_b200 import _lib
    from jax_sph_b200.engine import STATE_KEYS

    lib = _lib.load()
    n = len(state["r"])
    nbytes = C.c_size_t()
    _lib.check(lib.sphb200_workspace_bytes(C.byref(eng_cfg), n, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    bufs = [{k: torch.from_numpy(np.ascontiguousarray(state[k])).cuda() for k in STATE_KEYS
             if k in state} for _ in range(2)]
    err = torch.zeros(1, dtype=torch.int32, device="cuda")

    def struct(d):
        st = _lib.State()
        for k, v in d.items():
            setattr(st, k, v.data_ptr())
        return st

    sts = [struct(b) for b in bufs]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timed(fn, count):
        i0 = timed.i
        for i in range(i0, i0 + 3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(i0 + 3, i0 + 3 + count):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        timed.i = i0 + 3 + count
        return e0.elapsed_time(e1) / count

    timed.i = 0

    def call(fn):
        return lambda i: _lib.check(fn(C.byref(eng_cfg), n, float(meta["dt"]), C.byref(sts[i % 2]),
                                       C.byref(sts[(i + 1) % 2]), C.c_void_p(err.data_ptr()),
                                       C.c_void_p(ws.data_ptr()), nbytes.value, stream))

    ms = timed(call(lib.sphb200_advance_persistent), steps)
    lib.sphb200_workspace_release(C.c_void_p(ws.data_ptr()))
    # the same with the arrays kept in engine order (sphb200_advance_ordered): outputs and order of
    # one call are the inputs of the next
    order = [torch.arange(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")]

    def call_ordered(i):
        _lib.check(lib.sphb200_advance_ordered(
            C.byref(eng_cfg), n, float(meta["dt"]), C.byref(sts[i % 2]), C.c_void_p(order[i % 2].data_ptr()),
            C.byref(sts[(i + 1) % 2]), C.c_void_p(order[(i + 1) % 2].data_ptr()), C.c_void_p(err.data_ptr()),
            C.c_void_p(ws.data_ptr()), nbytes.value, stream))

    timed.i = 0
    ms_ordered = timed(call_ordered, steps)
    err_ordered = int(err.item())
    lib.sphb200_workspace_release(C.c_void_p(ws.data_ptr()))
    ms_scratch = timed(call(lib.sphb200_advance), min(steps, 3))
    return {"value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "device_error_word": int(err.item()),
            "what": "sphb200_advance_persistent on device pointers, the caller's workspace handed "
                    "in again call after call (the engine in it keeps the particles cell-sorted "
                    "and the neighbour lists; every call copies the full state in and the full "
                    "state out in the caller's particle order)",
            "engine_order": {"value": n / (ms_ordered * 1e-3), "unit": UNIT, "ms_per_step": ms_ordered,
                             "steps": steps, "device_error_word": err_ordered,
                             "what": "sphb200_advance_ordered: the caller keeps its arrays in the engine's "
                                     "slot order and an int32 order array carries the particle labels "
                                     "through the sorts; every call streams the full state (16 entries) "
                                     "in and out without a permutation"},
            "scratch_workspace": {"value": n / (ms_scratch * 1e-3), "unit": UNIT,
                                  "ms_per_step": ms_scratch,
                                  "what": "sphb200_advance: nothing kept between calls (pack, "
                                          "sort, search, step, unpack every call)"}}


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
    fp32 = fp32_peak()

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
    roof = roofline_of(acc, dim, n, edges, peaks, which, fp32,
                       counters["searches"] / max(counters["steps"], 1), counters.get("duo", False))
    # the C-ABI drop-in call and the other BASELINE configurations (resident engine freed first)
    cfg_copy = config_of(args, meta)
    plan = eng.plan()
    eng.close()
    del eng
    torch.cuda.empty_cache()
    stateless = stateless_advance(args, cfg_copy, state, meta, max(1, min(args.steps, 10)))
    others = {}
    if not args.no_configs:
        for tag, wl, nx_c in OTHER_CONFIGS:
            try:
                others[tag] = run_config(args, wl, nx_c)
            except Exception as ex:  # a failed side run must not hide the headline
                others[tag] = {"error": repr(ex)}
    cpu = cpu_baseline(args, bounded=True)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n),
                   "particles_per_gpu": n, "parallelism": "1 engine on 1 GPU (N > 1: the same box "
                   "cut into slabs, strong scaling)",
                   "l2_policy": "state (>1.8 GB) larger than L2", "plan": plan,
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
        "stateless_advance": stateless, "configs": others,
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
    roof = roofline_of(acc, dim, n_loc, edges, peaks, which, fp32_peak())  # rank 0's slab
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
                                  f"halo (one cutoff) + migration every step, transport: " +
                                  ("pack kernels store into the ring neighbours' symmetric-memory buffers "
                                   "over NVLink + signal flags (no collective on the data path)"
                                   if eng.transport == "direct" else "NCCL send/recv ring + all_reduce"),
                   "transport": eng.transport,
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
    ap.add_argument("--workload", default="tgv3d",
                    choices=["tgv3d", "tgv2d", "tgv2d_sph", "tgv2d_rie", "db2d", "ht3d"])
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the short runs of the other BASELINE configurations")
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
    ap.add_argument("--no-uniform-eta", action="store_true",
                    help="do not promise a uniform viscosity (SPHB200_HINT_UNIFORM_ETA off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
