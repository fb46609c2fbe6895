"""Golden vectors produced by the UNMODIFIED reference sources.

    python tests/golden/make_reference_golden.py            # all cases, both dtypes
    python tests/golden/make_reference_golden.py --only tgv2d_sph
    python tests/golden/make_reference_golden.py --long     # 200-step trajectories

jax / jaxlib cannot be installed in the build container (no network, not in the
wheelhouse), so the reference package under /root/reference is imported as it
lies there and executed against `tests/golden/jaxshim/` -- a torch-CPU
stand-in for the slice of the jax API the reference uses (`jax.numpy`, `vmap`,
`grad`, `ops.segment_sum`, `lax.cond`, ...) plus import stubs for omegaconf /
jraph / h5py / pyvista.  Every line of physics, case setup, neighbour search
and time integration that runs is the reference's own: `main.load_embedded_configs`
(main.py:20-47), `SimulationSetup.initialize` (jax_sph/case_setup.py:44-231),
`partition.neighbor_list` (jax_sph/partition.py:492-571 -> jax_md/partition.py),
`WCSPH.forward_wrapper` (jax_sph/solver.py:702-951), `si_euler`
(jax_sph/integrator.py:8-58).  The loop below restates only the wiring of
`simulate()` (jax_sph/simulate.py:28-134) because that function writes files
and does not return the state.

What differs from a real jax run: the floating-point executor (torch CPU
kernels instead of XLA:CPU -- same IEEE operations, possibly different
summation order inside segment_sum) and `jax.random` (the noise added to the
initial lattice comes from a torch generator; it is only ever used as an
INPUT here: state0 is stored and fed to the oracle and to the engine).

Each tests/golden/ref_<case>.npz holds: the config scalars, state0 (the state
returned by `initialize()`), the reference neighbour list of state0 as
canonical sorted (sender, receiver) pairs, the state after `advance(0.0)`
(= WCSPH.forward + bc_fn) and after NSTEPS `advance(dt)` calls, in float32
and in float64 (x64 enabled before the reference is imported, as main.py:66
does).
"""

import argparse
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
NSTEPS = 20
OUT_KEYS = ("r", "u", "v", "rho", "p", "dudt", "dvdt", "drhodt", "T", "dTdt")

# name -> (reference CLI overrides as for main.py, oracle.cases.make_case kwargs)
CASES = {
    # BASELINE configs[0]: cases/tgv.yaml solver.name=SPH solver.tvf=0.0 (2 500 particles)
    "tgv2d_sph": (["config=cases/tgv.yaml", "solver.name=SPH", "solver.tvf=0.0"],
                  dict(case="tgv", dim=2, dx=0.02)),
    "tgv2d_tvf": (["config=cases/tgv.yaml", "solver.tvf=1.0", "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=2, dx=0.02, tvf=1.0)),
    "tgv2d_rie": (["config=cases/tgv.yaml", "solver.name=RIE", "solver.density_evolution=True",
                   "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=2, dx=0.02, solver="RIE", density_evolution=True)),
    # validation/tgv3d.sh:20 at nx=16
    "tgv3d_tvf": (["config=cases/tgv.yaml", "case.dim=3", "case.dx=0.39269908169872414",
                   "case.viscosity=0.02", "solver.tvf=1.0", "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=3, dx=0.39269908169872414, tvf=1.0, viscosity=0.02)),
    "tgv3d_rie": (["config=cases/tgv.yaml", "case.dim=3", "case.dx=0.5235987755982988",
                   "case.viscosity=0.02", "solver.name=RIE", "solver.density_evolution=True"],
                  dict(case="tgv", dim=3, dx=0.5235987755982988, viscosity=0.02, solver="RIE",
                       density_evolution=True)),
    "db2d": (["config=cases/db.yaml", "case.dx=0.04", "solver.dt=null"],
             dict(case="db", dim=2, dx=0.04)),
    "db2d_renorm": (["config=cases/db.yaml", "case.dx=0.05", "solver.dt=null",
                     "solver.density_renormalize=True"],
                    dict(case="db", dim=2, dx=0.05, density_renormalize=True)),
    "db2d_rie": (["config=cases/db.yaml", "case.dx=0.05", "solver.dt=null", "solver.name=RIE",
                  "solver.artificial_alpha=0.0"],
                 dict(case="db", dim=2, dx=0.05, solver="RIE", artificial_alpha=0.0)),
    "ht2d": (["config=cases/ht.yaml"], dict(case="ht", dim=2, dx=0.02)),
    "ht3d": (["config=cases/ht.yaml", "case.dim=3", "case.dx=0.04"],
             dict(case="ht", dim=3, dx=0.04)),
    "cf2d_wc2k": (["config=cases/cf.yaml", "case.dx=0.04", "solver.dt=null", "kernel.name=WC2K",
                   "kernel.h_factor=1.3"],
                  dict(case="cf", dim=2, dx=0.04, kernel="WC2K", h_factor=1.3)),
    "cf2d_freeslip": (["config=cases/cf.yaml", "case.dx=0.04", "solver.dt=null",
                       "solver.free_slip=True"],
                      dict(case="cf", dim=2, dx=0.04, free_slip=True)),
    "pf2d": (["config=cases/pf.yaml", "case.dx=0.04", "solver.dt=null"],
             dict(case="pf", dim=2, dx=0.04)),
    # Delta-SPH: reference tests/test_pf2d.py:107, test_cf2d.py:111
    # (summation density + velocity diffusion) and validation/db2d.sh:8 (density diffusion,
    # gamma = 7)
    "pf2d_delta": (["config=cases/pf.yaml", "case.dx=0.04", "solver.dt=null", "solver.name=DELTA"],
                   dict(case="pf", dim=2, dx=0.04, solver="DELTA")),
    "cf2d_delta": (["config=cases/cf.yaml", "case.dx=0.04", "solver.dt=null", "solver.name=DELTA"],
                   dict(case="cf", dim=2, dx=0.04, solver="DELTA")),
    "db2d_delta": (["config=cases/db.yaml", "case.dx=0.05", "solver.dt=null", "solver.name=DELTA",
                    "eos.gamma=7.0", "solver.artificial_alpha=0.0"],
                   dict(case="db", dim=2, dx=0.05, solver="DELTA", gamma=7.0, artificial_alpha=0.0)),
    # the five kernels outside QSK / WC2K (kernel.py:26-48, :106-201)
    "tgv2d_csk": (["config=cases/tgv.yaml", "case.dx=0.04", "solver.tvf=1.0", "kernel.name=CSK",
                   "kernel.h_factor=1.3", "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=2, dx=0.04, tvf=1.0, kernel="CSK", h_factor=1.3)),
    "tgv2d_wc4k": (["config=cases/tgv.yaml", "case.dx=0.04", "solver.tvf=1.0", "kernel.name=WC4K",
                    "kernel.h_factor=1.5", "case.r0_noise_factor=0.25"],
                   dict(case="tgv", dim=2, dx=0.04, tvf=1.0, kernel="WC4K", h_factor=1.5)),
    "tgv3d_wc6k": (["config=cases/tgv.yaml", "case.dim=3", "case.dx=0.6283185307179586",
                    "case.viscosity=0.02", "solver.tvf=1.0", "kernel.name=WC6K",
                    "kernel.h_factor=1.5", "case.r0_noise_factor=0.25"],
                   dict(case="tgv", dim=3, dx=0.6283185307179586, viscosity=0.02, tvf=1.0,
                        kernel="WC6K", h_factor=1.5)),
    "tgv2d_gk": (["config=cases/tgv.yaml", "case.dx=0.04", "solver.name=RIE",
                  "solver.density_evolution=True", "kernel.name=GK", "case.r0_noise_factor=0.25"],
                 dict(case="tgv", dim=2, dx=0.04, solver="RIE", density_evolution=True, kernel="GK")),
    "db2d_sgk": (["config=cases/db.yaml", "case.dx=0.05", "solver.dt=null", "kernel.name=SGK"],
                 dict(case="db", dim=2, dx=0.05, kernel="SGK")),
    "tgv3d_delta": (["config=cases/tgv.yaml", "case.dim=3", "case.dx=0.6283185307179586",
                     "case.viscosity=0.02", "solver.name=DELTA", "solver.density_evolution=True",
                     "case.r0_noise_factor=0.25"],
                    dict(case="tgv", dim=3, dx=0.6283185307179586, viscosity=0.02, solver="DELTA",
                         density_evolution=True)),
    "tgv2d_delta": (["config=cases/tgv.yaml", "case.dx=0.04", "solver.name=DELTA",
                     "solver.density_evolution=True", "case.r0_noise_factor=0.25"],
                    dict(case="tgv", dim=2, dx=0.04, solver="DELTA", density_evolution=True)),
}


# north_star: "tolerance-matched 200-step trajectories for every listed case".  The same cases at
# sizes the stand-in steps through in minutes; only state0 and the final state are stored
# (tests/golden/ref200_<case>.npz).
LONG_STEPS = 200
LONG_CASES = {
    "tgv2d_sph": (["config=cases/tgv.yaml", "solver.name=SPH", "solver.tvf=0.0", "case.dx=0.04",
                   "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=2, dx=0.04)),
    "tgv2d_tvf": (["config=cases/tgv.yaml", "solver.tvf=1.0", "case.dx=0.04",
                   "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=2, dx=0.04, tvf=1.0)),
    "tgv2d_rie": (["config=cases/tgv.yaml", "solver.name=RIE", "solver.density_evolution=True",
                   "case.dx=0.04", "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=2, dx=0.04, solver="RIE", density_evolution=True)),
    "tgv3d_tvf": (["config=cases/tgv.yaml", "case.dim=3", "case.dx=0.6283185307179586",
                   "case.viscosity=0.02", "solver.tvf=1.0", "case.r0_noise_factor=0.25"],
                  dict(case="tgv", dim=3, dx=0.6283185307179586, tvf=1.0, viscosity=0.02)),
    "db2d": (["config=cases/db.yaml", "case.dx=0.08", "solver.dt=null"],
             dict(case="db", dim=2, dx=0.08)),
    "ht2d": (["config=cases/ht.yaml", "case.dx=0.04"], dict(case="ht", dim=2, dx=0.04)),
    "cf2d": (["config=cases/cf.yaml", "case.dx=0.05", "solver.dt=null"],
             dict(case="cf", dim=2, dx=0.05)),
    # 3D at sizes where the engine's interior tiles exist (24 cells per axis in the periodic box)
    "tgv3d_tvf24": (["config=cases/tgv.yaml", "case.dim=3", "case.dx=0.2617993877991494",
                     "case.viscosity=0.02", "solver.tvf=1.0", "case.r0_noise_factor=0.25"],
                    dict(case="tgv", dim=3, dx=0.2617993877991494, tvf=1.0, viscosity=0.02)),
    "ht3d": (["config=cases/ht.yaml", "case.dim=3", "case.dx=0.04"],
             dict(case="ht", dim=3, dx=0.04)),
}


def canonical_pairs(idx, n):
    """(2, E) padded edge list -> sorted int64 keys sender * n + receiver (padding dropped)."""
    recv, send = np.asarray(idx[0], dtype=np.int64), np.asarray(idx[1], dtype=np.int64)
    ok = (recv < n) & (send < n)
    return np.sort(send[ok] * n + recv[ok])


def pack_pairs(keys, n):
    """Sorted pair keys -> (per-sender counts int32, receivers in sender-major order uint16/int32)."""
    send, recv = keys // n, keys % n
    counts = np.bincount(send, minlength=n).astype(np.int32)
    return counts, recv.astype(np.uint16 if n <= 65535 else np.int32)


def unpack_pairs(counts, recv):
    """Inverse of pack_pairs: sorted int64 keys sender * n + receiver."""
    n = len(counts)
    send = np.repeat(np.arange(n, dtype=np.int64), counts)
    return send * n + recv.astype(np.int64)


def run_kernels(x64):
    """w(r) and grad_w(r) of every kernel class of jax_sph/kernel.py (vmapped as solver.py:726,730
    does) on a grid that includes r = 0, the knots and the cutoff: tests/golden/kernel_tables.npz."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    sys.path.insert(0, REF)
    import jax

    jax.config.update("jax_enable_x64", bool(x64))
    import jax.numpy as jnp

    from jax_sph import kernel as K

    tag = "f64" if x64 else "f32"
    h = 0.0123
    out = {"h": np.float64(h)}
    classes = dict(CSK=K.CubicKernel, QSK=K.QuinticKernel, WC2K=K.WendlandC2Kernel,
                   WC4K=K.WendlandC4Kernel, WC6K=K.WendlandC6Kernel, GK=K.GaussianKernel,
                   SGK=K.SuperGaussianKernel)
    for name, cls in classes.items():
        for dim in (2, 3):
            k = cls(h=h, dim=dim)
            q = np.concatenate([np.linspace(0.0, 3.3, 265), np.array([1.0, 2.0, 3.0])])
            r = jnp.array(np.sort(q) * h)
            out[f"{name}_{dim}_{tag}_r"] = np.array(r)
            out[f"{name}_{dim}_{tag}_w"] = np.array(jax.vmap(k.w)(r))
            out[f"{name}_{dim}_{tag}_gw"] = np.array(jax.vmap(k.grad_w)(r))
            out[f"{name}_{dim}_cutoff"] = np.float64(k.cutoff)
    np.savez(os.path.join(HERE, f"_tmp_kernels_{tag}.npz"), **out)


def run_case(name, x64, long=False):
    """Executed in a child process: one case, one dtype, through the reference's own code."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    sys.path.insert(0, REF)
    os.chdir(REF)
    import jax

    jax.config.update("jax_enable_x64", bool(x64))  # main.py:65-66, before jax_sph is imported
    import main as refmain
    from omegaconf import OmegaConf

    from jax_sph import partition
    from jax_sph.case_setup import load_case
    from jax_sph.integrator import si_euler
    from jax_sph.jax_md.partition import Sparse
    from jax_sph.solver import WCSPH
    from jax_sph.utils import Tag

    cli, _ = (LONG_CASES if long else CASES)[name]
    nsteps = LONG_STEPS if long else NSTEPS
    cli_args = OmegaConf.from_dotlist(cli + ["dtype=" + ("float64" if x64 else "float32")])
    cfg = refmain.load_embedded_configs(cli_args)

    # jax_sph/simulate.py:28-93, verbatim wiring
    Case = load_case(os.path.dirname(cfg.config), cfg.case.source)
    case = Case(cfg)
    (cfg, box_size, state, g_ext_fn, bc_fn, nw_fn, eos_fn, key, displacement_fn,
     shift_fn) = case.initialize()
    solver = WCSPH(
        displacement_fn, eos_fn, g_ext_fn, cfg.case.dx, cfg.case.dim, cfg.solver.dt,
        cfg.case.c_ref, cfg.solver.eta_limiter, cfg.solver.diff_delta, cfg.solver.diff_alpha,
        cfg.solver.name, cfg.kernel.name, cfg.kernel.h_factor, cfg.solver.is_bc_trick,
        cfg.solver.density_evolution, cfg.solver.artificial_alpha, cfg.solver.free_slip,
        cfg.solver.density_renormalize, cfg.solver.heat_conduction)
    forward = solver.forward_wrapper()
    neighbor_fn = partition.neighbor_list(
        displacement_fn, box_size, r_cutoff=solver._kernel_fn.cutoff, backend=cfg.nl.backend,
        capacity_multiplier=1.25, mask_self=False, format=Sparse,
        num_particles_max=state["r"].shape[0], num_partitions=cfg.nl.num_partitions,
        pbc=np.array(cfg.case.pbc))
    num_particles = (state["tag"] != Tag.PAD_VALUE).sum()
    neighbors = neighbor_fn.allocate(state["r"], num_particles=num_particles)
    advance = si_euler(cfg.solver.tvf, forward, shift_fn, bc_fn, nw_fn)
    advance = jax.jit(advance)  # simulate.py:93 (the shim's jit only converts numpy leaves)

    out = {}
    n = int(state["r"].shape[0])
    tag = "f64" if x64 else "f32"
    for k, v in state.items():
        out[f"state0_{tag}_{k}"] = np.array(v)
    keys = canonical_pairs(np.array(neighbors.idx), n)
    assert len(np.unique(keys)) == len(keys)
    out[f"pairs_{tag}_counts"], out[f"pairs_{tag}_recv"] = pack_pairs(keys, n)
    out[f"idx_capacity_{tag}"] = np.int64(neighbors.idx.shape[1])

    def snap(prefix, st):
        for k in OUT_KEYS:
            out[f"{prefix}_{tag}_{k}"] = np.array(st[k])

    # simulate.py:110-111: the dt = 0 call that initialises the accelerations
    state0 = {k: np.array(v) for k, v in state.items()}
    _state, _nbrs = advance(0.0, state0, neighbors)
    assert not bool(_nbrs.did_buffer_overflow)
    if not long:
        snap("forward", _state)
    else:
        # simulate.py:110-111 discards this call: the loop starts from the initial state
        del out[f"pairs_{tag}_counts"], out[f"pairs_{tag}_recv"]
    # simulate.py:114-131 without IO
    for step in range(nsteps):
        state_, neighbors_ = advance(cfg.solver.dt, state, neighbors)
        if bool(neighbors_.did_buffer_overflow):
            neighbors = neighbor_fn.allocate(state["r"], num_particles=num_particles)
            state, neighbors = advance(cfg.solver.dt, state, neighbors)
        else:
            state, neighbors = state_, neighbors_
    snap("advance", state)
    if not long:
        keys = canonical_pairs(np.array(neighbors.idx), n)
        out[f"pairs_end_{tag}_counts"], out[f"pairs_end_{tag}_recv"] = pack_pairs(keys, n)

    meta = dict(
        dt=float(cfg.solver.dt), dx=float(cfg.case.dx), dim=int(cfg.case.dim),
        box_size=[float(b) for b in np.asarray(box_size)], c_ref=float(cfg.case.c_ref),
        u_ref=float(cfg.case.u_ref), viscosity=float(cfg.case.viscosity),
        solver=str(cfg.solver.name), tvf=float(cfg.solver.tvf), kernel=str(cfg.kernel.name),
        h_factor=float(cfg.kernel.h_factor), cutoff=float(solver._kernel_fn.cutoff),
        p_ref=float(getattr(eos_fn, "p_ref", 0.0)), p_bg=float(eos_fn.p_bg),
        eos=type(eos_fn).__name__, nsteps=nsteps, n=n, pbc=[bool(b) for b in cfg.case.pbc],
        cli=cli)
    out[f"meta_{tag}"] = np.array(json.dumps(meta))
    np.savez(os.path.join(HERE, f"_tmp_{'long_' if long else ''}{name}_{tag}.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--child", nargs=2, metavar=("CASE", "X64"))
    ap.add_argument("--long", action="store_true", help="the 200-step trajectories (ref200_*.npz)")
    ap.add_argument("--kernels", action="store_true", help="kernel tables (kernel_tables.npz)")
    a = ap.parse_args()
    if a.child:
        if a.child[0] == "__kernels__":
            run_kernels(int(a.child[1]))
        else:
            run_case(a.child[0], int(a.child[1]), long=a.long)
        return
    if a.kernels:
        merged = {}
        for x64 in (0, 1):
            subprocess.run([sys.executable, "-W", "ignore", os.path.abspath(__file__), "--child",
                            "__kernels__", str(x64)], check=True, stdout=subprocess.DEVNULL)
            tmp = os.path.join(HERE, f"_tmp_kernels_{'f64' if x64 else 'f32'}.npz")
            with np.load(tmp) as z:
                merged.update({k: z[k] for k in z.files})
            os.remove(tmp)
        np.savez_compressed(os.path.join(HERE, "kernel_tables.npz"), **merged)
        print("kernel_tables.npz", len(merged), "arrays")
        return
    for name, (cli, kw) in (LONG_CASES if a.long else CASES).items():
        if a.only and name not in a.only:
            continue
        merged = {"make_case_kwargs": np.array(json.dumps(kw))}
        for x64 in (0, 1):
            subprocess.run([sys.executable, "-W", "ignore", os.path.abspath(__file__), "--child",
                            name, str(x64)] + (["--long"] if a.long else []), check=True,
                           stdout=subprocess.DEVNULL)
            tmp = os.path.join(HERE, f"_tmp_{'long_' if a.long else ''}{name}_{'f64' if x64 else 'f32'}.npz")
            with np.load(tmp) as z:
                merged.update({k: z[k] for k in z.files})
            os.remove(tmp)
        path = os.path.join(HERE, f"{'ref200' if a.long else 'ref'}_{name}.npz")
        np.savez_compressed(path, **merged)
        print(name, int(json.loads(str(merged["meta_f32"]))["n"]), "particles ->",
              os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()
