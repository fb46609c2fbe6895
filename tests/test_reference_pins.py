"""CPU: the NumPy oracle pinned against output of the UNMODIFIED reference.

tests/golden/ref_<case>.npz were produced by executing /root/reference's own
`main.load_embedded_configs`, `SimulationSetup.initialize`,
`partition.neighbor_list`, `WCSPH.forward_wrapper` and `si_euler` against a
torch-backed stand-in for jax (tests/golden/make_reference_golden.py, which
cannot run on the GPU box: /root/reference does not travel).  Here the oracle
(oracle/) is run on the reference's own initial state and must reproduce

* the case setup (dt, box, every field of state0; positions up to the noise the
  stand-in RNG draws differently),
* the neighbour set of state0 and of the final state, bit for bit,
* WCSPH.forward and the state after 20 advance() calls: float64 to 1e-9
  relative (same formulas, summation order aside), float32 within the parity
  tolerance of tests/_util.py.
"""

import glob
import json
import os

import numpy as np
import pytest

from _util import GOLDEN, assert_close, max_err

REF_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
NAMES = [os.path.basename(p)[4:-4] for p in REF_FILES]
OUT_KEYS = ("r", "u", "v", "rho", "p", "dudt", "dvdt", "drhodt", "T", "dTdt")


def unpack_pairs(counts, recv):
    n = len(counts)
    send = np.repeat(np.arange(n, dtype=np.int64), counts)
    return send * n + recv.astype(np.int64)


def load_ref(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    kw = json.loads(str(z["make_case_kwargs"]))
    return z, kw


def oracle_setup(z, kw, tag):
    """oracle.cases.make_case for the case, then the reference's own state0 as input."""
    from oracle import cases

    dtype = np.float32 if tag == "f32" else np.float64
    setup = cases.make_case(dtype=dtype, **kw)
    own_state = setup.state
    meta = json.loads(str(z[f"meta_{tag}"]))
    setup.state = {k: np.array(z[f"state0_{tag}_{k}"]) for k in own_state}
    return setup, own_state, meta


def oracle_run(setup, nsteps, dt):
    from oracle import integrator
    from oracle.solver import WCSPH

    solver = WCSPH(setup.displacement_fn, setup.eos, setup.g_ext_fn, setup.dx, setup.dim, setup.dt,
                   setup.c_ref, setup.eta_limiter, setup.diff_delta, setup.diff_alpha, setup.solver, setup.kernel,
                   setup.h_factor, setup.is_bc_trick, setup.density_evolution,
                   setup.artificial_alpha, setup.free_slip, setup.density_renormalize,
                   setup.heat_conduction, dtype=setup.dtype)
    nfn = integrator.make_neighbors_fn(setup.box_size, solver._kernel_fn.cutoff)
    adv = integrator.si_euler(setup.tvf, solver.forward, setup.shift_fn, setup.bc_fn, setup.nw_fn)
    st = {k: np.array(v, copy=True) for k, v in setup.state.items()}
    idx = None
    for _ in range(nsteps):
        st, idx = adv(dt, st, nfn)
    return st, idx


def pair_keys(idx, n):
    idx = np.asarray(idx)
    ok = (idx[0] < n) & (idx[1] < n)
    return np.sort(idx[1][ok].astype(np.int64) * n + idx[0][ok].astype(np.int64))


def test_reference_goldens_present():
    assert len(NAMES) >= 10, "tests/golden/ref_*.npz missing (make_reference_golden.py)"


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", NAMES)
def test_case_setup_matches_reference(name, tag):
    """oracle.cases.make_case == SimulationSetup.initialize (case_setup.py:44-231)."""
    z, kw = load_ref(name)
    setup, own, meta = oracle_setup(z, kw, tag)
    eps = np.finfo(setup.dtype).eps
    assert abs(setup.dt - meta["dt"]) <= 1e-12 * meta["dt"]
    assert np.allclose(setup.box_size, meta["box_size"], rtol=1e-12)
    assert setup.dim == meta["dim"] and len(own["r"]) == meta["n"]
    assert abs(setup.c_ref - meta["c_ref"]) <= 1e-12 * meta["c_ref"]
    factor = [float(c.split("=")[1]) for c in meta["cli"] if c.startswith("case.r0_noise_factor")]
    factor = factor[0] if factor else (0.05 if kw["case"] == "ht" else 0.0)  # cases/ht.yaml
    if factor != 0.0:
        # jax.random is not reproducible without jax, so the noise drawn by the stand-in RNG
        # differs from the oracle's: compare against the oracle's NOISE-FREE lattice instead
        # and check the noise statistics (case_setup.py:138-144: std = factor * dx, fluid only)
        from oracle import cases

        own = cases.make_case(dtype=setup.dtype, **dict(kw, r0_noise_factor=0.0)).state
    for k, mine in own.items():
        ref = z[f"state0_{tag}_{k}"]
        assert mine.shape == ref.shape, k
        if k == "tag":  # int32, int64 under x64 (case_setup.py:243)
            assert np.array_equal(mine, ref)
            continue
        assert mine.dtype == ref.dtype, k
        scale = max(1.0, float(np.abs(ref).max()))
        if factor != 0.0 and k in ("r", "u", "v"):
            if k == "r":
                box = np.asarray(meta["box_size"])[None, :]
                d = (ref - mine + 0.5 * box) % box - 0.5 * box
                fluid = own["tag"] == 0
                assert max_err(d[~fluid], 0 * d[~fluid]) <= 4 * eps * scale  # walls get no noise
                std = d[fluid].std()
                assert abs(std - factor * setup.dx) <= 0.15 * factor * setup.dx, (std, factor)
                assert abs(d[fluid].mean()) <= 0.1 * factor * setup.dx
            continue  # u, v are functions of the noisy r
        # wall normals: unit vectors from a float64 KD-tree query (utils.py:169-194), cast down
        tol = 2e-5 if (k == "nw" and tag == "f32") else 4 * eps * scale
        assert max_err(mine, ref) <= tol, (k, max_err(mine, ref))


@pytest.mark.parametrize("name", NAMES)
def test_neighbor_sets_match_reference(name):
    """oracle.partition == jax_md cell list + prune (jax_md/partition.py:832-909), bit exact."""
    from oracle import partition

    z, kw = load_ref(name)
    for tag in ("f32", "f64"):
        setup, _, meta = oracle_setup(z, kw, tag)
        n = meta["n"]
        ref = unpack_pairs(z[f"pairs_{tag}_counts"], z[f"pairs_{tag}_recv"])
        idx = partition.neighbor_pairs(setup.state["r"], np.asarray(meta["box_size"]), meta["cutoff"])
        assert np.array_equal(pair_keys(idx, n), ref), f"{name} {tag}: neighbour set differs"
        if n <= 3000 and tag == "f32":  # the literal restatement of the cell-list algorithm too
            nl = partition.neighbor_list_reference(setup.state["r"], np.asarray(meta["box_size"]),
                                                   meta["cutoff"])
            assert np.array_equal(pair_keys(nl.idx, n), ref)


@pytest.mark.parametrize("name", NAMES)
def test_forward_and_advance_f64_match_reference(name):
    """Same formulas => float64 agreement far below the float32 parity tolerance."""
    z, kw = load_ref(name)
    setup, _, meta = oracle_setup(z, kw, "f64")
    fwd, _ = oracle_run(setup, 1, 0.0)
    adv, idx = oracle_run(setup, meta["nsteps"], meta["dt"])
    for prefix, st, rt in (("forward", fwd, 1e-9), ("advance", adv, 1e-8)):
        for k in OUT_KEYS:
            ref = z[f"{prefix}_f64_{k}"]
            # the same float32-style floor, scaled to float64 (eps ratio 2^-29)
            tol = rt * max(float(np.abs(ref).max()), 1e-300) + 2.0 ** -29 * 100 * (
                assert_tol(k, ref, setup))
            assert max_err(st[k], ref) <= tol, (name, prefix, k, max_err(st[k], ref), tol)
    ref_end = unpack_pairs(z["pairs_end_f64_counts"], z["pairs_end_f64_recv"])
    assert np.array_equal(pair_keys(idx, meta["n"]), ref_end)


def assert_tol(k, ref, setup):
    from _util import tolerance

    return tolerance(k, ref, setup)


@pytest.mark.parametrize("name", NAMES)
def test_forward_and_advance_f32_match_reference(name):
    """float32 oracle vs float32 reference within the parity tolerance (tests/_util.py)."""
    z, kw = load_ref(name)
    setup, _, meta = oracle_setup(z, kw, "f32")
    fwd, _ = oracle_run(setup, 1, 0.0)
    for k in OUT_KEYS:
        assert_close(k, fwd[k], z[f"forward_f32_{k}"], setup, what=f"{name} forward")
    adv, _ = oracle_run(setup, meta["nsteps"], meta["dt"])
    for k in OUT_KEYS:
        assert_close(k, adv[k], z[f"advance_f32_{k}"], setup, factor=5.0, what=f"{name} advance")


# ---- 200-step trajectories (north_star: "tolerance-matched 200-step trajectories") ----------
LONG_NAMES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "ref200_*.npz")))


def load_long(name, tag):
    from oracle import cases

    z = np.load(os.path.join(GOLDEN, f"ref200_{name}.npz"))
    kw = json.loads(str(z["make_case_kwargs"]))
    meta = json.loads(str(z[f"meta_{tag}"]))
    setup = cases.make_case(dtype=np.float32 if tag == "f32" else np.float64, **kw)
    setup.state = {k: np.array(z[f"state0_{tag}_{k}"]) for k in setup.state}
    assert abs(setup.dt - meta["dt"]) <= 1e-12 * meta["dt"]
    return z, setup, meta


def test_long_goldens_present():
    assert len(LONG_NAMES) >= 6, "tests/golden/ref200_*.npz missing (make_reference_golden.py --long)"


@pytest.mark.parametrize("name", LONG_NAMES)
def test_200_steps_match_reference(name):
    """Oracle vs the reference's own 200-step run: float64 to 1e-6 relative (rounding-order
    differences grow along the trajectory), float32 within the trajectory tolerance."""
    z, setup, meta = load_long(name, "f64")
    adv, _ = oracle_run(setup, meta["nsteps"], meta["dt"])
    for k in OUT_KEYS:
        ref = z[f"advance_f64_{k}"]
        tol = 1e-6 * max(float(np.abs(ref).max()), 1e-300) + 1e-6 * assert_tol(k, ref, setup)
        assert max_err(adv[k], ref) <= tol, (name, k, max_err(adv[k], ref), tol)
    z, setup, meta = load_long(name, "f32")
    adv, _ = oracle_run(setup, meta["nsteps"], meta["dt"])
    for k in ("r", "u", "v", "rho", "T"):
        assert_close(k, adv[k], z[f"advance_f32_{k}"], setup, factor=15.0, what=f"{name} 200 steps")


# ---- kernel tables -------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["CSK", "QSK", "WC2K", "WC4K", "WC6K", "GK", "SGK"])
@pytest.mark.parametrize("dim", [2, 3])
def test_kernels_match_reference_tables(name, dim):
    """w(r) and jax.grad(w)(r) of every kernel class of jax_sph/kernel.py on a grid including
    r = 0, the knots and the cutoff (tests/golden/kernel_tables.npz): the oracle's closed forms
    agree to rounding in float64 and to 2e-6 of the peak in float32."""
    from oracle import kernel as K

    z = np.load(os.path.join(GOLDEN, "kernel_tables.npz"))
    h = float(z["h"])
    for tag, dt, tol in (("f64", np.float64, 1e-13), ("f32", np.float32, 2e-6)):
        k = K.KERNELS[name](h=h, dim=dim, dtype=dt)
        assert abs(k.cutoff - float(z[f"{name}_{dim}_cutoff"])) < 1e-15
        r = z[f"{name}_{dim}_{tag}_r"]
        for what, mine in (("w", k.w(r)), ("gw", k.grad_w(r))):
            ref = z[f"{name}_{dim}_{tag}_{what}"]
            assert max_err(mine, ref) <= tol * np.abs(ref).max(), (name, dim, tag, what)
