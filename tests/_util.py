"""Shared helpers for the parity tests.

Tolerances (float32 engine vs float32 oracle), DESIGN.md section 6
------------------------------------------------------------------
`north_star` asks for relative 1e-5 in FP32 plus a documented drift bound
against float64.  For fields that are plain sums of same-sign terms (rho, T,
positions, velocities) the bound is applied as is: |a - b| <= 1e-5 * max|b|.
Pressure and the accelerations sit behind the stiff equation of state
p = p_ref (rho / rho_ref - 1): a last-bit difference of rho (the two
implementations add the ~100 kernel values in a different order) becomes
p_ref * eps, and the pressure-gradient sum turns that into
~p_ref * eps / (rho dx).  That floor is a property of float32 WCSPH, present
between the reference's own float32 and float64 runs, so the tolerance is

    |a - b| <= RTOL * max|b| + NOISE * floor(field)

with floor(p) = p_ref, floor(dudt) = p_ref / (rho_ref dx),
floor(dvdt) = |p_fn(0)| / (rho_ref dx), floor(drhodt) = rho_ref c_ref / dx * 5e-2
(the continuity sum divides the velocity noise by dx; calibrated on the reference's own
float32 vs float64 runs, tests/test_reference_pins.py),
and `drift_ok` additionally checks that the engine is not further from the
float64 oracle than the float32 oracle is (times a small factor).
"""

import os

import numpy as np

RTOL = 1e-5
NOISE = 4e-6  # a few float32 ulps of the EoS input

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def floors(setup):
    dx, rho0 = setup.dx, setup.rho_ref
    p_bg_tvf = abs(-setup.p_ref + setup.p_bg) if setup.solver != "RIE" else abs(
        -100 * setup.u_ref**2 * rho0 + setup.p_bg)
    acc = max(setup.p_ref, p_bg_tvf) / (rho0 * dx)
    return {
        "p": setup.p_ref if setup.solver != "RIE" else 100 * setup.u_ref**2 * rho0,
        "dudt": setup.p_ref / (rho0 * dx),
        "dvdt": p_bg_tvf / (rho0 * dx),
        "drhodt": rho0 * setup.c_ref / dx * 5e-2,
        "dTdt": 1e-2 / dx,
        # one step of the acceleration floor: u += dt dudt, r += dt v
        "u": setup.dt * acc,
        "v": setup.dt * acc,
        "r": float(np.max(setup.box_size)) * 1e-2 + setup.dt**2 * acc,
    }


def tolerance(key, ref, setup):
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    return RTOL * scale + NOISE * floors(setup).get(key, 0.0) + 1e-30


def max_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0


def assert_close(key, got, ref, setup, factor=1.0, what=""):
    tol = factor * tolerance(key, ref, setup)
    err = max_err(got, ref)
    assert err <= tol, f"{what} {key}: max abs err {err:.3e} > tol {tol:.3e} (max|ref| {np.abs(ref).max():.3e})"
    return err, tol


def drift_ok(key, got, ref32, ref64, setup, factor=2.0):
    """Engine vs float64 no worse than `factor` x (float32 oracle vs float64) + one RTOL unit."""
    e_ours = max_err(got, ref64)
    e_ref = max_err(ref32, ref64)
    slack = tolerance(key, ref64, setup)
    assert e_ours <= factor * e_ref + slack, (
        f"{key}: |engine - f64| = {e_ours:.3e} exceeds {factor} x |oracle32 - f64| = {e_ref:.3e} "
        f"+ {slack:.3e}")
    return e_ours, e_ref


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    state0 = {k[len("state0_"):]: z[k] for k in z.files if k.startswith("state0_")}
    return z, state0


def canonical_pairs(idx, n):
    idx = np.asarray(idx)
    idx = idx[:, idx[0] < n]
    order = np.lexsort((idx[0], idx[1]))
    return np.ascontiguousarray(idx[:, order])
