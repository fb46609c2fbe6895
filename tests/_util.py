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


def periodic_err(a, b, box):
    """max |a - b| of positions up to periodic images: a particle a rounding error away from the
    seam is wrapped to the other side of the box by one implementation and not by the other."""
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64).reshape(-1)[: d.shape[-1]]
    d -= box * np.round(d / box)
    return float(np.abs(d).max()) if d.size else 0.0


AMPLIFIED = 1e-3


def assert_close(key, got, ref, setup, factor=1.0, what="", ref64=None):
    """`ref64` (the float64 oracle, optional): particles on which the reference's OWN float32 and
    float64 results disagree are ill-conditioned -- e.g. the Shepard averages sum(w f) / (sum(w)
    + EPS) of the outermost wall layer, whose only fluid neighbours sit ON the cutoff (sum(w) ~
    EPS): a last-bit change of sum(w) moves the quotient by O(1) between the reference's two
    precisions.  There, and only there, the bound is widened by AMPLIFIED x |ref32 - ref64| of
    that very particle (a summation-order change is a ~1e-7 relative perturbation of the sums,
    the precision change an O(1) one)."""
    tol = factor * tolerance(key, ref, setup)
    if ref64 is not None and key != "r":
        a, b = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
        allow = tol + AMPLIFIED * np.abs(b - np.asarray(ref64, dtype=np.float64))
        bad = np.abs(a - b) > allow
        assert not bad.any(), (f"{what} {key}: {int(bad.sum())} entries off, worst abs err "
                               f"{np.abs(a - b)[bad].max():.3e} > tol {tol:.3e}")
        return float(np.abs(a - b).max()), tol
    err = periodic_err(got, ref, setup.box_size) if key == "r" else max_err(got, ref)
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
