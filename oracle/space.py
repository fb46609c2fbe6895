"""Periodic space (oracle; test infrastructure only).

Restates jax_sph/jax_md/space.py:170-209 (periodic_displacement, square_distance,
distance, periodic_shift) and :232-286 (periodic).
"""

import numpy as np


def _mod(a, b):
    # jnp.mod == np.mod for floats: fmod + sign fix-up towards the divisor.
    return np.mod(a, b)


def periodic_displacement(side, dR):
    """space.py:170-181: mod(dR + side*0.5, side) - 0.5*side, in dR's dtype."""
    dt = dR.dtype
    side = np.asarray(side, dtype=dt)
    half = (side * dt.type(0.5)).astype(dt)
    return (_mod((dR + half).astype(dt), side) - half).astype(dt)


def square_distance(dR):
    """space.py:184-192. Sum over the last axis, left to right."""
    sq = dR * dR
    acc = sq[..., 0].copy()
    for k in range(1, dR.shape[-1]):
        acc = acc + sq[..., k]
    return acc


def distance(dR):
    """space.py:195-204: safe sqrt (0 where d2 == 0)."""
    d2 = square_distance(dR)
    out = np.zeros_like(d2)
    m = d2 > 0
    out[m] = np.sqrt(d2[m])
    return out


def periodic_shift(side, R, dR):
    """space.py:207-209."""
    dt = R.dtype
    side = np.asarray(side, dtype=dt)
    return _mod((R + dR).astype(dt), side).astype(dt)


def periodic(side):
    """space.py:232-286 -> (displacement_fn, shift_fn), vectorised over rows."""

    def displacement_fn(Ra, Rb):
        return periodic_displacement(side, (Ra - Rb).astype(Ra.dtype))

    def shift_fn(R, dR):
        return periodic_shift(side, R, dR)

    displacement_fn.side = np.asarray(side, dtype=np.float64)
    shift_fn.side = displacement_fn.side
    return displacement_fn, shift_fn
