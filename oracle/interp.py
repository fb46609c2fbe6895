"""SPH interpolation onto probe points (oracle; test infrastructure only).

Restates ``sph_interpolator(...)->interp_vel`` of jax_sph/utils.py:299-443, which
the reference's Poiseuille / Couette tests (tests/test_pf2d.py:82-103,
tests/test_cf2d.py) use to compare a trajectory frame with the analytical
profile.  h5 files are replaced by in-memory state dicts.
"""

import numpy as np

from . import partition, space
from .kernel import QuinticKernel
from .solver import FLUID, WALL_TAGS


def interp_vel(state, box_size, dx, dim, r_target, prop="u", dim_ind=0):
    r = state["r"].astype(np.float64)
    tag = state["tag"]
    N = len(r)
    eps = np.finfo(np.float64).eps
    kernel_fn = QuinticKernel(h=dx, dim=dim, dtype=np.float64)
    idx = partition.neighbor_pairs(r, box_size, 3 * dx)
    i_s, j_s = idx
    dr = space.periodic_displacement(box_size, r[i_s] - r[j_s])
    dist = space.distance(dr)
    w_dist = kernel_fn.w(dist)
    w_j_s_fluid = w_dist * np.where(tag[j_s] == FLUID, 1.0, 0.0)
    w_i_sum = np.bincount(i_s, weights=w_j_s_fluid, minlength=N)
    vel = state[prop].astype(np.float64)
    x_wall_unnorm = np.stack(
        [np.bincount(i_s, weights=w_j_s_fluid * vel[j_s, k], minlength=N) for k in range(vel.shape[1])],
        axis=1,
    )
    x_wall = x_wall_unnorm / (w_i_sum[:, None] + eps)
    mask_bc = np.isin(tag, WALL_TAGS)
    vel = np.where(mask_bc[:, None], 2 * vel - x_wall, vel)

    dist = (((r_target[:, None] - r[None, :]) ** 2).sum(axis=-1)) ** 0.5
    w = kernel_fn.w(dist)
    w_norm = w.sum(axis=-1) * dx**dim
    u_val = (w * vel[:, dim_ind][None, :]).sum(axis=1) * dx**dim
    return u_val / w_norm
