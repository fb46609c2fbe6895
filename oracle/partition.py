"""Neighbour search (oracle; test infrastructure only).

Two builders of the same object -- the reference's Sparse neighbour list
``idx[2, E]`` with row 0 = receiver/candidate, row 1 = sender (ascending),
padding value N:

* ``neighbor_list_reference``: faithful restatement of the jax-md algorithm the
  reference runs (dense cell buffers, 3^d shifted candidate blocks, distance
  prune, cumsum compaction, capacity rules and error bits):
  jax_sph/jax_md/partition.py:114-156 (_cell_dimensions), :177-187 (hash
  constants), :195-199 (capacity), :202-224 (shift_array), :294-413
  (cell_list_fn), :832-856 (candidates), :885-909 (prune), :914-983
  (neighbor_fn / allocate).  Memory is N*3^d*K, so small N only.
* ``neighbor_pairs``: scalable builder (scipy cKDTree proposes candidates with a
  widened radius, membership is then decided by the reference metric in the
  state dtype, i.e. ``sum(periodic_displacement(..)**2) < cutoff**2``).  Returns
  exactly the same *set* (tests/test_oracle_pins.py cross-checks the two).
"""

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import space

# PartitionErrorCode, jax_md/partition.py:434-457
NEIGHBOR_LIST_OVERFLOW = 1 << 0
CELL_LIST_OVERFLOW = 1 << 1


def cutoff_sq(cutoff, dtype):
    """partition.py:820-822: weak-typed python scalar squared in the position dtype."""
    c = np.dtype(dtype).type(cutoff)
    return np.dtype(dtype).type(c * c)


def metric_sq(box, Ra, Rb):
    """_displacement_or_metric_to_metric_sq of space.periodic(side=box)."""
    dR = space.periodic_displacement(box, (Ra - Rb).astype(Ra.dtype))
    return space.square_distance(dR)


def cell_dimensions(dim, box, minimum_cell_size, dtype):
    """jax_md/partition.py:114-156; box is cast to f32 first (:818)."""
    box32 = np.asarray(box, dtype=np.float32).reshape(-1)
    if box32.size == 1:
        box32 = np.repeat(box32, dim)
    mcs = np.float32(minimum_cell_size) if np.dtype(dtype) == np.float32 else np.float64(
        minimum_cell_size
    )
    cells_per_side = np.floor(box32 / mcs)
    cell_size = box32 / cells_per_side
    cells_per_side = cells_per_side.astype(np.int32)
    if (cells_per_side < 3).any():
        raise ValueError("Box must be at least 3x the size of the grid spacing")
    return box32, cell_size, cells_per_side, int(np.prod(cells_per_side))


def particle_hashes(position, cell_size, cells_per_side):
    """jax_md/partition.py:367-368 with the multipliers of :177-187 (x fastest)."""
    dim = position.shape[1]
    mult = np.concatenate([[1], np.cumprod(cells_per_side[:-1])]).astype(np.int32)
    indices = (position / cell_size.astype(position.dtype)).astype(np.int32)  # trunc
    return (indices * mult[:dim]).sum(axis=1).astype(np.int32)


def estimate_cell_capacity(position, box, cell_size_min, multiplier):
    """jax_md/partition.py:195-199 (+ count_cell_filling :159-174)."""
    dim = position.shape[1]
    _, cell_size, cps, count = cell_dimensions(dim, box, cell_size_min, position.dtype)
    h = particle_hashes(position, cell_size, cps)
    filling = np.bincount(h, minlength=count)
    return int(filling.max() * multiplier)


def _shift_array(arr, dindex):
    """jax_md/partition.py:202-224.  arr is indexed [z][y][x] (or [y][x]) and
    dindex = (dx, dy[, dz]) acts on axes 0,1,2 exactly as the reference does."""
    if len(dindex) == 2:
        dx, dy = dindex
        dz = 0
    else:
        dx, dy, dz = dindex
    if dx < 0:
        arr = np.concatenate((arr[1:], arr[:1]))
    elif dx > 0:
        arr = np.concatenate((arr[-1:], arr[:-1]))
    if dy < 0:
        arr = np.concatenate((arr[:, 1:], arr[:, :1]), axis=1)
    elif dy > 0:
        arr = np.concatenate((arr[:, -1:], arr[:, :-1]), axis=1)
    if dz < 0:
        arr = np.concatenate((arr[:, :, 1:], arr[:, :, :1]), axis=2)
    elif dz > 0:
        arr = np.concatenate((arr[:, :, -1:], arr[:, :, :-1]), axis=2)
    return arr


@dataclass
class NeighborList:
    """Mirror of jax_md/partition.py:620-671 (fields the callers read)."""

    idx: np.ndarray  # (2, max_occupancy) int32
    reference_position: np.ndarray
    error_code: int
    cell_list_capacity: Optional[int]
    max_occupancy: int
    cell_size: Optional[float]

    @property
    def did_buffer_overflow(self):
        return bool(self.error_code & (NEIGHBOR_LIST_OVERFLOW | CELL_LIST_OVERFLOW))


def neighbor_list_reference(
    position,
    box,
    r_cutoff,
    mask_self=False,
    capacity_multiplier=1.25,
    neighbors: Optional[NeighborList] = None,
    extra_capacity=0,
):
    """allocate (neighbors=None) or update of the jaxmd_vmap backend, Sparse format.

    Follows neighbor_fn, jax_md/partition.py:914-983, with dr_threshold = 0
    (simulate.py never passes one), i.e. the list is rebuilt on every call.
    """
    N, dim = position.shape
    dtype = position.dtype
    c2 = cutoff_sq(r_cutoff, dtype)
    box32 = np.asarray(box, dtype=np.float32).reshape(-1)
    err = 0 if neighbors is None else neighbors.error_code

    cmp_cut = np.float32(r_cutoff) if dtype == np.float32 else np.float64(r_cutoff)
    use_cells = bool(np.all(cmp_cut < box32 / np.float32(3.0)))  # :929-931
    cell_capacity = None
    if use_cells:
        _, cell_size, cps, count = cell_dimensions(dim, box, r_cutoff, dtype)
        if neighbors is None:
            cell_capacity = estimate_cell_capacity(position, box, r_cutoff, capacity_multiplier)
            cell_capacity += extra_capacity
        else:
            cell_capacity = neighbors.cell_list_capacity
        K = cell_capacity
        hashes = particle_hashes(position, cell_size, cps)
        sort_map = np.argsort(hashes, kind="stable")  # :377
        sorted_hash = hashes[sort_map]
        sorted_cell_id = np.mod(np.arange(N, dtype=np.int32), K) + sorted_hash * K  # :385-386
        cell_id = np.full((count * K,), N, dtype=np.int32)
        # out-of-range scatter indices are dropped by XLA (mode defaults to drop for .at[].set)
        ok = (sorted_cell_id >= 0) & (sorted_cell_id < count * K)
        cell_id[sorted_cell_id[ok]] = sort_map[ok].astype(np.int32)
        occupancy = np.bincount(np.clip(hashes, 0, count - 1), minlength=count)
        if occupancy.max() > K:
            err |= CELL_LIST_OVERFLOW
        shape = tuple(int(x) for x in cps[::-1]) + (K,)  # unflatten_cell_buffer :227-240
        idb = cell_id.reshape(shape)
        blocks = [idb]
        for dindex in np.ndindex(*([3] * dim)):  # _neighboring_cells :190-192
            d = np.array(dindex, dtype=np.int32) - 1
            if np.all(d == 0):
                continue
            blocks.append(_shift_array(idb, d))
        cell_idx = np.concatenate(blocks, axis=-1)  # (..., 3^d K)
        flat_ids = idb.reshape(-1)
        cand = np.zeros((N + 1, cell_idx.shape[-1]), dtype=np.int32)
        rows = np.repeat(cell_idx.reshape(-1, cell_idx.shape[-1]), K, axis=0)
        cand[flat_ids] = rows  # copy_values_from_cell :848-851
        cand = cand[:-1]
    else:
        cand = np.broadcast_to(np.arange(N, dtype=np.int32)[None, :], (N, N)).copy()  # :825-830

    if mask_self:  # :858-863
        cand = np.where(cand == np.arange(N, dtype=np.int32)[:, None], N, cand)

    # prune_neighbor_list_sparse :885-909
    sender = np.broadcast_to(np.arange(N, dtype=np.int32)[:, None], cand.shape).reshape(-1)
    receiver = cand.reshape(-1)
    rc = np.minimum(receiver, N - 1)  # gather clamps
    dR = metric_sq(box, position[sender], position[rc])
    mask = (dR < c2) & (receiver < N)
    occupancy = int(mask.sum())
    out_r = receiver[mask]
    out_s = sender[mask]

    if neighbors is None:  # :956-971
        max_occupancy = int(occupancy * capacity_multiplier + N * extra_capacity)
        max_occupancy = min(max_occupancy, receiver.shape[0])
        cap_limit = N * (N - 1) if mask_self else N * N
        max_occupancy = min(max_occupancy, cap_limit)
    else:
        max_occupancy = neighbors.max_occupancy
    idx = np.full((2, max_occupancy), N, dtype=np.int32)
    k = min(occupancy, max_occupancy)
    idx[0, :k] = out_r[:k]
    idx[1, :k] = out_s[:k]
    if occupancy > max_occupancy:
        err |= NEIGHBOR_LIST_OVERFLOW
    return NeighborList(
        idx,
        position.copy(),
        err,
        cell_capacity,
        max_occupancy,
        float(r_cutoff) if use_cells else None,
    )


def brute_force_pairs(position, box, r_cutoff, mask_self=False):
    """All-pairs definition of the neighbour set (small N)."""
    N = position.shape[0]
    c2 = cutoff_sq(r_cutoff, position.dtype)
    ii, jj = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    # edge (receiver=ii, sender=jj) exists iff metric(r[sender], r[receiver]) < c2
    # (prune_neighbor_list_sparse evaluates d(position[sender], position[receiver]))
    d = metric_sq(box, position[jj.ravel()], position[ii.ravel()])
    m = d < c2
    if mask_self:
        m &= ii.ravel() != jj.ravel()
    return canonical_pairs(np.stack([ii.ravel()[m], jj.ravel()[m]]).astype(np.int32))


def neighbor_pairs(position, box, r_cutoff, mask_self=False, widen=1e-3):
    """Scalable builder of the same set; returns idx (2, E) sorted by (sender, receiver)."""
    from scipy.spatial import cKDTree

    N, dim = position.shape
    dtype = position.dtype
    box64 = np.asarray(box, dtype=np.float64).reshape(-1)
    c2 = cutoff_sq(r_cutoff, dtype)
    p64 = np.mod(position.astype(np.float64), box64)
    p64 = np.where(p64 >= box64, 0.0, p64)
    tree = cKDTree(p64, boxsize=box64)
    pairs = tree.query_pairs(float(r_cutoff) * (1.0 + widen), output_type="ndarray")
    a = pairs[:, 0].astype(np.int32)
    b = pairs[:, 1].astype(np.int32)
    # The reference metric is evaluated per directed edge as
    # d(position[sender], position[receiver]); (a,b) and (b,a) can differ in the
    # last bit at exact ties, so each direction is decided separately.
    k_ab = metric_sq(box, position[b], position[a]) < c2  # receiver a, sender b
    k_ba = metric_sq(box, position[a], position[b]) < c2  # receiver b, sender a
    recv = np.concatenate([a[k_ab], b[k_ba]])
    send = np.concatenate([b[k_ab], a[k_ba]])
    if not mask_self:
        ar = np.arange(N, dtype=np.int32)
        recv = np.concatenate([recv, ar])
        send = np.concatenate([send, ar])
    return canonical_pairs(np.stack([recv, send]).astype(np.int32))


def canonical_pairs(idx, N=None):
    """Strip padding and sort by (sender, receiver) -- the comparison form of
    tests/test_neighbors.py:77-82."""
    idx = np.asarray(idx)
    if N is not None:
        idx = idx[:, idx[0] < N]
    order = np.lexsort((idx[0], idx[1]))
    return np.ascontiguousarray(idx[:, order])


def tie_band(position, box, r_cutoff, idx, ulps=4):
    """Number of listed edges whose d^2 lies within `ulps` ulp of cutoff^2
    (report the tie band with every set comparison)."""
    c2 = cutoff_sq(r_cutoff, position.dtype)
    d = metric_sq(box, position[idx[0]], position[idx[1]])
    return int((np.abs(d - c2) <= ulps * np.spacing(c2)).sum())
