"""Neighbour search, host mirror of jax_sph/partition.py:485-571.

``neighbor_list(...)`` keeps the reference signature and returns
``NeighborListFns(allocate, update)``; the ``"sphb200"`` backend (the only one)
builds the Sparse list with the CUDA cell pipeline + sweep of libsphb200.so.
The reference's backend names are accepted as aliases so that
``cfg.nl.backend`` values from existing YAML files keep working
(simulate.py:75-88).

``NeighborList`` mirrors jax_md/partition.py:620-671: ``idx`` is (2, capacity)
int32 with row 0 = receiver, row 1 = sender (ascending), padding value N;
``did_buffer_overflow`` follows PartitionErrorCode (:434-457).  Within one
sender the receivers are ascending (the reference leaves that order
unspecified; tests compare sorted pairs, tests/test_neighbors.py:77-82).
"""

import ctypes as C
import dataclasses
from typing import Any, Callable, Optional

import numpy as np

from . import _lib
from .engine import Engine, make_config

NEIGHBOR_LIST_OVERFLOW = 1 << 0  # jax_md/partition.py:434-457
CELL_LIST_OVERFLOW = 1 << 1


class NeighborListFormat:
    Dense, Sparse, OrderedSparse = 0, 1, 2


Sparse = NeighborListFormat.Sparse
BACKENDS = ("sphb200", "jaxmd_vmap", "jaxmd_scan", "matscipy")


@dataclasses.dataclass
class NeighborList:
    idx: Any
    reference_position: Any
    error_code: int
    cell_list_capacity: Optional[int]
    max_occupancy: int
    format: int
    cell_size: Optional[float]
    update_fn: Callable

    @property
    def did_buffer_overflow(self) -> bool:
        return bool(self.error_code & (NEIGHBOR_LIST_OVERFLOW | CELL_LIST_OVERFLOW))

    def update(self, position, **kwargs):
        return self.update_fn(position, self, **kwargs)


@dataclasses.dataclass
class NeighborListFns:
    allocate: Callable
    update: Callable

    def __iter__(self):
        return iter((self.allocate, self.update))


def neighbor_list(
    displacement_or_metric, box_size, r_cutoff: float, backend: str = "sphb200",
    dr_threshold: float = 0.0, capacity_multiplier: float = 1.25, disable_cell_list: bool = False,
    mask_self: bool = True, custom_mask_function=None, fractional_coordinates: bool = False,
    format=Sparse, num_particles_max: Optional[int] = None, num_partitions: int = 1, pbc=None,
) -> NeighborListFns:
    """Same arguments as jax_sph/partition.py:492-507.  Unsupported options raise."""
    if backend not in BACKENDS:
        raise ValueError(f"unknown neighbour-list backend {backend!r}")
    if format != Sparse:
        raise NotImplementedError("only the Sparse format is on the hot path (simulate.py:85)")
    if custom_mask_function is not None or fractional_coordinates:
        raise NotImplementedError("custom masks / fractional coordinates are not supported")
    if pbc is not None and not all(np.asarray(pbc).reshape(-1)):
        raise NotImplementedError("non-periodic boxes are not supported (space.periodic only)")
    box = np.asarray(box_size, dtype=np.float64).reshape(-1)
    engines = {}

    def _engine(position):
        n, dim = position.shape
        key = (n, dim)
        if key not in engines:
            b = box if box.size == dim else np.repeat(box, dim)
            # dx only sizes the staging buffers: estimate it from the number density
            dx = float((np.prod(b) / max(n, 1)) ** (1.0 / dim))
            cfg = make_config(dim, b, dx, 0.0, r_cutoff=float(r_cutoff))
            engines[key] = Engine(cfg, n)
        return engines[key]

    def _build(position, capacity, prev_err):
        eng = _engine(position)
        eng.upload({"r": position})
        if capacity is None:  # allocate: count, then size like jax_md/partition.py:956-971
            _, count = eng.neighbor_list(0, mask_self=mask_self)
            return eng, count
        idx, count = eng.neighbor_list(capacity, mask_self=mask_self)
        err = prev_err | eng.error()
        code = (NEIGHBOR_LIST_OVERFLOW if err & _lib.ERR_NEIGHBOR_OVERFLOW else 0) | (
            CELL_LIST_OVERFLOW if err & (_lib.ERR_CELL_OVERFLOW | _lib.ERR_STAGE_OVERFLOW) else 0)
        return idx, code

    def allocate(position, extra_capacity: int = 0, **kwargs) -> NeighborList:
        n = position.shape[0]
        eng, count = _build(position, None, 0)
        eng.error()  # clear
        cap = int(count * capacity_multiplier + n * extra_capacity)
        cap = min(cap, n * (n - 1) if mask_self else n * n)
        idx, code = _build(position, cap, 0)
        return NeighborList(idx, position, code, None, cap, Sparse, float(r_cutoff), update)

    def update(position, neighbors: NeighborList, **kwargs) -> NeighborList:
        idx, code = _build(position, neighbors.max_occupancy, neighbors.error_code)
        return dataclasses.replace(neighbors, idx=idx, reference_position=position, error_code=code)

    return NeighborListFns(allocate, update)
