"""jax.ops stand-in.  Test infrastructure only."""
import torch as _torch

from ._core import T


def segment_sum(data, segment_ids, num_segments=None, indices_are_sorted=False, **kw):
    data, ids = T(data), T(segment_ids).long()
    n = int(num_segments) if num_segments is not None else int(ids.max()) + 1
    ok = (ids >= 0) & (ids < n)  # jax drops out-of-range segment ids
    out = _torch.zeros((n,) + tuple(data.shape[1:]), dtype=data.dtype)
    if bool(ok.all()):
        out.index_add_(0, ids, data)
    else:
        out.index_add_(0, ids[ok], data[ok])
    return out
