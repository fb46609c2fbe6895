"""jax.numpy.linalg stand-in.  Test infrastructure only."""
import torch as _torch

from .._core import T


def inv(x):
    return _torch.linalg.inv(T(x))


def norm(x, axis=None, keepdims=False, ord=None):
    x = T(x)
    if axis is None:
        return _torch.linalg.norm(x)
    return _torch.linalg.norm(x, dim=axis, keepdim=keepdims)


def det(x):
    return _torch.linalg.det(T(x))
