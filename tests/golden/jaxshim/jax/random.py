"""jax.random stand-in: NOT bit-compatible with threefry; seeded torch generators.  Test infrastructure only.

The golden generator never compares noise produced here with anything: cases are run
with r0_noise_factor = 0 or with positions passed in explicitly.
"""
import torch as _torch

from ._core import fdt, to_dtype


def PRNGKey(seed):
    return _torch.tensor([0, int(seed)], dtype=_torch.int64)


key = PRNGKey


def split(key, num=2):
    return [_torch.tensor([int(key[1]) + 1, int(key[1]) * 7919 + i], dtype=_torch.int64) for i in range(num)]


def _gen(key):
    g = _torch.Generator()
    g.manual_seed(int(key[0]) * 1000003 + int(key[1]))
    return g


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    u = _torch.rand(tuple(shape), generator=_gen(key), dtype=to_dtype(dtype) or fdt())
    return u * (maxval - minval) + minval


def normal(key, shape=(), dtype=None):
    return _torch.randn(tuple(shape), generator=_gen(key), dtype=to_dtype(dtype) or fdt())
