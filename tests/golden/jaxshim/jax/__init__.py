"""Stand-in for the `jax` top-level package (see _core.py).  Test infrastructure only."""

import functools as _ft

import numpy as _np
import torch as _torch

from . import _core

_core.install()

from . import numpy  # noqa: E402
from . import core, lax, ops, random, tree_util  # noqa: E402
from . import tree_util as tree  # noqa: E402

Array = _torch.Tensor


class _Config:
    def update(self, key, value):
        if key == "jax_enable_x64":
            _core.set_x64(value)

    @property
    def jax_enable_x64(self):
        return _core._X64[0]


config = _Config()


def _conv(x):
    if isinstance(x, (_np.ndarray, _np.generic)):
        return _core.T(x)
    return x


def _conv_arrays(x):
    # what tracing does to array arguments of a jitted function: numpy -> device array
    if isinstance(x, _np.ndarray) and x.dtype.kind in "fiub":
        return _core.T(x)
    return x


def jit(f=None, static_argnums=None, static_argnames=None, **kw):
    if f is None:
        return lambda g: jit(g, static_argnums=static_argnums, static_argnames=static_argnames)
    static = (static_argnums,) if isinstance(static_argnums, int) else tuple(static_argnums or ())
    names = (static_argnames,) if isinstance(static_argnames, str) else tuple(static_argnames or ())

    @_ft.wraps(f)
    def g(*args, **kwargs):
        args = tuple(a if i in static else tree_util.tree_map(_conv_arrays, a)
                     for i, a in enumerate(args))
        kwargs = {k: (v if k in names else tree_util.tree_map(_conv_arrays, v))
                  for k, v in kwargs.items()}
        return f(*args, **kwargs)
    return g


def vmap(f, in_axes=0, out_axes=0):
    def g(*args, **kwargs):
        args = tree_util.tree_map(_conv, args)
        ia = in_axes
        if isinstance(ia, list):
            ia = tuple(ia)
        return _torch.func.vmap(f, in_dims=ia, out_dims=out_axes)(*args, **kwargs)
    return g


def grad(f, argnums=0):
    def g(*args):
        args = tuple(_core.T(a) if not isinstance(a, _torch.Tensor) else a for a in args)
        return _torch.func.grad(f, argnums=argnums)(*args)
    return g


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = _core.to_dtype(dtype)


def eval_shape(f, *args, **kwargs):
    def mk(x):
        if isinstance(x, (ShapeDtypeStruct, core.ShapedArray)):
            return _torch.zeros(x.shape, dtype=_core.to_dtype(x.dtype))
        return x
    out = f(*tree_util.tree_map(mk, args), **kwargs)
    return tree_util.tree_map(
        lambda t: ShapeDtypeStruct(t.shape, t.dtype) if isinstance(t, _torch.Tensor) else t, out)


class custom_jvp:
    def __init__(self, f, nondiff_argnums=()):
        self.f = f
        _ft.update_wrapper(self, f)

    def defjvp(self, jvp):
        return jvp

    def __call__(self, *a, **k):
        return self.f(*a, **k)


def pure_callback(callback, result_shape, *args, **kw):
    out = callback(*tree_util.tree_map(lambda t: t.numpy() if isinstance(t, _torch.Tensor) else t, args))
    return tree_util.tree_map(_conv, out)


def device_get(x):
    return tree_util.tree_map(lambda t: t.numpy() if isinstance(t, _torch.Tensor) else t, x)


def devices(*a):
    return ["cpu"]
