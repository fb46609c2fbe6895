"""jax.numpy stand-in on torch CPU tensors (see jax/_core.py).  Test infrastructure only."""

import builtins as _b
import math as _math

import numpy as _np
import torch as _torch

from .._core import DType, T, binop_args, canon, fdt, idt, is_scalar, jmod, to_dtype
from . import linalg  # noqa: F401

pi = _math.pi
inf = _math.inf
newaxis = None
ndarray = _torch.Tensor

float32 = DType(_torch.float32, "float32")
float64 = DType(_torch.float64, "float64")
float16 = DType(_torch.float16, "float16")
int32 = DType(_torch.int32, "int32")
int64 = DType(_torch.int64, "int64")
int16 = DType(_torch.int16, "int16")
int8 = DType(_torch.int8, "int8")
uint8 = DType(_torch.uint8, "uint8")
uint32 = DType(_torch.int64, "uint32")
bool_ = DType(_torch.bool, "bool_")
float_ = float
int_ = int


class _Finfo:
    def __init__(self, d):
        fi = _torch.finfo(to_dtype(d))
        self.eps, self.max, self.min, self.tiny = fi.eps, fi.max, fi.min, fi.tiny
        self.dtype = to_dtype(d)


def finfo(d):
    return _Finfo(d)


class _Iinfo:
    def __init__(self, d):
        ii = _torch.iinfo(to_dtype(d))
        self.max, self.min = ii.max, ii.min


def iinfo(d):
    return _Iinfo(d)


def array(x, dtype=None, copy=True):
    t = T(x, dtype)
    return t.clone() if isinstance(x, _torch.Tensor) else t


def asarray(x, dtype=None):
    return T(x, dtype)


def _shape(s):
    if isinstance(s, (int, _np.integer)):
        return (int(s),)
    if isinstance(s, _torch.Tensor):
        return tuple(int(v) for v in s.reshape(-1).tolist())
    return tuple(int(v) for v in s)


def zeros(shape, dtype=None):
    return _torch.zeros(_shape(shape), dtype=to_dtype(dtype) or fdt())


def ones(shape, dtype=None):
    return _torch.ones(_shape(shape), dtype=to_dtype(dtype) or fdt())


def full(shape, fill_value, dtype=None):
    if dtype is None:
        if isinstance(fill_value, _torch.Tensor):
            dt = fill_value.dtype
        elif isinstance(fill_value, (bool, _np.bool_)):
            dt = _torch.bool
        elif isinstance(fill_value, (int, _np.integer)):
            dt = idt()
        else:
            dt = fdt()
    else:
        dt = to_dtype(dtype)
    fv = fill_value.item() if isinstance(fill_value, _torch.Tensor) and fill_value.dim() == 0 else fill_value
    if isinstance(fv, _torch.Tensor):
        return _torch.broadcast_to(fv.to(dt), _shape(shape)).clone()
    return _torch.full(_shape(shape), fv, dtype=dt)


def zeros_like(x, dtype=None):
    return _torch.zeros_like(T(x), dtype=to_dtype(dtype))


def ones_like(x, dtype=None):
    return _torch.ones_like(T(x), dtype=to_dtype(dtype))


def full_like(x, v, dtype=None):
    return _torch.full_like(T(x), v, dtype=to_dtype(dtype))


def arange(*a, dtype=None):
    a = [int(v) if isinstance(v, (_torch.Tensor, _np.integer)) else v for v in a]
    if dtype is None:
        dtype = fdt() if _b.any(isinstance(v, float) for v in a) else idt()
    return _torch.arange(*a, dtype=to_dtype(dtype))


def linspace(a, b, n, dtype=None):
    return _torch.linspace(a, b, n, dtype=to_dtype(dtype) or fdt())


def where(c, a=None, b=None):
    c = T(c)
    if c.dtype != _torch.bool:
        c = c != 0
    if a is None:
        return _torch.where(c)
    a, b = binop_args(a, b)
    return _torch.where(c, a, b)


def _bin(fn):
    def f(a, b):
        a, b = binop_args(a, b)
        return fn(a, b)
    return f


maximum = _bin(_torch.maximum)
minimum = _bin(_torch.minimum)
logical_and = _bin(_torch.logical_and)
logical_or = _bin(_torch.logical_or)
mod = jmod
remainder = jmod
power = _bin(_torch.pow)
add = _bin(_torch.add)
subtract = _bin(_torch.sub)
multiply = _bin(_torch.mul)
divide = _bin(_torch.true_divide)
equal = _bin(_torch.eq)
not_equal = _bin(_torch.ne)
less = _bin(_torch.lt)
greater = _bin(_torch.gt)
arctan2 = _bin(_torch.atan2)


def _un(fn, floating=True):
    def f(x):
        x = T(x)
        if floating and not x.dtype.is_floating_point:
            x = x.to(fdt())
        return fn(x)
    return f


sqrt = _un(_torch.sqrt)
sin = _un(_torch.sin)
cos = _un(_torch.cos)
tan = _un(_torch.tan)
exp = _un(_torch.exp)
log = _un(_torch.log)
tanh = _un(_torch.tanh)
abs = _un(_torch.abs, False)
absolute = abs
floor = _un(_torch.floor)
ceil = _un(_torch.ceil)
sign = _un(_torch.sign, False)
square = _un(_torch.square, False)
logical_not = _un(_torch.logical_not, False)
isnan = _un(_torch.isnan)
isfinite = _un(_torch.isfinite)
round = _un(_torch.round)


def _red(name):
    def f(x, axis=None, keepdims=False, dtype=None):
        x = T(x)
        if name in ("sum", "prod") and x.dtype == _torch.bool:
            x = x.to(idt())
        if axis is None:
            r = getattr(_torch, name)(x)
        else:
            if isinstance(axis, (tuple, list)) and name in ("amin", "amax", "sum", "mean"):
                r = getattr(_torch, name)(x, dim=tuple(axis), keepdim=keepdims)
            else:
                r = getattr(_torch, name)(x, dim=axis, keepdim=keepdims)
        if name in ("sum", "prod") and r.dtype == _torch.int64:
            r = r.to(canon(_torch.int64) if x.dtype != _torch.int64 else _torch.int64)
        return r
    return f


sum = _red("sum")
prod = _red("prod")
mean = _red("mean")
min = _red("amin")
max = _red("amax")
amin = min
amax = max


def any(x, axis=None):
    x = T(x)
    return _torch.any(x != 0) if axis is None else _torch.any(x != 0, dim=axis)


def all(x, axis=None):
    x = T(x)
    return _torch.all(x != 0) if axis is None else _torch.all(x != 0, dim=axis)


def argmin(x, axis=None):
    return _torch.argmin(T(x), dim=axis).to(idt())


def argmax(x, axis=None):
    return _torch.argmax(T(x), dim=axis).to(idt())


def argsort(x, axis=-1, stable=True, kind=None):
    return _torch.argsort(T(x), dim=axis, stable=True).to(idt())


def sort(x, axis=-1):
    return _torch.sort(T(x), dim=axis, stable=True).values


def cumsum(x, axis=None, dtype=None):
    x = T(x)
    if x.dtype == _torch.bool:
        x = x.to(idt())
    if axis is None:
        x, axis = x.reshape(-1), 0
    r = _torch.cumsum(x, dim=axis)
    if not x.dtype.is_floating_point:
        r = r.to(x.dtype if x.dtype != _torch.bool else idt())
    return r


def cumprod(x, axis=None, dtype=None):
    x = T(x)
    if axis is None:
        x, axis = x.reshape(-1), 0
    return _torch.cumprod(x, dim=axis).to(x.dtype)


def dot(a, b):
    a, b = binop_args(a, b)
    if a.dim() == 1 and b.dim() == 1:
        return (a * b).sum()
    if a.dim() == 0 or b.dim() == 0:
        return a * b
    return _torch.matmul(a, b)


matmul = dot


def outer(a, b):
    a, b = T(a), T(b)
    return a.reshape(-1)[:, None] * b.reshape(-1)[None, :]


def einsum(spec, *ops):
    return _torch.einsum(spec, *[T(o) for o in ops])


def tensordot(a, b, axes=2):
    return _torch.tensordot(T(a), T(b), dims=axes)


def concatenate(xs, axis=0):
    xs = [T(x) for x in xs]
    dt = xs[0].dtype
    for x in xs[1:]:
        dt = _torch.promote_types(dt, x.dtype)
    return _torch.cat([x.to(dt) for x in xs], dim=axis)


def stack(xs, axis=0):
    return _torch.stack([T(x) for x in xs], dim=axis)


def vstack(xs):
    xs = [T(x) for x in xs]
    xs = [x[None] if x.dim() == 1 else x for x in xs]
    return _torch.cat(xs, dim=0)


def hstack(xs):
    xs = [T(x) for x in xs]
    return _torch.cat(xs, dim=0 if xs[0].dim() == 1 else 1)


def reshape(x, shape):
    return T(x).reshape(_shape(shape) if not isinstance(shape, int) else (shape,))


def ravel(x):
    return T(x).reshape(-1)


def broadcast_to(x, shape):
    return _torch.broadcast_to(T(x), _shape(shape))


def tile(x, reps):
    x = T(x)
    reps = (reps,) if isinstance(reps, int) else tuple(reps)
    return x.repeat(*reps) if len(reps) >= x.dim() else x.repeat(*((1,) * (x.dim() - len(reps)) + reps))


def transpose(x, axes=None):
    x = T(x)
    return x.permute(*axes) if axes is not None else x.T


def expand_dims(x, axis):
    return T(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return T(x).squeeze() if axis is None else T(x).squeeze(axis)


def clip(x, a_min=None, a_max=None, min=None, max=None):
    lo = a_min if a_min is not None else min
    hi = a_max if a_max is not None else max
    return _torch.clamp(T(x), lo, hi)


def isin(x, test):
    return _torch.isin(T(x), T(test))


def isscalar(x):
    return is_scalar(x) or (isinstance(x, _torch.Tensor) and x.dim() == 0 and False)


def size(x, axis=None):
    x = T(x)
    return x.numel() if axis is None else x.shape[axis]


def shape(x):
    return tuple(T(x).shape)


def ndim(x):
    return T(x).dim()


def triu(x, k=0):
    return _torch.triu(T(x), diagonal=k)


def diag(x):
    return _torch.diag(T(x))


def eye(n, dtype=None):
    return _torch.eye(n, dtype=to_dtype(dtype) or fdt())


def pad(x, pad_width, mode="constant", constant_values=0):
    x = T(x)
    pw = _np.asarray(pad_width)
    if pw.ndim == 0:
        pw = _np.tile(pw, (x.dim(), 2))
    elif pw.ndim == 1:
        pw = _np.tile(pw, (x.dim(), 1))
    flat = []
    for lo, hi in pw[::-1]:
        flat += [int(lo), int(hi)]
    return _torch.nn.functional.pad(x, flat, value=constant_values)


def meshgrid(*xs, indexing="xy"):
    return _torch.meshgrid(*[T(x) for x in xs], indexing=indexing)


def unique(x, **kw):
    return _torch.unique(T(x))


def count_nonzero(x):
    return _torch.count_nonzero(T(x))


def nonzero(x):
    return _torch.nonzero(T(x), as_tuple=True)


def result_type(*a):
    dt = None
    for x in a:
        t = to_dtype(x) if not isinstance(x, _torch.Tensor) else x.dtype
        dt = t if dt is None else _torch.promote_types(dt, t)
    return dt


def issubdtype(a, b):
    a = to_dtype(a)
    if b in (_np.integer, int):
        return not a.is_floating_point and a != _torch.bool
    if b in (_np.floating, float):
        return a.is_floating_point
    return a == to_dtype(b)


integer = _np.integer
floating = _np.floating
