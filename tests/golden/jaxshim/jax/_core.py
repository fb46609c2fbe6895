"""Torch-backed stand-in for the slice of the jax API that the reference uses.

TEST INFRASTRUCTURE ONLY.  jax/jaxlib are not installable in the build
container, so the unmodified reference sources under /root/reference are
executed by importing them against this package instead: `jax.numpy` maps
to torch CPU tensor functions, `vmap`/`grad` to `torch.func`, `segment_sum`
to an ordered index_add.  It exists for one purpose: producing the golden
vectors of tests/golden/ref_*.npz from the reference's own code
(tests/golden/make_reference_golden.py).  It is never imported by the
product, by bench.py or by the GPU tests.

Semantics of jax that the reference relies on and that are reproduced here:
  * default dtype float32/int32 unless `jax_enable_x64` is set;
  * Python scalars are weakly typed (torch has the same rule);
  * out-of-range gather indices are clamped, out-of-range segment ids and
    scatter indices are dropped;
  * `jnp.mod` takes the sign of the divisor (fmod + fix-up);
  * `x += y` on an array rebinds instead of mutating;
  * subgradient of `maximum` at a tie is 1/2.
"""

import builtins
import math

import numpy as np
import torch

_X64 = [False]


def set_x64(flag):
    _X64[0] = bool(flag)
    torch.set_default_dtype(torch.float64 if flag else torch.float32)


def fdt():
    return torch.float64 if _X64[0] else torch.float32


def idt():
    return torch.int64 if _X64[0] else torch.int32


class DType:
    """A dtype object that is also a scalar constructor, like jnp.float32."""

    def __init__(self, tdt, name):
        self.t = tdt
        self.name = name
        self.__name__ = name
        self.dtype = np.dtype(name.rstrip("_"))  # lets numpy accept it as a dtype

    def __call__(self, x):
        return torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(self.t)

    def __eq__(self, other):
        return to_dtype(other) == self.t if other is not None else False

    def __hash__(self):
        return hash(self.t)

    def __repr__(self):
        return "jaxshim." + self.name


def to_dtype(d):
    if d is None:
        return None
    if isinstance(d, DType):
        return d.t
    if isinstance(d, torch.dtype):
        return d
    if d is float:
        return fdt()
    if d is int:
        return idt()
    if d is bool:
        return torch.bool
    nd = np.dtype(d)
    t = {
        "float64": torch.float64, "float32": torch.float32, "float16": torch.float16,
        "int64": torch.int64, "int32": torch.int32, "int16": torch.int16,
        "int8": torch.int8, "uint8": torch.uint8, "bool": torch.bool,
        "uint32": torch.int64, "uint64": torch.int64, "uint16": torch.int32,
    }[nd.name]
    return canon(t)


def canon(t):
    """Canonicalise a torch dtype the way jax does without x64."""
    if not _X64[0]:
        if t == torch.float64:
            return torch.float32
        if t == torch.int64:
            return torch.int32
    return t


def T(x, dtype=None):
    """Anything array-like -> torch tensor with jax's dtype canonicalisation."""
    if isinstance(x, torch.Tensor):
        out = x
    elif isinstance(x, (bool, np.bool_)):
        out = torch.tensor(bool(x))
    elif isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        out = torch.tensor(int(x), dtype=idt())
    elif isinstance(x, (float, np.floating)):
        out = torch.tensor(float(x), dtype=fdt())
    elif isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in x):
        out = torch.stack([T(e) for e in x])
    else:
        a = np.asarray(x)
        if a.dtype == object:
            raise TypeError("cannot convert object array")
        if a.dtype.kind == "u" and a.dtype.itemsize > 1:
            a = a.astype(np.int64)
        out = torch.from_numpy(np.ascontiguousarray(a).copy())
        out = out.to(canon(out.dtype))
    if dtype is not None:
        out = out.to(to_dtype(dtype))
    return out


def is_scalar(x):
    return isinstance(x, (int, float, bool, np.generic))


def binop_args(a, b):
    """Convert the two operands of a jnp binary function; keep Python scalars weak."""
    if is_scalar(a) and is_scalar(b):
        return T(a), T(b)
    if is_scalar(a):
        b = T(b)
        return _weak(a, b), b
    if is_scalar(b):
        a = T(a)
        return a, _weak(b, a)
    return T(a), T(b)


def _weak(s, ref):
    if isinstance(s, np.generic):
        s = s.item()
    if isinstance(s, float) and not ref.dtype.is_floating_point:
        return torch.tensor(s, dtype=fdt())
    if isinstance(s, bool):
        return torch.tensor(s)
    return torch.tensor(s, dtype=ref.dtype)


# --------------------------------------------------------------------------
# torch.Tensor patches that give tensors the jax.Array surface the reference uses


class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtRef(self.arr, idx)


def _fix_index(idx, shape0=None):
    if isinstance(idx, tuple):
        return tuple(_fix_index(i) for i in idx)
    if isinstance(idx, np.ndarray):
        idx = torch.from_numpy(idx)
    if isinstance(idx, torch.Tensor) and idx.dtype not in (torch.bool, torch.int64):
        idx = idx.long()
    return idx


class _AtRef:
    def __init__(self, arr, idx):
        self.arr = arr
        self.idx = idx

    def _prep(self, val):
        out = self.arr.clone()
        idx = _fix_index(self.idx)
        val = T(val).to(out.dtype) if not isinstance(val, torch.Tensor) else val.to(out.dtype)
        # jax drops out-of-bounds scatter indices (mode="drop" is the default for .at[].set)
        if isinstance(idx, torch.Tensor) and idx.dtype == torch.int64 and out.dim() >= 1:
            n = out.shape[0]
            ok = (idx >= -n) & (idx < n)
            if not bool(ok.all()):
                if val.dim() > 0 and val.shape[:idx.dim()] == idx.shape:
                    val = val[ok]
                idx = idx[ok]
        return out, idx, val

    def set(self, val, **kw):
        out, idx, val = self._prep(val)
        out[idx] = val
        return out

    def add(self, val, **kw):
        out, idx, val = self._prep(val)
        if isinstance(idx, torch.Tensor) and idx.dtype == torch.int64:
            out.index_put_((idx,), val.expand(idx.shape + out.shape[1:]) if val.dim() == 0 else val,
                           accumulate=True)
        else:
            out[idx] = out[idx] + val
        return out


_orig_getitem = torch.Tensor.__getitem__


def _getitem(self, idx):
    if isinstance(idx, np.ndarray):
        idx = torch.from_numpy(idx)
    if isinstance(idx, torch.Tensor) and idx.dtype != torch.bool and self.dim() >= 1:
        n = self.shape[0]
        idx = idx.long().clamp(-n, n - 1)  # jax gather clamps
    elif isinstance(idx, tuple):
        idx = tuple(i.long() if isinstance(i, torch.Tensor) and i.dtype not in (torch.bool, torch.int64)
                    else (torch.from_numpy(i) if isinstance(i, np.ndarray) else i) for i in idx)
    return _orig_getitem(self, idx)


def _patch_binops():
    names = ["add", "sub", "mul", "truediv", "floordiv", "pow", "mod",
             "lt", "le", "gt", "ge", "eq", "ne", "and", "or", "matmul"]
    for n in names:
        for pre in ("__", "__r"):
            meth = pre + n + "__"
            orig = getattr(torch.Tensor, meth, None)
            if orig is None:
                continue

            def make(orig):
                def f(self, other):
                    if isinstance(other, (np.ndarray, np.generic, list)):
                        other = T(other)
                    elif isinstance(other, DType):
                        return NotImplemented
                    return orig(self, other)
                return f

            setattr(torch.Tensor, meth, make(orig))
    # augmented assignment rebinds (jax arrays are immutable)
    for n in ["add", "sub", "mul", "truediv"]:
        def make_i(n):
            def f(self, other):
                return getattr(self, "__" + n + "__")(other)
            return f
        setattr(torch.Tensor, "__i" + n + "__", make_i(n))


_orig_mod = None


def _tensor_mod(self, other):
    return jmod(self, other)


def jmod(a, b):
    a, b = binop_args(a, b)
    if a.dtype.is_floating_point or b.dtype.is_floating_point:
        return torch.remainder(a, b)
    return torch.remainder(a, b)


def install():
    if getattr(torch.Tensor, "_jaxshim", False):
        return
    torch.Tensor._jaxshim = True
    torch.Tensor.at = property(lambda self: _AtIndexer(self))
    torch.Tensor.astype = lambda self, d: self.to(to_dtype(d))
    torch.Tensor.__getitem__ = _getitem
    torch.Tensor.block_until_ready = lambda self: self
    torch.Tensor.copy = lambda self: self.clone()
    _orig_array = torch.Tensor.__array__

    _patch_binops()
    # numpy-style keyword names on the methods the reference calls
    for name in ("sum", "min", "max", "mean", "any", "all", "prod"):
        orig = getattr(torch.Tensor, name)

        def make(orig, name):
            def f(self, *a, axis=None, keepdims=False, **kw):
                if a and axis is None:
                    axis, a = a[0], a[1:]
                if axis is None:
                    return orig(self)
                r = orig(self, dim=axis, keepdim=keepdims)
                return r.values if name in ("min", "max") else r
            return f

        setattr(torch.Tensor, name, make(orig, name))
    # np.argsort(x) on an array object calls x.argsort(axis=..., kind=..., order=...): numpy's
    # sort of a jax array is what the reference's tests rely on (stable for these sizes)
    orig_argsort = torch.Tensor.argsort

    def argsort(self, *a, axis=None, kind=None, order=None, stable=None, **kw):
        if a or kw:
            return orig_argsort(self, *a, stable=True, **kw)
        return orig_argsort(self, dim=-1 if axis is None else axis, stable=True)

    torch.Tensor.argsort = argsort
    orig_transpose = torch.Tensor.transpose

    def transpose(self, *axes):  # numpy / jax: x.transpose(1, 2, 0) or x.transpose((1, 2, 0))
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if len(axes) == 0:
            return self.permute(*reversed(range(self.dim())))
        if len(axes) == self.dim() and len(axes) != 2:
            return self.permute(*axes)
        if len(axes) == 2 and self.dim() == 2 and set(axes) == {0, 1}:
            return self.permute(*axes)
        return orig_transpose(self, *axes)

    torch.Tensor.transpose = transpose
    orig_reshape = torch.Tensor.reshape
    torch.Tensor.ravel = lambda self: orig_reshape(self, -1)
    torch.Tensor.item_ = torch.Tensor.item
    set_x64(False)
