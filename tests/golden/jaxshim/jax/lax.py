"""jax.lax stand-in (eager).  Test infrastructure only."""
import torch as _torch

from . import tree_util
from ._core import T, to_dtype


def stop_gradient(x):
    # like every lax primitive, stop_gradient turns a Python scalar into a (weakly typed)
    # array of the default dtype: jax_md's `cutoff ** 2` (jax_md/partition.py:820-821) is
    # therefore evaluated in float32 when x64 is off, not in Python double precision
    def one(t):
        if isinstance(t, _torch.Tensor):
            return t.detach()
        if isinstance(t, (float, int)) and not isinstance(t, bool):
            return T(t)
        return t
    return tree_util.tree_map(one, x)


def iota(dtype, size):
    return _torch.arange(int(size), dtype=to_dtype(dtype))


def cond(pred, *args):
    if len(args) == 4 and callable(args[1]) and callable(args[3]) and not callable(args[0]):
        # legacy form cond(pred, true_operand, true_fun, false_operand, false_fun)
        t_op, t_fn, f_op, f_fn = args
        return t_fn(t_op) if bool(pred) else f_fn(f_op)
    true_fn, false_fn, operands = args[0], args[1], args[2:]
    return true_fn(*operands) if bool(pred) else false_fn(*operands)


def fori_loop(lo, hi, body, init):
    val = init
    for i in range(int(lo), int(hi)):
        val = body(i, val)
    return val


def while_loop(cond_fn, body, init):
    val = init
    while bool(cond_fn(val)):
        val = body(val)
    return val


def scan(f, init, xs, length=None):
    carry = init
    n = length if xs is None else len(tree_util.tree_leaves(xs)[0])
    ys = []
    for i in range(int(n)):
        x = None if xs is None else tree_util.tree_map(lambda t: t[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    if ys and ys[0] is not None:
        ys = tree_util.tree_map(lambda *l: _torch.stack([T(v) for v in l]), *ys)
    else:
        ys = None
    return carry, ys


def dynamic_slice(x, start, sizes):
    x = T(x)
    idx = []
    for d, (s, n) in enumerate(zip(start, sizes)):
        s = int(s)
        s = max(0, min(s, x.shape[d] - int(n)))  # jax clamps the start index
        idx.append(slice(s, s + int(n)))
    return x[tuple(idx)]


def dynamic_update_slice(x, upd, start):
    x = T(x).clone()
    idx = tuple(slice(int(s), int(s) + n) for s, n in zip(start, upd.shape))
    x[idx] = upd
    return x
