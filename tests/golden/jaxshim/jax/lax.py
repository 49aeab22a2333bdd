"""jax.lax stand-in: eager control flow (see _core.py)."""
import numpy as np

from . import numpy as jnp
from ._core import _raw, wrap
from .tree_util import tree_flatten, tree_map, tree_unflatten


def cond(pred, true_fun, false_fun, *operands, **kw):
    return true_fun(*operands) if bool(np.asarray(pred)) else false_fun(*operands)


def switch(index, branches, *operands):
    i = int(np.clip(int(np.asarray(index)), 0, len(branches) - 1))   # lax.switch clamps
    return branches[i](*operands)


def select(pred, on_true, on_false):
    return jnp.where(pred, on_true, on_false)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(np.asarray(cond_fun(val))):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(jnp.int32(i), val)
    return val


SCAN_STACK = []   # index of the running iteration of every active scan (innermost last)


def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = int(length)
    else:
        leaves, _ = tree_flatten(xs)
        n = int(leaves[0].shape[0]) if leaves else int(length)
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        x = None if xs is None else tree_map(lambda a: wrap(a)[i], xs)
        SCAN_STACK.append(i)
        try:
            carry, y = f(carry, x)
        finally:
            SCAN_STACK.pop()
        ys[i] = y
    if n == 0 or ys[0] is None:
        return carry, None
    leaves0, treedef = tree_flatten(ys[0])
    cols = [[] for _ in leaves0]
    for y in ys:
        lv, _ = tree_flatten(y)
        for c, v in zip(cols, lv):
            c.append(v)
    return carry, tree_unflatten(treedef, [jnp.stack(c) for c in cols])


def dynamic_slice_in_dim(operand, start_index, slice_size, axis=0):
    a = _raw(operand)
    n = a.shape[axis]
    s = int(np.asarray(start_index))
    if s < 0:
        s += n
    s = max(0, min(s, n - slice_size))       # XLA clamps so that the slice stays in bounds
    idx = [slice(None)] * a.ndim
    idx[axis] = slice(s, s + slice_size)
    return wrap(a[tuple(idx)].copy())


def dynamic_slice(operand, start_indices, slice_sizes):
    a = _raw(operand)
    idx = []
    for ax, (s, sz) in enumerate(zip(start_indices, slice_sizes)):
        s = int(np.asarray(s))
        if s < 0:
            s += a.shape[ax]
        s = max(0, min(s, a.shape[ax] - sz))
        idx.append(slice(s, s + sz))
    return wrap(a[tuple(idx)].copy())


def dynamic_update_slice(operand, update, start_indices):
    a = _raw(operand).copy()
    u = _raw(update)
    idx = []
    for ax, s in enumerate(start_indices):
        s = int(np.asarray(s))
        if s < 0:
            s += a.shape[ax]
        s = max(0, min(s, a.shape[ax] - u.shape[ax]))
        idx.append(slice(s, s + u.shape[ax]))
    a[tuple(idx)] = u
    return wrap(a)


def stop_gradient(x):
    return x


def bitcast_convert_type(x, new_dtype):
    return wrap(_raw(x).view(jnp._dt(new_dtype)))


def pmean(x, axis_name=None):
    return x


def round(x, rounding_method=0):
    a = _raw(x)
    return wrap(np.where(a >= 0, np.floor(a + 0.5), np.ceil(a - 0.5)).astype(a.dtype))
