"""NumPy-backed stand-in for the `jax` package (TEST INFRASTRUCTURE; see _core.py for the emulated semantics)."""
import functools
import types

import numpy as _np

from . import _core, flatten_util, lax, numpy, random, tree_util
from ._core import Array, wrap as _wrap
from .tree_util import tree_flatten, tree_leaves, tree_map, tree_unflatten

__version__ = "0.0.shim"
SHIM = True
tree = tree_util
typing = types.SimpleNamespace(ArrayLike=object, DTypeLike=object)


def _to_arrays(x):
    if isinstance(x, (_np.ndarray, _np.generic)):
        return _wrap(x)
    return x


def jit(fun=None, static_argnums=None, static_argnames=None, **kw):
    """Eager: numpy inputs are canonicalised to (32-bit) Arrays as tracing would."""
    if fun is None:
        return lambda f: jit(f, static_argnums=static_argnums, static_argnames=static_argnames, **kw)
    if isinstance(static_argnums, int):
        static_argnums = (static_argnums,)
    static = set(static_argnums or ())

    @functools.wraps(fun)
    def wrapped(*args, **kwargs):
        args = tuple(a if i in static else tree_map(_to_arrays, a) for i, a in enumerate(args))
        return fun(*args, **kwargs)
    return wrapped


def vmap(fun, in_axes=0, out_axes=0, axis_name=None, **kw):
    """Python loop over the mapped axis; outputs are stacked leaf-wise along out_axes."""
    @functools.wraps(fun)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("vmap in_axes must match the arguments")
        n = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            if isinstance(ax, (tuple, list, dict)):
                raise NotImplementedError("nested in_axes")
            for leaf in tree_leaves(a):
                n = int(_np.asarray(leaf).shape[ax])
                break
            if n is not None:
                break
        if n is None:
            raise ValueError("vmap needs at least one mapped argument")
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree_map(lambda l: _wrap(_np.take(_core._raw(l), i, axis=ax)), a)
                  for a, ax in zip(args, axes)]
            _core.VMAP_STACK.append(i)
            try:
                outs.append(fun(*sl))
            finally:
                _core.VMAP_STACK.pop()
        leaves0, td = tree_flatten(outs[0])
        cols = [[] for _ in leaves0]
        for o in outs:
            lv, _ = tree_flatten(o)
            for c, v in zip(cols, lv):
                c.append(v)
        oax = out_axes if isinstance(out_axes, int) else 0
        return tree_unflatten(td, [numpy.stack(c, axis=oax) for c in cols])
    return mapped


vmapped = vmap


def block_until_ready(x):
    return x


def devices(*a, **k):
    return [types.SimpleNamespace(platform="cpu", id=0, device_kind="numpy-shim")]


def device_count(*a, **k):
    return 1


def local_device_count(*a, **k):
    return 1


def default_backend():
    return "cpu"


def device_put(x, *a, **k):
    return tree_map(_to_arrays, x)


def device_get(x):
    return x


def named_scope(*a, **k):
    import contextlib
    return contextlib.nullcontext()


class _Config:
    def update(self, *a, **k):
        pass

    jax_enable_x64 = False


config = _Config()
debug = types.SimpleNamespace(print=lambda *a, **k: None, callback=lambda *a, **k: None, breakpoint=lambda *a, **k: None)


class _Profiler:
    @staticmethod
    def start_trace(*a, **k):
        pass

    @staticmethod
    def stop_trace(*a, **k):
        pass

    @staticmethod
    def trace(*a, **k):
        import contextlib
        return contextlib.nullcontext()

    @staticmethod
    def save_device_memory_profile(*a, **k):
        pass


profiler = _Profiler()
ops = types.SimpleNamespace()


def _segment_sum(data, segment_ids, num_segments=None, **kw):
    d, s = _core._raw(data), _core._raw(segment_ids)
    n = int(num_segments) if num_segments is not None else int(s.max()) + 1
    out = _np.zeros((n,) + d.shape[1:], d.dtype)
    ok = (s >= 0) & (s < n)
    _np.add.at(out, s[ok], d[ok])
    return _wrap(out)


ops.segment_sum = _segment_sum
