"""jax.numpy stand-in over NumPy (see _core.py)."""
import builtins

import numpy as np

from ._core import (Array, _is_py, _kind, _raw, _seq_sum, canon_dtype, mean_, result_dtype, sum_, wrap, _cast)

ndarray = Array
newaxis = None
inf = np.inf
nan = np.nan
pi = np.pi


class _DType:
    """Callable dtype object: jnp.int32(3) -> 0-d Array; usable wherever a dtype is expected."""

    def __init__(self, dt):
        self.dtype = np.dtype(dt)
        self.__name__ = self.dtype.name

    def __call__(self, x):
        with np.errstate(all="ignore"):
            return wrap(np.asarray(_raw(x) if not _is_py(x) else x).astype(self.dtype))

    def __repr__(self):
        return f"jnp.{self.dtype.name}"

    def __eq__(self, o):
        try:
            return self.dtype == np.dtype(getattr(o, "dtype", o))
        except TypeError:
            return False

    def __hash__(self):
        return hash(self.dtype)


int8, int16, int32, int64 = _DType(np.int8), _DType(np.int16), _DType(np.int32), _DType(np.int32)
uint8, uint16, uint32, uint64 = _DType(np.uint8), _DType(np.uint16), _DType(np.uint32), _DType(np.uint32)
float16, float32, float64 = _DType(np.float16), _DType(np.float32), _DType(np.float32)
bool_ = _DType(np.bool_)


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, _DType):
        return dtype.dtype
    if dtype is int:
        return np.dtype(np.int32)
    if dtype is float:
        return np.dtype(np.float32)
    if dtype is bool:
        return np.dtype(bool)
    return canon_dtype(dtype)


def _shape(s):
    if isinstance(s, (int, np.integer)) or (isinstance(s, np.ndarray) and s.ndim == 0):
        return (int(s),)
    return tuple(int(x) for x in s)


def array(x, dtype=None, copy=True, ndmin=0):
    a = _raw(x)
    if dtype is not None:
        with np.errstate(all="ignore"):
            a = a.astype(_dt(dtype))
    if ndmin:
        a = np.array(a, ndmin=ndmin)
    return wrap(a.copy() if copy else a)


def asarray(x, dtype=None):
    return array(x, dtype, copy=False)


def zeros(shape, dtype=None):
    return wrap(np.zeros(_shape(shape), _dt(dtype) or np.float32))


def ones(shape, dtype=None):
    return wrap(np.ones(_shape(shape), _dt(dtype) or np.float32))


def empty(shape, dtype=None):
    return zeros(shape, dtype)


def full(shape, fill_value, dtype=None):
    fv = fill_value if _is_py(fill_value) else _raw(fill_value)
    dt = _dt(dtype) or result_dtype(fv)
    with np.errstate(all="ignore"):
        return wrap(np.full(_shape(shape), np.asarray(fv).astype(dt), dt))


def zeros_like(a, dtype=None):
    a = _raw(a)
    return wrap(np.zeros(a.shape, _dt(dtype) or a.dtype))


def ones_like(a, dtype=None):
    a = _raw(a)
    return wrap(np.ones(a.shape, _dt(dtype) or a.dtype))


def full_like(a, fill_value, dtype=None):
    a = _raw(a)
    return full(a.shape, fill_value, _dt(dtype) or a.dtype)


def arange(start, stop=None, step=None, dtype=None):
    args = [int(x) if isinstance(x, (np.ndarray, np.integer)) and np.asarray(x).dtype.kind in "iu" else x
            for x in (start, stop, step) if x is not None]
    a = np.arange(*args)
    return wrap(a.astype(_dt(dtype)) if dtype is not None else a)


def where(condition, x=None, y=None, *, size=None, fill_value=None):
    c = _raw(condition)
    if x is None and y is None:
        idx = np.nonzero(c)
        if size is None:
            return tuple(wrap(i) for i in idx)
        fv = 0 if fill_value is None else fill_value
        out = []
        for d, i in enumerate(idx):
            f = fv[d] if isinstance(fv, (tuple, list)) else fv
            r = np.full(size, f, np.int32)
            k = builtins.min(size, i.shape[0])
            r[:k] = i[:k]
            out.append(wrap(r))
        return tuple(out)
    xs = [v if _is_py(v) else _raw(v) for v in (x, y)]
    dt = result_dtype(*xs)
    return wrap(np.where(c.astype(bool), _cast(xs[0], dt), _cast(xs[1], dt)))


def select(condlist, choicelist, default=0):
    return wrap(np.select([_raw(c) for c in condlist], [_raw(c) for c in choicelist], default))


def _promote_list(arrs):
    xs = [v if _is_py(v) else _raw(v) for v in arrs]
    dt = result_dtype(*xs)
    return [np.asarray(_cast(v, dt)) for v in xs]


def concatenate(arrs, axis=0, dtype=None):
    r = np.concatenate(_promote_list(list(arrs)), axis=axis)
    return wrap(r.astype(_dt(dtype)) if dtype is not None else r)


def stack(arrs, axis=0, dtype=None):
    r = np.stack(_promote_list(list(arrs)), axis=axis)
    return wrap(r.astype(_dt(dtype)) if dtype is not None else r)


def vstack(arrs):
    return wrap(np.vstack(_promote_list(list(arrs))))


def hstack(arrs):
    return wrap(np.hstack(_promote_list(list(arrs))))


def _u(ufunc):
    def f(*a, **k):
        return ufunc(*[wrap(x) if not _is_py(x) else x for x in a], **k) if builtins.any(not _is_py(x) for x in a) \
            else ufunc(wrap(a[0]), *a[1:], **k)
    f.__name__ = ufunc.__name__
    return f


abs = absolute = _u(np.absolute)
maximum, minimum = _u(np.maximum), _u(np.minimum)
sign = _u(np.sign)
ceil, floor = _u(np.ceil), _u(np.floor)
exp, log, sqrt = _u(np.exp), _u(np.log), _u(np.sqrt)
logical_and, logical_or, logical_not = _u(np.logical_and), _u(np.logical_or), _u(np.logical_not)
add, subtract, multiply, divide, true_divide = _u(np.add), _u(np.subtract), _u(np.multiply), _u(np.true_divide), _u(np.true_divide)
floor_divide, mod, remainder, power, negative = _u(np.floor_divide), _u(np.remainder), _u(np.remainder), _u(np.power), _u(np.negative)
equal, not_equal, less, less_equal, greater, greater_equal = (_u(np.equal), _u(np.not_equal), _u(np.less),
                                                             _u(np.less_equal), _u(np.greater), _u(np.greater_equal))
isnan, isinf, isfinite = _u(np.isnan), _u(np.isinf), _u(np.isfinite)


def divmod(a, b):
    return np.divmod(wrap(a) if not _is_py(a) else a, wrap(b) if not _is_py(b) else b)


def round(a, decimals=0):
    a = _raw(a)
    if a.dtype.kind != "f":
        return wrap(a)
    return wrap(np.round(a, decimals))      # half to even, as jnp.round


around = round


def clip(a, a_min=None, a_max=None, *, min=None, max=None):
    lo = a_min if a_min is not None else min
    hi = a_max if a_max is not None else max
    r = wrap(a)
    if lo is not None:
        r = maximum(r, lo)
    if hi is not None:
        r = minimum(r, hi)
    return r


def sum(a, axis=None, dtype=None, keepdims=False, where=None):
    return sum_(a, axis=axis, dtype=dtype, keepdims=keepdims, where=where)


def mean(a, axis=None, dtype=None, keepdims=False):
    return mean_(a, axis=axis, keepdims=keepdims)


def std(a, axis=None, **kw):
    return wrap(np.std(_raw(a).astype(np.float32), axis=axis))


def percentile(a, q, axis=None, **kw):
    return wrap(np.percentile(_raw(a).astype(np.float32), q, axis=axis))


def max(a, axis=None, keepdims=False, initial=None, where=None):
    kw = {}
    if initial is not None:
        kw["initial"] = initial
    if where is not None:
        kw["where"] = _raw(where)
    return wrap(np.max(_raw(a), axis=axis, keepdims=keepdims, **kw))


def min(a, axis=None, keepdims=False, initial=None, where=None):
    kw = {}
    if initial is not None:
        kw["initial"] = initial
    if where is not None:
        kw["where"] = _raw(where)
    return wrap(np.min(_raw(a), axis=axis, keepdims=keepdims, **kw))


amax, amin = max, min


def any(a, axis=None, keepdims=False):
    return wrap(np.any(_raw(a), axis=axis, keepdims=keepdims))


def all(a, axis=None, keepdims=False):
    return wrap(np.all(_raw(a), axis=axis, keepdims=keepdims))


def argmax(a, axis=None):
    return wrap(np.argmax(_raw(a), axis=axis))


def argmin(a, axis=None):
    return wrap(np.argmin(_raw(a), axis=axis))


def argsort(a, axis=-1, stable=True, descending=False, **kw):
    a = _raw(a)
    if descending:
        return wrap(np.flip(a.shape[axis] - 1 - np.argsort(np.flip(a, axis), axis=axis, kind="stable"), axis))
    return wrap(np.argsort(a, axis=axis, kind="stable"))


def sort(a, axis=-1, **kw):
    return wrap(np.sort(_raw(a), axis=axis, kind="stable"))


def cumsum(a, axis=None, dtype=None):
    a = _raw(a)
    return wrap(np.cumsum(a, axis=axis, dtype=_dt(dtype) or (a.dtype if a.dtype.kind != "b" else np.int32)))


def dot(a, b):
    a, b = _promote_list([a, b])
    if a.dtype.kind == "f" and a.ndim == 1 and b.ndim == 1:
        return wrap(_seq_sum(a * b, None))
    return wrap(np.dot(a, b))


def unique(a, return_index=False, return_inverse=False, return_counts=False, axis=None, *, size=None, fill_value=None):
    a = _raw(a)
    res = np.unique(a, return_index=return_index, return_inverse=return_inverse, return_counts=return_counts, axis=axis)
    multi = isinstance(res, tuple)
    vals = res[0] if multi else res
    if size is not None:
        k = builtins.min(size, vals.shape[0])
        fv = vals[0] if fill_value is None else fill_value     # jnp: default fill is the minimum value
        out = np.full((size,) + vals.shape[1:], fv, vals.dtype)
        out[:k] = vals[:k]
        vals = out
    if not multi:
        return wrap(vals)
    rest = list(res[1:])
    fixed = []
    names = [n for n, f in (("index", return_index), ("inverse", return_inverse), ("counts", return_counts)) if f]
    for n, r in zip(names, rest):
        if size is not None and n != "inverse":
            o = np.zeros((size,), r.dtype)
            k = builtins.min(size, r.shape[0])
            o[:k] = r[:k]
            r = o
        if n == "inverse":
            r = r.reshape(a.shape) if axis is None else r
        fixed.append(wrap(r))
    return (wrap(vals), *fixed)


def resize(a, new_shape):
    return wrap(np.resize(_raw(a), _shape(new_shape)))


def reshape(a, shape, *a2, **k):
    return wrap(_raw(a).reshape(_shape(shape)))


def ravel(a):
    return wrap(_raw(a).reshape(-1))


def squeeze(a, axis=None):
    return wrap(np.squeeze(_raw(a), axis=axis))


def expand_dims(a, axis):
    return wrap(np.expand_dims(_raw(a), axis))


def transpose(a, axes=None):
    return wrap(np.transpose(_raw(a), axes))


def tile(a, reps):
    return wrap(np.tile(_raw(a), reps))


def repeat(a, repeats, axis=None, total_repeat_length=None):
    return wrap(np.repeat(_raw(a), _raw(repeats) if not _is_py(repeats) else repeats, axis=axis))


def atleast_1d(a):
    return wrap(np.atleast_1d(_raw(a)))


def flip(a, axis=None):
    return wrap(np.flip(_raw(a), axis))


def take(a, indices, axis=None, **kw):
    return wrap(np.take(_raw(a), _raw(indices), axis=axis, mode="clip"))


def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    a = _raw(x)
    if a.dtype.kind != "f":
        return wrap(a)
    return wrap(np.nan_to_num(a, nan=nan, posinf=posinf, neginf=neginf))


def isscalar(x):
    return np.isscalar(x) or (isinstance(x, np.ndarray) and x.ndim == 0)


def set_printoptions(*a, **k):
    np.set_printoptions(*a, **k)


def iinfo(dt):
    return np.iinfo(_dt(dt))


def finfo(dt):
    return np.finfo(_dt(dt))

int_ = int32
float_ = float32
uint = uint32
integer = np.integer
floating = np.floating
number = np.number
dtype = np.dtype
bool = bool_  # noqa: A001
