"""NumPy-backed emulation of the slice of the JAX API that JaxMARL-HFT's environment uses (TEST INFRASTRUCTURE).

Purpose: execute the UNMODIFIED reference sources (/root/reference/gymnax_exchange) in a container without jax/jaxlib,
to generate golden vectors that pin the CPU oracle.  It emulates JAX *semantics*, not XLA numerics:

  * x64 disabled: every array is canonicalised to <= 32 bits; int32 arithmetic wraps;
  * weak typing: Python scalars never widen an array dtype; int (+) Python float -> float32;
  * int / int -> float32 true division with BOTH operands converted to float32 first;
  * float32 floor_divide / remainder follow jax._src.numpy.ufuncs._float_divmod;
  * gather indices: negative wraps once, then clamps; scatter out of bounds is dropped;
  * jnp.where(mask, size=, fill_value=), jnp.unique(size=, fill_value=), stable argsort;
  * lax.cond / switch / while_loop / scan and vmap run eagerly (vmap = Python loop over the batch + stack), which is
    observationally equal to the traced program for pure functions;
  * float32 reductions are summed left-to-right (XLA leaves the order unspecified).
jax.random is a deterministic stand-in (NOT threefry): every draw is recorded in ``random.TRACE`` so the golden
generator can hand the same draws to the oracle as inputs.
"""
import builtins
import operator

import numpy as np

_PY_SCALARS = (bool, int, float)


def canon_dtype(dt):
    dt = np.dtype(dt)
    if dt == np.float64:
        return np.dtype(np.float32)
    if dt == np.int64:
        return np.dtype(np.int32)
    if dt == np.uint64:
        return np.dtype(np.uint32)
    if dt == np.complex128:
        return np.dtype(np.complex64)
    return dt


def _is_py(x):
    return type(x) in _PY_SCALARS


def _raw(x):
    """-> plain ndarray with a canonical dtype (wraps int64 -> int32 silently, like jnp.asarray with x64 off)."""
    if isinstance(x, np.ndarray):
        a = x.view(np.ndarray) if type(x) is not np.ndarray else x
    else:
        if isinstance(x, (list, tuple)):
            x = [(_raw(e) if not _is_py(e) else e) for e in x]
        a = np.asarray(x)
    cd = canon_dtype(a.dtype)
    if cd != a.dtype:
        with np.errstate(all="ignore"):
            a = a.astype(cd)
    return a


def _kind(dt):
    k = np.dtype(dt).kind
    return 0 if k == "b" else (1 if k in "iu" else 2)


def result_dtype(*xs):
    """JAX type promotion for the dtypes that occur here (bool < ints < float32; Python scalars are weak)."""
    strong = [canon_dtype(x.dtype) for x in xs if not _is_py(x)]
    weak = [type(x) for x in xs if _is_py(x)]
    if strong:
        if any(_kind(d) == 2 for d in strong):
            fl = [d for d in strong if _kind(d) == 2]
            dt = fl[0]
            for d in fl[1:]:
                dt = np.promote_types(dt, d)
            dt = canon_dtype(dt)
        elif any(_kind(d) == 1 for d in strong):
            ints = [d for d in strong if _kind(d) == 1]
            dt = ints[0]
            for d in ints[1:]:
                dt = canon_dtype(np.promote_types(dt, d))
        else:
            dt = np.dtype(bool)
        if float in weak and _kind(dt) < 2:
            dt = np.dtype(np.float32)
        elif int in weak and _kind(dt) == 0:
            dt = np.dtype(np.int32)
        return dt
    if float in weak:
        return np.dtype(np.float32)
    if int in weak:
        return np.dtype(np.int32)
    return np.dtype(bool)


def _cast(x, dt):
    with np.errstate(all="ignore"):
        if _is_py(x):
            if _kind(dt) == 1 and isinstance(x, int) and not isinstance(x, bool):
                return np.array(x).astype(dt)      # wraps like lax.convert_element_type
            return np.asarray(x, dtype=dt)
        return x if x.dtype == dt else x.astype(dt)


_INEXACT = {np.true_divide, np.exp, np.log, np.sqrt, np.ceil, np.floor, np.exp2, np.log2, np.log10, np.expm1,
            np.log1p, np.sin, np.cos, np.tanh, np.rint, np.trunc}
_COMPARE = {np.equal, np.not_equal, np.less, np.less_equal, np.greater, np.greater_equal}
_LOGICAL = {np.logical_and, np.logical_or, np.logical_not, np.logical_xor}
_KEEP = {np.isnan, np.isinf, np.isfinite, np.signbit}


def _float_divmod(x1, x2):
    """jax._src.numpy.ufuncs._float_divmod (float32)."""
    with np.errstate(all="ignore"):
        mod = np.fmod(x1, x2)
        div = (x1 - mod) / x2
        ind = (np.sign(mod) != np.sign(x2)) & (mod != 0)
        mod = np.where(ind, mod + x2, mod)
        div = np.where(ind, div - np.asarray(1, x1.dtype), div)
        # lax.round: half away from zero
        rdiv = np.where(div >= 0, np.floor(div + 0.5), np.ceil(div - 0.5)).astype(x1.dtype)
    return rdiv, mod.astype(x1.dtype)


def _int_floordiv(a, b):
    """jnp.floor_divide on ints: lax.div (truncating) with the sign fix-up; x // 0 follows XLA (-1)."""
    with np.errstate(all="ignore"):
        bz = b == 0
        bs = np.where(bz, 1, b)
        q = np.floor_divide(a, bs)
        return np.where(bz, np.asarray(-1, q.dtype), q).astype(q.dtype)


def apply_ufunc(ufunc, method, inputs, kwargs):
    ins = [x if _is_py(x) else _raw(x) for x in inputs]
    if method == "__call__":
        if ufunc in _LOGICAL:
            ins = [np.asarray(x).astype(bool) if not _is_py(x) else bool(x) for x in ins]
            return wrap(ufunc(*ins, **kwargs))
        if ufunc in _KEEP:
            return wrap(ufunc(*[np.asarray(x) for x in ins], **kwargs))
        dt = result_dtype(*ins)
        if ufunc in _INEXACT and _kind(dt) < 2:
            dt = np.dtype(np.float32)
        if ufunc in (np.invert, np.bitwise_and, np.bitwise_or, np.bitwise_xor, np.left_shift, np.right_shift):
            ci = [_cast(x, dt) for x in ins]
            return wrap(ufunc(*ci, **kwargs))
        if ufunc in (np.add, np.subtract, np.multiply, np.negative) and _kind(dt) == 0:
            dt = np.dtype(bool) if ufunc is not np.negative else dt   # bool arithmetic stays bool in JAX (add == or)
        ci = [_cast(x, dt) for x in ins]
        with np.errstate(all="ignore"):
            if ufunc is np.floor_divide:
                res = _float_divmod(*ci)[0] if _kind(dt) == 2 else _int_floordiv(*ci)
            elif ufunc in (np.remainder, np.mod):
                if _kind(dt) == 2:
                    res = _float_divmod(*ci)[1]
                else:
                    bz = ci[1] == 0
                    res = np.where(bz, ci[0], np.remainder(ci[0], np.where(bz, 1, ci[1]))).astype(dt)
            elif ufunc is np.divmod:
                if _kind(dt) == 2:
                    res = _float_divmod(*ci)
                else:
                    res = (_int_floordiv(*ci), np.remainder(ci[0], np.where(ci[1] == 0, 1, ci[1])).astype(dt))
            elif ufunc is np.power and _kind(dt) == 1:
                res = np.power(ci[0].astype(np.int64), ci[1].astype(np.int64)).astype(dt)
            else:
                res = ufunc(*ci, **kwargs)
        if isinstance(res, tuple):
            return tuple(wrap(r) for r in res)
        return wrap(res)
    if method == "reduce":
        a = ins[0]
        kw = {k: v for k, v in kwargs.items() if k in ("axis", "keepdims", "initial", "where")}
        if "where" in kw:
            kw["where"] = _raw(kw["where"])
        if ufunc is np.add and a.dtype.kind == "f":
            return wrap(_seq_sum(a, kw.get("axis", 0), kw.get("keepdims", False)))
        with np.errstate(all="ignore"):
            res = ufunc.reduce(a, dtype=(a.dtype if ufunc in (np.add, np.multiply) and a.dtype.kind != "b" else None), **kw)
        return wrap(res)
    if method == "accumulate":
        with np.errstate(all="ignore"):
            return wrap(ufunc.accumulate(ins[0], **{k: v for k, v in kwargs.items() if k in ("axis",)}, dtype=ins[0].dtype))
    raise NotImplementedError(f"ufunc method {method} of {ufunc}")


def _seq_sum(a, axis=None, keepdims=False):
    """float32 sum, strictly left to right along ``axis`` (None = flattened C order)."""
    a = np.asarray(a)
    if axis is None:
        flat = a.reshape(-1)
        r = np.cumsum(flat, dtype=a.dtype)[-1] if flat.size else np.asarray(0, a.dtype)
        return np.asarray(r, a.dtype).reshape((1,) * a.ndim) if keepdims else np.asarray(r, a.dtype)
    if isinstance(axis, tuple):
        out = a
        for ax in sorted((x % a.ndim for x in axis), reverse=True):
            out = _seq_sum(out, ax, keepdims)
        return out
    if a.shape[axis] == 0:
        return np.zeros(np.sum(a, axis=axis, keepdims=keepdims).shape, a.dtype)
    r = np.take(np.cumsum(a, axis=axis, dtype=a.dtype), -1, axis=axis)
    return np.expand_dims(r, axis) if keepdims else r


def _norm_index(idx, shape, clamp):
    """JAX index normalisation for integer indices: negative wraps once; gather clamps, scatter marks OOB (-> None)."""
    tup = idx if isinstance(idx, tuple) else (idx,)
    out, dim, oob = [], 0, False
    n_real = builtins.sum(1 for t in tup if t is not None and t is not Ellipsis)
    for t in tup:
        if t is None:
            out.append(t); continue
        if t is Ellipsis:
            dim += len(shape) - n_real
            out.append(t); continue
        if isinstance(t, slice):
            out.append(t); dim += 1; continue
        if isinstance(t, (list, tuple)):
            t = np.asarray(t)
        if isinstance(t, np.ndarray) and t.dtype == bool:
            out.append(t.view(np.ndarray)); dim += t.ndim; continue
        if isinstance(t, (int, np.integer)) or (isinstance(t, np.ndarray) and t.dtype.kind in "iu"):
            n = shape[dim] if dim < len(shape) else 1
            a = np.asarray(t).astype(np.int64)
            a = np.where(a < 0, a + n, a)
            bad = (a < 0) | (a >= n)
            if bad.any():
                oob = True
                a = np.clip(a, 0, builtins.max(n - 1, 0))
            out.append(int(a) if a.ndim == 0 else a)
            dim += 1
            continue
        out.append(t); dim += 1
    res = tuple(out) if isinstance(idx, tuple) else out[0]
    return res, oob


class Array(np.ndarray):
    """jax.Array stand-in."""
    __array_priority__ = 1000

    def __new__(cls, x):
        return _raw(x).view(cls)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if out is not None:
            raise NotImplementedError("out= is not part of the JAX API")
        return apply_ufunc(ufunc, method, inputs, kwargs)

    def __array_function__(self, func, types, args, kwargs):
        def strip(o):
            if isinstance(o, Array):
                return o.view(np.ndarray)
            if isinstance(o, (list, tuple)):
                return type(o)(strip(e) for e in o)
            if isinstance(o, dict):
                return {k: strip(v) for k, v in o.items()}
            return o
        res = func(*strip(args), **strip(kwargs))
        if isinstance(res, np.ndarray) or isinstance(res, np.generic):
            return wrap(res)
        if isinstance(res, (tuple, list)):
            return type(res)(wrap(r) if isinstance(r, (np.ndarray, np.generic)) else r for r in res)
        return res

    def __getitem__(self, idx):
        if isinstance(idx, Array):
            idx = idx.view(np.ndarray)
        elif isinstance(idx, tuple):
            idx = tuple(i.view(np.ndarray) if isinstance(i, Array) else i for i in idx)
        nidx, _ = _norm_index(idx, self.shape, clamp=True)
        return wrap(self.view(np.ndarray)[nidx])

    def __setitem__(self, idx, v):
        raise TypeError("JAX arrays are immutable; use .at[].set()")

    def __iter__(self):
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (self[i] for i in range(self.shape[0]))

    def __hash__(self):
        return id(self)

    def __bool__(self):
        return bool(self.view(np.ndarray))

    def __index__(self):
        return int(self.view(np.ndarray))

    def __int__(self):
        return int(self.view(np.ndarray))

    def __float__(self):
        return float(self.view(np.ndarray))

    def __format__(self, spec):
        return format(self.view(np.ndarray).item() if self.ndim == 0 else self.view(np.ndarray), spec)

    def __reduce__(self):
        return (Array, (self.view(np.ndarray).copy(),))

    @property
    def at(self):
        return _At(self)

    def astype(self, dtype, *a, **k):
        with np.errstate(all="ignore"):
            return wrap(self.view(np.ndarray).astype(canon_dtype(dtype)))

    def sum(self, axis=None, dtype=None, keepdims=False, **kw):
        return sum_(self, axis=axis, dtype=dtype, keepdims=keepdims)

    def mean(self, axis=None, dtype=None, keepdims=False, **kw):
        return mean_(self, axis=axis, keepdims=keepdims)

    def max(self, axis=None, keepdims=False, **kw):
        return wrap(np.max(self.view(np.ndarray), axis=axis, keepdims=keepdims))

    def min(self, axis=None, keepdims=False, **kw):
        return wrap(np.min(self.view(np.ndarray), axis=axis, keepdims=keepdims))

    def any(self, axis=None, keepdims=False, **kw):
        return wrap(np.any(self.view(np.ndarray), axis=axis, keepdims=keepdims))

    def all(self, axis=None, keepdims=False, **kw):
        return wrap(np.all(self.view(np.ndarray), axis=axis, keepdims=keepdims))

    def reshape(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return wrap(self.view(np.ndarray).reshape(tuple(int(s) for s in shape)))

    def flatten(self, *a, **k):
        return wrap(self.view(np.ndarray).reshape(-1).copy())

    def ravel(self, *a, **k):
        return wrap(self.view(np.ndarray).reshape(-1))

    def squeeze(self, axis=None):
        return wrap(np.squeeze(self.view(np.ndarray), axis=axis))

    def transpose(self, *axes):
        return wrap(self.view(np.ndarray).transpose(*axes))

    @property
    def T(self):
        return wrap(self.view(np.ndarray).T)

    def block_until_ready(self):
        return self

    def item(self, *a):
        return self.view(np.ndarray).item(*a)

    def tolist(self):
        return self.view(np.ndarray).tolist()


def wrap(x):
    if isinstance(x, Array):
        return x
    return _raw(x).view(Array)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        if isinstance(idx, Array):
            idx = idx.view(np.ndarray)
        elif isinstance(idx, tuple):
            idx = tuple(i.view(np.ndarray) if isinstance(i, Array) else i for i in idx)
        self.arr, self.idx = arr, idx

    def _apply(self, v, op):
        base = self.arr.view(np.ndarray).copy()
        nidx, oob = _norm_index(self.idx, base.shape, clamp=False)
        if oob:   # scatter with out-of-bounds indices: those updates are dropped (mode=FILL_OR_DROP)
            tup = self.idx if isinstance(self.idx, tuple) else (self.idx,)
            if all(isinstance(t, (int, np.integer, slice)) or (isinstance(t, np.ndarray) and t.ndim == 0) for t in tup):
                return wrap(base)
            raise NotImplementedError("partially out-of-bounds vector scatter")
        with np.errstate(all="ignore"):
            val = _raw(v) if not _is_py(v) else v
            if op == "set":
                base[nidx] = np.asarray(val).astype(base.dtype) if not _is_py(val) else val
            elif op == "add":
                np.add.at(base, nidx, np.asarray(val).astype(base.dtype))
            elif op == "mul":
                np.multiply.at(base, nidx, np.asarray(val).astype(base.dtype))
            elif op == "min":
                np.minimum.at(base, nidx, np.asarray(val).astype(base.dtype))
            elif op == "max":
                np.maximum.at(base, nidx, np.asarray(val).astype(base.dtype))
        return wrap(base)

    def set(self, v, **kw):
        return self._apply(v, "set")

    def add(self, v, **kw):
        return self._apply(v, "add")

    def multiply(self, v, **kw):
        return self._apply(v, "mul")

    def min(self, v, **kw):
        return self._apply(v, "min")

    def max(self, v, **kw):
        return self._apply(v, "max")

    def get(self, **kw):
        return self.arr[self.idx]


def sum_(a, axis=None, dtype=None, keepdims=False, where=None):
    a = _raw(a)
    if where is not None:
        a = np.where(_raw(where), a, np.zeros((), a.dtype))
    if dtype is not None:
        a = a.astype(canon_dtype(dtype))
    if a.dtype.kind == "b":
        a = a.astype(np.int32)
    if a.dtype.kind == "f":
        return wrap(_seq_sum(a, axis, keepdims))
    with np.errstate(all="ignore"):
        return wrap(np.sum(a, axis=axis, keepdims=keepdims, dtype=a.dtype))


def mean_(a, axis=None, keepdims=False):
    a = _raw(a)
    if a.dtype.kind != "f":
        a = a.astype(np.float32)       # jnp.mean: sum(x, dtype=float32) / n
    s = _seq_sum(a, axis, keepdims)
    n = a.size if axis is None else (np.prod([a.shape[x] for x in axis]) if isinstance(axis, tuple) else a.shape[axis])
    with np.errstate(all="ignore"):
        return wrap((s / np.float32(n)).astype(np.float32))


VMAP_STACK = []   # index of the running element of every active vmap (outermost first)
