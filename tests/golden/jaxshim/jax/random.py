"""jax.random stand-in: deterministic, NOT threefry.  Every draw is appended to TRACE as
(function, caller function name, key, args, result) so a golden-vector generator can replay the same draws."""
import sys

import numpy as np

from . import numpy as jnp
from ._core import _raw, wrap

TRACE = []
_RECORD = True


def _rng(key, tag):
    k = _raw(key).astype(np.uint32).reshape(-1)
    return np.random.default_rng(np.concatenate([k.astype(np.uint64), np.array([tag], np.uint64)]))


def _caller():
    f = sys._getframe(2)
    return f"{f.f_code.co_name}:{f.f_lineno}"


def PRNGKey(seed):
    s = int(np.asarray(seed)) & 0xFFFFFFFFFFFFFFFF
    return wrap(np.array([s >> 32, s & 0xFFFFFFFF], np.uint32))


key = PRNGKey


def split(key, num=2):
    r = _rng(key, 1)
    return wrap(r.integers(0, 2**32, size=(int(num), 2), dtype=np.uint64).astype(np.uint32))


def fold_in(key, data):
    r = _rng(key, 2 + (int(np.asarray(data)) & 0xFFFFFF) * 16)
    return wrap(r.integers(0, 2**32, size=(2,), dtype=np.uint64).astype(np.uint32))


def randint(key, shape, minval, maxval, dtype=None):
    r = _rng(key, 3)
    lo, hi = int(np.asarray(minval)), int(np.asarray(maxval))
    out = wrap(r.integers(lo, max(hi, lo + 1), size=tuple(shape)).astype(np.int32))
    TRACE.append(("randint", _caller(), _raw(key).copy(), (tuple(shape), lo, hi), _raw(out).copy()))
    return out


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    r = _rng(key, 4)
    return wrap((r.random(size=tuple(shape)) * (maxval - minval) + minval).astype(np.float32))


def normal(key, shape=(), dtype=None):
    return wrap(_rng(key, 5).standard_normal(size=tuple(shape)).astype(np.float32))


def permutation(key, x, axis=0, independent=False):
    r = _rng(key, 6)
    if isinstance(x, (int, np.integer)) or np.asarray(x).ndim == 0:
        n = int(np.asarray(x))
        perm = r.permutation(n).astype(np.int32)
        TRACE.append(("permutation", _caller(), _raw(key).copy(), (n,), perm.copy()))
        return wrap(perm)
    a = _raw(x)
    perm = r.permutation(a.shape[axis]).astype(np.int32)
    TRACE.append(("permutation", _caller(), _raw(key).copy(), (a.shape[axis],), perm.copy()))
    return wrap(np.take(a, perm, axis=axis))


def choice(key, a, shape=(), replace=True, p=None, axis=0):
    """jax.random.choice.  With ``p`` (scalar draw, replace=True) it follows jax/_src/random.py literally:
    ``p_cuml = cumsum(p); r = p_cuml[-1] * (1 - uniform(key)); ind = searchsorted(p_cuml, r)`` in float32, and the
    uniform draw is recorded (args = (outermost vmap index, innermost scan index)) so that a replay can feed it back."""
    from . import lax as _lax
    from . import _core
    r = _rng(key, 7)
    arr = _raw(a)
    if arr.ndim == 0:
        arr = np.arange(int(arr))
    if p is not None and tuple(shape) == () and replace:
        u = np.float32(int(r.integers(0, 2 ** 23)) / 2.0 ** 23)          # [0, 1), 23 mantissa bits as jax.random.uniform
        p_cuml = np.cumsum(_raw(p).astype(np.float32), dtype=np.float32)
        rr = np.float32(p_cuml[-1] * np.float32(np.float32(1.0) - u))
        idx = min(int(np.searchsorted(p_cuml, rr, side="left")), arr.shape[0] - 1)
        pos = (_core.VMAP_STACK[0] if _core.VMAP_STACK else -1, _lax.SCAN_STACK[-1] if _lax.SCAN_STACK else -1)
        TRACE.append(("choice", _caller(), _raw(key).copy(), pos, float(u)))
        return wrap(arr[idx])
    pp = None
    if p is not None:
        pp = _raw(p).astype(np.float64)
        pp = pp / pp.sum() if pp.sum() > 0 else None
    idx = r.choice(arr.shape[0], size=tuple(shape) if shape != () else None, replace=replace, p=pp)
    out = wrap(arr[idx])
    TRACE.append(("choice", _caller(), _raw(key).copy(), (arr.shape[0],), np.asarray(idx).copy()))
    return out


def bernoulli(key, p=0.5, shape=None):
    r = _rng(key, 8)
    return wrap(r.random(size=tuple(shape or ())) < float(np.asarray(p)))
