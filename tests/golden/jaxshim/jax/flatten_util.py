import numpy as np

from . import numpy as jnp
from ._core import _raw, result_dtype, wrap
from .tree_util import tree_flatten, tree_unflatten


def ravel_pytree(pytree):
    """Leaves raveled and concatenated in flatten order (dict keys sorted), dtype = promotion of all leaves."""
    leaves, td = tree_flatten(pytree)
    raws = [l if isinstance(l, (bool, int, float)) else _raw(l) for l in leaves]
    if not raws:
        return jnp.zeros((0,), jnp.float32), (lambda flat: pytree)
    dt = result_dtype(*raws)
    arrs = [np.asarray(r, dtype=dt if isinstance(r, (bool, int, float)) else None) for r in raws]
    shapes = [a.shape for a in arrs]
    flat = np.concatenate([a.astype(dt).reshape(-1) for a in arrs])

    def unravel(v):
        v = _raw(v)
        out, o = [], 0
        for s, a in zip(shapes, arrs):
            n = int(np.prod(s)) if s else 1
            out.append(wrap(v[o:o + n].reshape(s).astype(a.dtype)))
            o += n
        return tree_unflatten(td, out)
    return wrap(flat), unravel
