"""chex stand-in: only the type aliases / no-op assertions the reference touches."""
import dataclasses as _dc

import jax as _jax

Array = _jax.Array
ArrayTree = object
PRNGKey = _jax.Array
Scalar = object
Numeric = object
Shape = tuple


def dataclass(cls=None, **kw):
    from flax import struct
    return struct.dataclass(cls) if cls is not None else (lambda c: struct.dataclass(c))


def assert_gpu_available(*a, **k):
    pass


def assert_shape(*a, **k):
    pass


def assert_equal_shape(*a, **k):
    pass
