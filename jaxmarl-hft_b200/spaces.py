"""Minimal action / observation space types: what the trainer reads is ``.n`` and ``.shape``
(gymnax_exchange/jaxen/from_JAXMARL/spaces.py:19-73, ippo_rnn_JAXMARL.py:525-529)."""
import numpy as np


class Discrete:
    def __init__(self, num_categories: int, dtype=np.int32):
        assert num_categories >= 0
        self.n = int(num_categories)
        self.shape = ()
        self.dtype = dtype

    def sample(self, rng: np.random.Generator):
        return rng.integers(0, self.n, dtype=self.dtype)

    def contains(self, x) -> bool:
        return bool(0 <= int(x) < self.n)


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self, rng: np.random.Generator):
        return rng.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))


class MultiDiscrete:
    """from_JAXMARL/spaces.py:45: a vector of categorical variables (the EXE fixed_prices action space)."""

    def __init__(self, num_categories, dtype=np.int32):
        self.n = list(num_categories)
        self.num_categories = np.asarray(num_categories, dtype=np.int64)
        self.shape = (len(self.num_categories),)
        self.dtype = dtype

    def sample(self, rng: np.random.Generator):
        return (rng.random(self.shape) * self.num_categories).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= 0) and np.all(x < self.num_categories))
