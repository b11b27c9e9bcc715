"""Shim of gym.spaces (TEST HARNESS ONLY) -- shapes/dtypes/bounds and `contains`."""
from collections import OrderedDict

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = np.random.RandomState()

    @property
    def shape(self):
        return self._shape

    def seed(self, seed=None):
        self._np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape).astype(dtype)
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), shape).astype(dtype)
        super().__init__(shape, dtype)

    def sample(self):
        return self._np_random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        super().__init__((), np.int64)

    def sample(self):
        return int(self._np_random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype)

    def sample(self):
        return (self._np_random.random_sample(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))


class MultiBinary(Space):
    def __init__(self, n):
        self.n = n
        super().__init__((n,) if np.isscalar(n) else tuple(n), np.int8)

    def sample(self):
        return self._np_random.randint(0, 2, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        if spaces is None:
            spaces = kw
        self.spaces = OrderedDict(spaces)
        super().__init__(None, None)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)
        super().__init__(None, None)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(p) for s, p in zip(self.spaces, x))
