"""Minimal gym.spaces stand-ins (Box / MultiDiscrete / Dict) used when `gym` is not installed.
The reference builds these in Env_Train_GenNBV.update_observation_space (env_train_gennbv.py:459-492) and
flattens them in env_wrapper_gennbv_train.py:59-87; only shape / bounds / dtype / nvec are consumed downstream."""
from collections import OrderedDict

import numpy as np

try:                                               # pragma: no cover - gym is absent in the build image
    from gym.spaces import Box, Dict, MultiDiscrete  # type: ignore
except Exception:
    class Space:
        def __init__(self, shape=None, dtype=None):
            self._shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)

        @property
        def shape(self):
            return self._shape

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self._shape).astype(np.float32)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self._shape).astype(np.float32)

        def __repr__(self):
            return f"Box({self.shape}, {self.dtype})"

    class MultiDiscrete(Space):
        def __init__(self, nvec, dtype=np.int64):
            self.nvec = np.asarray(nvec, dtype=dtype)
            super().__init__(self.nvec.shape, dtype)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Dict(Space):
        def __init__(self, spaces):
            super().__init__(None, None)
            self.spaces = OrderedDict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()
