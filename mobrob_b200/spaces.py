"""Minimal ``Box`` space with gymnasium 0.28.1's seeding/sampling semantics for bounded
float boxes (the only kind mobrob uses: src/mobrob/envs/wrapper.py:250-264, engine.py:314).
Used when gymnasium itself is not installed."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self._shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self._shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self._shape).copy()
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    def seed(self, seed=None):
        ss = np.random.SeedSequence(seed)
        self._np_random = np.random.Generator(np.random.PCG64(ss))
        return [ss.entropy]

    def sample(self):
        bounded = self.bounded_below & self.bounded_above
        sample = np.empty(self._shape)
        unb = ~self.bounded_below & ~self.bounded_above
        sample[unb] = self.np_random.normal(size=unb[unb].shape)
        low_only = self.bounded_below & ~self.bounded_above
        sample[low_only] = self.np_random.exponential(size=low_only[low_only].shape) + self.low[low_only]
        up_only = ~self.bounded_below & self.bounded_above
        sample[up_only] = -self.np_random.exponential(size=up_only[up_only].shape) + self.high[up_only]
        sample[bounded] = self.np_random.uniform(low=self.low[bounded], high=self.high[bounded],
                                                 size=bounded[bounded].shape)
        return sample.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self._shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self._shape}, {self.dtype})"
