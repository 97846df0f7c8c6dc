"""gymnasium.spaces.Box when gymnasium is installed, else a minimal look-alike with the attributes the
reference's callers touch (shape, dtype, low, high, sample, contains; soccer_env.py:34,67)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box  # type: ignore
except Exception:  # gymnasium absent (this image)
    class Box:  # type: ignore
        def __init__(self, low, high, shape, dtype=np.float32, seed=None):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            bounded = np.isfinite(self.low) & np.isfinite(self.high)
            u = self._rng.uniform(lo, hi)
            n = self._rng.normal(size=self.shape)
            return np.where(bounded, u, n).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

        __contains__ = contains

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, other):
            return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                    and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

try:  # pragma: no cover
    from pettingzoo import ParallelEnv  # type: ignore
except Exception:
    class ParallelEnv:  # type: ignore
        """Stand-in base class when pettingzoo is not installed."""
        metadata: dict = {}
