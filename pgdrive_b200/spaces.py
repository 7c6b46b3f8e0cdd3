"""Minimal Box space (gym is not a dependency; the reference carries its own copy of gym's spaces in
/root/reference/pgdrive/utils/space.py:309-470)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rs = np.random.RandomState()

    def seed(self, seed=None):
        self._rs = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return self._rs.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.allclose(self.low, other.low) and \
            np.allclose(self.high, other.high)


class MultiDiscrete:
    """gym.spaces.MultiDiscrete([n0, n1]) stand-in for ``discrete_action=True`` (base_vehicle.py:721-727)."""
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(np.int64)
        self._rs = np.random.RandomState()

    def seed(self, seed=None):
        self._rs = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return (self._rs.random_sample(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))

    __contains__ = contains

    def __repr__(self):
        return "MultiDiscrete(%s)" % (self.nvec.tolist(), )
