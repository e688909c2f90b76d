"""`gymnasium.spaces` stand-ins (Box, Dict) used when gymnasium is not installed (it is absent from this image).
Only what QuadrupedEnv's surface needs: shape / low / high / dtype, sample(), contains(), mapping access."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - depends on the image
    from gymnasium import Env  # type: ignore
    from gymnasium.spaces import Box, Dict  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium not installed
    HAVE_GYMNASIUM = False

    class Env:  # minimal gym.Env surface
        metadata: dict = {}
        action_space = None
        observation_space = None

        def step(self, action):
            raise NotImplementedError

        def reset(self, **kwargs):
            raise NotImplementedError

        def render(self, *a, **k):
            raise NotImplementedError

        def close(self):
            pass

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low) if np.ndim(low) else np.shape(high)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape).astype(self.dtype)
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            """gymnasium semantics: N(0,1) where unbounded, shifted exponential where half-bounded, uniform where bounded."""
            lo_b, hi_b = np.isfinite(self.low), np.isfinite(self.high)
            out = np.empty(self.shape, dtype=np.float64)
            both, none = lo_b & hi_b, ~lo_b & ~hi_b
            out[none] = self._rng.normal(size=int(none.sum()))
            out[both] = self._rng.uniform(self.low[both], self.high[both])
            lo_only, hi_only = lo_b & ~hi_b, ~lo_b & hi_b
            out[lo_only] = self.low[lo_only] + self._rng.exponential(size=int(lo_only.sum()))
            out[hi_only] = self.high[hi_only] - self._rng.exponential(size=int(hi_only.sum()))
            return out.astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f'Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})'

    class Dict:
        def __init__(self, spaces=None, **kw):
            self.spaces = OrderedDict(spaces or {})
            self.spaces.update(kw)

        def __getitem__(self, k):
            return self.spaces[k]

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def values(self):
            return self.spaces.values()

        def sample(self):
            return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

        def contains(self, x):
            return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def __repr__(self):
            return 'Dict(' + ', '.join(f'{k}: {v}' for k, v in self.spaces.items()) + ')'
