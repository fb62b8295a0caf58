"""Observation / action spaces of the two envs, as the reference declares them.

`gym` (0.21, what the reference pins through SB3 1.8) and `gymnasium` are optional: when either is importable the
real `Box` is returned so SB3 accepts the env unchanged; otherwise a duck-typed Box with the attributes SB3, the
ES trainer (`fitness_functions.py:40-51,112-114`) and the TRPO agent read (`shape, low, high, dtype, sample,
contains`) stands in.
"""
import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __contains__(self, x):
        return self.contains(x)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return (hasattr(other, "low") and hasattr(other, "high") and np.array_equal(self.low, other.low)
                and np.array_equal(self.high, other.high))


def _box_cls():
    for mod in ("gym.spaces", "gymnasium.spaces"):
        try:
            return __import__(mod, fromlist=["Box"]).Box
        except Exception:
            continue
    return Box


def make_box(low, high):
    cls = _box_cls()
    low = np.asarray(low, dtype=np.float32)
    high = np.asarray(high, dtype=np.float32)
    if cls is Box:
        return Box(low, high)
    return cls(low=low, high=high, dtype=np.float32)


def swing_spaces():
    """swingracket_env.py:29-39"""
    action = make_box([-1, -1, -1, -1, -1, -1], [1, 1, 1, 1, 1, 1])
    obs = make_box([-20, -10, -20, -10, -15, -5], [20, 10, 20, 10, 0, 5])
    return obs, action


def hit_spaces():
    """tennisbot_env.py:37-55"""
    action = make_box([-1.0, -1.0], [1.0, 1.0])
    obs = make_box([-20, -20, -5, -5, -5, -5] + [-20, -20, 0, -10, -10, -10],
                   [20, 20, 5, 5, 5, 5] + [20, 20, 10, 10, 10, 10])
    return obs, action


def spaces_for(env_id):
    if env_id in ("SwingRacket-v0", "swing", 0):
        return swing_spaces()
    if env_id in ("Tennisbot-v0", "hit", 1):
        return hit_spaces()
    raise KeyError(env_id)
