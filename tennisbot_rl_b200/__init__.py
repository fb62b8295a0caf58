"""tennisbot_rl_b200 - B200-native env step for youliangtan/tennisbot-rl's SwingRacket-v0 and Tennisbot-v0.

The per-step dynamics the reference delegates to PyBullet run as hand-written sm_100a CUDA kernels behind a C ABI
(include/tennisbot_b200.h); this package is the thin Python host side: ctypes binding, device-resident batch,
SB3-style VecEnv adapter, single-env gym classes and the multi-GPU shard helper.
"""
from ._lib import TennisbotLibraryError, param_names, scene_constant  # noqa: F401

__all__ = ["TennisBatch", "TennisVecEnv", "SwingRacketEnv", "TennisbotEnv", "make", "TennisbotLibraryError"]


def __getattr__(name):  # lazy: importing the package must not need torch / a GPU
    if name == "TennisBatch":
        from .batch import TennisBatch
        return TennisBatch
    if name == "TennisVecEnv":
        from .vec_env import TennisVecEnv
        return TennisVecEnv
    if name in ("SwingRacketEnv", "TennisbotEnv", "make"):
        from . import envs
        return getattr(envs, name)
    raise AttributeError(name)
