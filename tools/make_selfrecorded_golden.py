#!/usr/bin/env python3
"""SELF-RECORDED trajectories in the golden format: produced by the in-repo CPU oracle, NOT by PyBullet.

    python tools/make_selfrecorded_golden.py            # -> tests/golden/selfrecorded_traj.npz

They pin nothing about Bullet.  Their purpose is that the plumbing of tests/test_golden_pybullet.py - replay from the file's
reset state / placement record / action tape through the oracle and through the CUDA path, comparison of every recorded
quantity - runs in CI, so that the day a real recording (tools/record_golden_pybullet.py) is dropped next to it, the only thing
that can fail is the physics.  The Tennisbot-v0 episodes alternate random actions and the scripted tracker, as the recorder's do.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import binding as ob  # noqa: E402
from tennisbot_rl_b200 import trajectory as tj  # noqa: E402


def record(env_id, episodes, seed):
    rng = np.random.default_rng(seed)
    w = tj.EpisodeWriter(env_id)
    o = ob.OracleEnv(env_id, 1, seed=seed, auto_reset=False)
    for ep in range(episodes):
        o.reset()
        s = o.get_state()[0].copy()
        w.begin(s, tj.init_from_reset_state(env_id, s))
        for t in range(tj.MAX_STEPS[env_id]):
            a = rng.uniform(-1, 1, tj.ACT_DIM[env_id]).astype(np.float32)
            if env_id == "Tennisbot-v0" and ep % 2:
                a = np.array([0.2 * a[0], np.clip(4.0 * (s[14] - s[1]) - 1.5 * s[8], -1, 1)], np.float32)
            out = o.step(a[None], want_obs64=True)
            s = o.get_state()[0].copy()
            ev = int(out["events"][0])
            w.step(a, s, out["obs64"][0], float(out["reward"][0]), bool(out["done"][0]), [ev & 1, ev & 2, ev & 4])
            if out["done"][0]:
                break
    return w


def main():
    out = ROOT / "tests" / "golden" / "selfrecorded_traj.npz"
    writers = [record("SwingRacket-v0", 12, 3), record("Tennisbot-v0", 4, 4)]
    tj.save(out, writers, {"producer": "tools/make_selfrecorded_golden.py", "engine": "oracle (oracle/tb_oracle.c) - NOT PyBullet",
                           "engine_params": {}, "racket_scale": 1.0})
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
