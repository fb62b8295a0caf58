#!/usr/bin/env python3
"""Record golden per-step trajectories from the UNMODIFIED reference envs on real PyBullet, in the harness's format
(tennisbot_rl_b200/trajectory.py).

Cannot run in the build image (pybullet, gym, simple_pid, matplotlib are absent); run it wherever those exist:

    python tools/record_golden_pybullet.py --reference /path/to/tennisbot-rl --out tests/golden/pybullet_traj.npz

Per env id it plays `--episodes` episodes with a fixed action tape.  For every episode it stores the state RIGHT AFTER reset()
(before step 1: racket COM pose, ball pose, spawn position / goal / d0 for SwingRacket-v0, shoot force and the ball's
re-placed position for Tennisbot-v0 - tennisbot_env.py:227-246), the placement record tb_reset_from needs, and per step the
canonical state, observation, reward, done and the three contact predicates, plus getPhysicsEngineParameters() as structured
fields.  tests/test_golden_pybullet.py then drives the oracle and the CUDA path from the file: that is what turns "parity
unpinned" into pinned, and where every recalled Bullet constant (tb_set_param names) gets corrected.
"""
import argparse
import random
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tennisbot_rl_b200 import trajectory as tj  # noqa: E402


def canonical_state(p, env, env_id, done):
    rec = np.zeros(tj.STATE_WORDS)
    pos, quat = p.getBasePositionAndOrientation(env.racket.id, env.client)  # COM frame (racket.py:131)
    vel, ang = p.getBaseVelocity(env.racket.id, env.client)
    rec[0:3], rec[3:7], rec[7:10], rec[10:13] = pos, quat, vel, ang
    bpos, _ = p.getBasePositionAndOrientation(env.ball.id, env.client)
    bvel, bang = p.getBaseVelocity(env.ball.id, env.client)
    rec[13:16], rec[16:19], rec[19:22] = bpos, bvel, bang
    if env_id == "SwingRacket-v0":
        rec[22:25] = env.spawn_pos
        rec[25:27] = env.goal
        rec[27] = env.initial_dist_to_goal
    else:
        rec[22:25] = env.ball_shoot_force
    rec[29] = env.step_count
    rec[30] = float(done)
    return rec


def contacts(p, env, env_id):
    return [len(p.getContactPoints(env.racket.id, env.ball.id)) > 0, len(p.getContactPoints(env.court.id, env.ball.id)) > 0,
            env_id == "SwingRacket-v0" and len(p.getContactPoints(env.goal_obj.id, env.ball.id)) > 0]


def record(env_id, episodes, seed):
    import gym
    import pybullet as p
    import tennisbot  # noqa: F401  (the reference's package: registers the ids)

    kw = dict(use_gui=False, delay_mode=False) if env_id == "SwingRacket-v0" else dict(use_gui=False)
    if env_id == "Tennisbot-v0":
        import tennisbot.envs.tennisbot_env as te

        te.DELAY_MODE = False  # module constant: the 1/240 s sleep per step (tennisbot_env.py:20,124-126)
    env = gym.make(env_id, **kw).unwrapped
    rng = np.random.default_rng(seed)
    w = tj.EpisodeWriter(env_id)
    for ep in range(episodes):
        random.seed(seed * 1000 + ep)     # the reference draws placements from the GLOBAL generators (swingracket_env.py:161-173)
        np.random.seed(seed * 1000 + ep)
        env.reset()
        s0 = canonical_state(p, env, env_id, False)
        w.begin(s0, tj.init_from_reset_state(env_id, s0, getattr(env, "racket_scale", 1.0)))
        for t in range(tj.MAX_STEPS[env_id]):
            a = rng.uniform(-1, 1, tj.ACT_DIM[env_id]).astype(np.float32)
            if env_id == "Tennisbot-v0" and ep % 2:  # every second episode steers at the ball: contacts with the racket
                rp, bp, rv = s_prev[1] if t else s0[1], s_prev[14] if t else s0[14], s_prev[8] if t else 0.0
                a = np.array([0.2 * a[0], np.clip(4.0 * (bp - rp) - 1.5 * rv, -1, 1)], np.float32)
            ob, r, done, _ = env.step(a)
            s_prev = canonical_state(p, env, env_id, done)
            w.step(a, s_prev, np.asarray(ob, np.float64), r, done, contacts(p, env, env_id))
            if done:
                break
    params = {k: (float(v) if isinstance(v, (int, float)) else str(v)) for k, v in p.getPhysicsEngineParameters().items()}
    try:
        import pkg_resources

        version = pkg_resources.get_distribution("pybullet").version
    except Exception:
        version = "unknown"
    env.close()
    return w, {"engine": f"pybullet {version}", "engine_params": params}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="path to a tennisbot-rl checkout")
    ap.add_argument("--out", default="tests/golden/pybullet_traj.npz")
    ap.add_argument("--episodes", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    writers, meta = [], {"producer": "tools/record_golden_pybullet.py", "racket_scale": 1.0}
    for env_id in ("SwingRacket-v0", "Tennisbot-v0"):
        w, m = record(env_id, args.episodes, args.seed)
        writers.append(w)
        meta.update(m)
    tj.save(args.out, writers, meta)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
