#!/usr/bin/env python3
"""Record golden per-step trajectories from the UNMODIFIED reference envs on real PyBullet.

Cannot run in the build image (pybullet, gym, simple_pid, matplotlib are absent); run it wherever those exist:

    python tools/record_golden_pybullet.py --reference /path/to/tennisbot-rl --out tests/golden/pybullet_traj.npz

For each env kind it forces fixed placements (by seeding Python's `random` / `np.random` right before reset() and
reading the placement back from the simulator), replays a fixed action sequence and stores, per step, the canonical
32-word state record of include/tennisbot_b200.h (racket COM pose / velocities, ball pose / velocities, aux, goal,
d0, step count), the observation, reward, done flag and the three contact predicates the envs query.  The parity
harness (tests/test_golden_pybullet.py) feeds the same placements / actions to the oracle and to the CUDA path.
"""
import argparse
import random
import sys

import numpy as np


def racket_state(p, env):
    pos, quat = p.getBasePositionAndOrientation(env.racket.id, env.client)
    vel, ang = p.getBaseVelocity(env.racket.id, env.client)
    return list(pos) + list(quat) + list(vel) + list(ang)


def ball_state(p, env):
    pos, _ = p.getBasePositionAndOrientation(env.ball.id, env.client)
    vel, ang = p.getBaseVelocity(env.ball.id, env.client)
    return list(pos) + list(vel) + list(ang)


def record(env_id, episodes, seed, max_steps):
    import gym
    import pybullet as p
    import tennisbot  # noqa: F401  (registers the ids)

    kw = dict(use_gui=False, delay_mode=False) if env_id == "SwingRacket-v0" else dict(use_gui=False)
    if env_id == "Tennisbot-v0":
        import tennisbot.envs.tennisbot_env as te

        te.DELAY_MODE = False  # module constant: the 1/240 s sleep per step
    env = gym.make(env_id, **kw).unwrapped
    rng = np.random.default_rng(seed)
    out = dict(state=[], obs=[], reward=[], done=[], contact=[], action=[], episode=[], params=[])
    for ep in range(episodes):
        random.seed(seed * 1000 + ep)
        np.random.seed(seed * 1000 + ep)
        env.reset()
        for t in range(max_steps):
            a = rng.uniform(-1, 1, env.action_space.shape[0]).astype(np.float32)
            ob, r, done, _ = env.step(a)
            rec = np.zeros(32)
            rec[0:13] = racket_state(p, env)
            rec[13:22] = ball_state(p, env)
            if env_id == "SwingRacket-v0":
                rec[22:25] = env.spawn_pos
                rec[25:27] = env.goal
                rec[27] = env.initial_dist_to_goal
            else:
                rec[22:25] = env.ball_shoot_force
            rec[29] = env.step_count
            rec[30] = float(done)
            rec[31] = ep
            contact = [len(p.getContactPoints(env.racket.id, env.ball.id)) > 0,
                       len(p.getContactPoints(env.court.id, env.ball.id)) > 0,
                       env_id == "SwingRacket-v0" and len(p.getContactPoints(env.goal_obj.id, env.ball.id)) > 0]
            out["state"].append(rec); out["obs"].append(np.asarray(ob, np.float64)); out["reward"].append(float(r))
            out["done"].append(bool(done)); out["contact"].append(contact); out["action"].append(a); out["episode"].append(ep)
            if done:
                break
    out["params"] = [str(p.getPhysicsEngineParameters())]
    env.close()
    return {k: np.asarray(v) for k, v in out.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="path to a tennisbot-rl checkout")
    ap.add_argument("--out", default="tests/golden/pybullet_traj.npz")
    ap.add_argument("--episodes", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    data = {}
    for env_id, steps in (("SwingRacket-v0", 26), ("Tennisbot-v0", 1001)):
        for k, v in record(env_id, args.episodes, args.seed, steps).items():
            data[f"{env_id}/{k}"] = v
    np.savez_compressed(args.out, **data)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
