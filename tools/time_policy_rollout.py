"""Where the time of a 26-step policy rollout goes at a given batch size: per-iteration rollout time while PPO trains (the
workload changes as the policy learns to hit the ball: longer flights, contact chains), and the per-kernel split of one rollout.
usage: time_policy_rollout.py [n_envs] [iters]"""
import sys
import time
import torch
sys.path.insert(0, ".")
from tennisbot_rl_b200.ppo import SwingPPO, EPISODE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ppo = SwingPPO(num_envs=n, seed=0, use_graph=True, fused_policy=True)
for it in range(iters):
    t0 = ppo.rollout_s
    st = ppo.rollout()
    dt = ppo.rollout_s - t0
    if it % 10 == 0 or it == iters - 1:
        eps = max(int(st[0]), 1)
        print("iter %3d rollout %.3f ms -> %.3e env-steps/s; return %.2f, mean episode length %.0f substeps, hits/ep %.2f" % (
            it, dt * 1e3, EPISODE * n / dt, st[6] / 2 ** 20 / eps, st[1] / eps, st[2] / eps), flush=True)
    ppo.update()
# per-kernel split of one eager rollout with the trained policy (events inside the library, synchronised per step)
ppo.env.set_kernel_timing(True)
ppo.use_graph = False
ppo.rollout()
a, b, k = ppo.env.kernel_timing()
print("trained policy, eager, timed per step: step_kernel %.3f ms + ff_kernel %.3f ms over %d steps" % (a, b, k))
d = ppo.env.ff_diagnostics()
print("last fast-forward: server visits %d, flights over after %.3f ms" % (d[1], d[2] * 1e-6))
