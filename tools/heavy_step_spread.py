"""Spread of the fast-forward step's time over many episodes (one CUDA-event pair per 26th step).
usage: heavy_step_spread.py [episodes] [n_envs]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch

episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
b = TennisBatch("SwingRacket-v0", n, seed=0)
ring = [torch.empty((n, 6), device="cuda").uniform_(-1, 1) for _ in range(32)]
b.reset()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(episodes)]
k = 0
for e in range(episodes + 2):
    for s in range(26):
        if s == 25 and e >= 2:
            ev[e - 2][0].record()
        b.step(ring[k % 32]); k += 1
        if s == 25 and e >= 2:
            ev[e - 2][1].record()
torch.cuda.synchronize()
t = np.array([a.elapsed_time(c) for a, c in ev])
print("fast-forward step over %d episodes: median %.3f ms, mean %.3f, min %.3f, p90 %.3f, p99 %.3f, max %.3f; above 1.1 x median: %d"
      % (episodes, np.median(t), t.mean(), t.min(), np.quantile(t, .9), np.quantile(t, .99), t.max(), (t > 1.1 * np.median(t)).sum()))
print("slowest:", np.sort(t)[-8:].round(3), "at episodes", np.argsort(t)[-8:])
