"""Per-launch timing of the step kernel over lock-step SwingRacket episodes (CUDA events on the launch stream).
usage: time_steps.py precision n_envs [episodes]"""
import sys
import numpy as np
import os
import torch
RING = int(os.environ.get("TB_RING", 32))  # pre-drawn action batches the steps cycle through (>= 26: i.i.d. within an episode)
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
prec = sys.argv[1]; n = int(sys.argv[2]); eps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
env = sys.argv[4] if len(sys.argv) > 4 else "SwingRacket-v0"
b = TennisBatch(env, n, precision=prec, seed=0)
for kv in os.environ.get("TB_PARAMS", "").split(","):  # e.g. TB_PARAMS=racket_court_contact=1
    if "=" in kv:
        b.set_param(kv.split("=")[0], float(kv.split("=")[1]))
b.reset()
acts = [torch.empty((n, b.act_dim), device="cuda").uniform_(-1, 1) for _ in range(RING)]
for t in range(26): b.step(acts[t % RING])
torch.cuda.synchronize()
b.read_stats(clear=True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(26 * eps + 1)]
ev[0].record()
for t in range(26 * eps):
    b.step(acts[t % RING]); ev[t + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(26 * eps)]).reshape(eps, 26)
st = b.read_stats()
light = ms[:, :25].mean(); heavy = ms[:, 25].mean(); total = ms.sum(1).mean()
print("%s n=%d light %.4f ms heavy %.3f ms episode %.3f ms -> %.3e env-steps/s, %.3e substeps/s, heavy-only %.3e substeps/s" % (
    prec, n, light, heavy, total, n * 26 / total * 1e3, st[8] / (ms.sum() * 1e-3), (st[8] - 25 * eps * n) / (ms[:, 25].sum() * 1e-3)))
