"""Small driver for ncu: lock-step SwingRacket batch, 27 step launches (the 26th is the fast-forward step)."""
import sys
import os
import torch
RING = int(os.environ.get("TB_RING", 32))  # pre-drawn action batches the steps cycle through (>= 26: i.i.d. within an episode)
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
env = sys.argv[3] if len(sys.argv) > 3 else "SwingRacket-v0"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 27
b = TennisBatch(env, n, precision=prec, seed=0)
b.reset()
acts = [torch.empty((n, b.act_dim), device="cuda").uniform_(-1, 1) for _ in range(RING)]  # the ring bench.py steps through
for t in range(steps):
    b.step(acts[t % RING])
torch.cuda.synchronize()
print(b.read_stats())
