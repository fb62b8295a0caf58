"""Small driver for ncu: lock-step SwingRacket batch, 27 step launches (the 26th is the fast-forward step)."""
import sys
import torch
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
env = sys.argv[3] if len(sys.argv) > 3 else "SwingRacket-v0"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 27
b = TennisBatch(env, n, precision=prec, seed=0)
b.reset()
acts = [torch.empty((n, b.act_dim), device="cuda").uniform_(-1, 1) for _ in range(4)]  # the ring bench.py steps through
for t in range(steps):
    b.step(acts[t % 4])
torch.cuda.synchronize()
print(b.read_stats())
