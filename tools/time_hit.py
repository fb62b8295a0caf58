"""Mean step time of Tennisbot-v0 once episodes have desynchronised (CUDA events on the launch stream).
usage: time_hit.py precision n_envs [warm_steps] [timed_steps]"""
import sys
import torch
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
prec = sys.argv[1]; n = int(sys.argv[2])
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
timed = int(sys.argv[4]) if len(sys.argv) > 4 else 500
b = TennisBatch("Tennisbot-v0", n, precision=prec, seed=0)
b.reset()
acts = [torch.empty((n, b.act_dim), device="cuda").uniform_(-1, 1) for _ in range(4)]
for t in range(warm): b.step(acts[t % 4])
torch.cuda.synchronize()
b.read_stats(clear=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(timed): b.step(acts[t % 4])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / timed
st = b.read_stats()
print("%s n=%d %.4f ms/step -> %.3e env-steps/s; episodes %d mean length %.1f hits %d" % (prec, n, ms, n / ms * 1e3, st[0], st[1] / max(st[0], 1), st[2]))
