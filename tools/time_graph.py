"""Env steps issued eagerly vs replayed from a CUDA graph (launch overhead matters at small batch sizes).
usage: time_graph.py env precision n_envs [steps_per_graph]"""
import sys
import torch
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
env, prec, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
k = int(sys.argv[4]) if len(sys.argv) > 4 else 26
assert k % 2 == 0, "capture an even number of steps (INTEGRATION.md)"
b = TennisBatch(env, n, precision=prec, seed=0)
b.reset()
acts = [torch.empty((n, b.act_dim), device="cuda").uniform_(-1, 1) for _ in range(4)]
def run():
    for t in range(k):
        b.step(acts[t % 4])
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        run()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = max(4, 2000 // k)
e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
eager = e0.elapsed_time(e1) / (reps * k)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
g.replay(); torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    g.replay()
e1.record(); torch.cuda.synchronize()
graph = e0.elapsed_time(e1) / (reps * k)
print("%s %s n=%d: eager %.4f ms/step (%.3e env-steps/s), CUDA graph of %d steps %.4f ms/step (%.3e env-steps/s)" % (
    env, prec, n, eager, n / eager * 1e3, k, graph, n / graph * 1e3))
