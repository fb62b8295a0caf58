import sys, torch
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
prec = sys.argv[1]; n = int(sys.argv[2])
b = TennisBatch("SwingRacket-v0", n, precision=prec, seed=0); b.reset()
a = torch.empty((n, 6), device="cuda").uniform_(-1, 1)
for t in range(25): b.step(a)
b.read_stats(clear=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b.step(a); e1.record(); torch.cuda.synchronize()
st = b.read_stats()
print("n=%d heavy %.3f ms thread-substeps %d lane-slots %d -> lane utilisation %.3f, warp-iterations per env-group %.1f" % (
    n, e0.elapsed_time(e1), st[8], st[9], st[8] / st[9], st[9] / 32 / (n / 32)))
