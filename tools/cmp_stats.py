import sys, torch
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
n = 1 << 20
b = TennisBatch("SwingRacket-v0", n, precision=sys.argv[1], seed=0)
b.reset()
g = torch.Generator("cuda").manual_seed(123)
acts = [torch.empty((n, 6), device="cuda").uniform_(-1, 1, generator=g) for _ in range(4)]
for t in range(52):
    b.step(acts[t % 4])
torch.cuda.synchronize()
print(sys.argv[1], b.read_stats().tolist())
st = b.get_state().cpu()
print("state checksum", float(st.double().abs().sum()))
