import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from tennisbot_rl_b200.batch import TennisBatch
prec = sys.argv[1]; n = int(sys.argv[2]); env = sys.argv[3] if len(sys.argv) > 3 else "SwingRacket-v0"
b = TennisBatch(env, n, precision=prec, seed=0); b.reset_host()
hb = b.host_buffers(); hb["actions"][...] = np.random.default_rng(0).uniform(-1, 1, hb["actions"].shape)
for _ in range(26): b.step_host(want_terminal=False, want_events=False)
t = []
for k in range(52):
    t0 = time.perf_counter(); b.step_host(want_terminal=False, want_events=False); t.append(time.perf_counter() - t0)
t = np.array(t).reshape(2, 26) * 1e3
print("%s %s n=%d e2e light %.3f ms heavy %.3f ms episode %.2f ms -> %.3e env-steps/s" % (env, prec, n, t[:, :25].mean(), t[:, 25].mean(), t.sum(1).mean(), n * 26 / t.sum(1).mean() * 1e3))
