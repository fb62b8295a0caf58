import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import binding as ob
from tests.harness import reference_reset_params
from tennisbot_rl_b200.batch import TennisBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
b = TennisBatch("SwingRacket-v0", n, seed=101, precision="f64"); o = ob.OracleEnv("SwingRacket-v0", n, seed=101, threads=16)
rng = np.random.default_rng(77)
init = reference_reset_params(o.kind, n, rng)
b.reset(init=init); o.reset(init=init)
for t in range(78):
    a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    pre = o.get_state() if (t % 26) == 25 else None
    g = [x.cpu().numpy() for x in b.step(torch.from_numpy(a).cuda())]
    r = o.step(a)
    eo = np.abs(g[0] - r["obs"]).max(1); er = np.abs(g[1] - r["reward"]); et = np.abs(g[3] - r["terminal_obs"]).max(1) * (r["done"] != 0)
    bad = np.nonzero((eo > 1e-5) | (er > 1e-5) | (et > 1e-5))[0]
    if len(bad):
        print("step", t, "bad envs", len(bad), bad[:10])
        for i in bad[:4]:
            print(" env", i, "obs", g[0][i], r["obs"][i], "rew", g[1][i], r["reward"][i], "term", g[3][i], r["terminal_obs"][i], "ev", g[4][i], r["events"][i])
            if pre is not None: print("  oracle pre-step state: rp", pre[i, 0:3], "bp", pre[i, 13:16], "bv", pre[i, 16:19], "step", pre[i, 29])
print("done")
