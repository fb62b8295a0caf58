import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import binding as ob
from tennisbot_rl_b200.batch import TennisBatch
n = 4096
for policy in ("random", "zero"):
    b = TennisBatch("SwingRacket-v0", n, seed=19, precision="f64", auto_reset=False); o = ob.OracleEnv("SwingRacket-v0", n, seed=19, threads=16, auto_reset=False)
    b.set_param("racket_court_contact", 1.0); o.set_param("racket_court_contact", 1.0)
    b.reset(); o.reset()
    rng = np.random.default_rng(6)
    for t in range(26):
        a = (rng.uniform(-1, 1, (n, 6)) if policy == "random" else np.zeros((n, 6))).astype(np.float32)
        g = b.step(torch.from_numpy(a).cuda()); r = o.step(a)
    gs = b.get_state().cpu().numpy(); os_ = o.get_state()
    err = np.abs(gs - os_)
    e_r = err[:, 0:13].max(1); e_b = err[:, 13:22].max(1)
    low = (r["events"] & 64) != 0
    print(policy, "low", low.mean(), "racket err quantiles", np.quantile(e_r, [0.5, 0.9, 0.99, 0.999, 1.0]), "ball err max", e_b.max(),
          "n(racket err>1e-9)", (e_r > 1e-9).sum(), "n(>1e-6)", (e_r > 1e-6).sum(), "steps equal", (gs[:, 29] == os_[:, 29]).all())
    bad = np.argsort(-e_r)[:3]
    for i in bad:
        print("  env", i, "err", e_r[i], "racket pos gpu", gs[i, 0:3], "cpu", os_[i, 0:3], "step", gs[i, 29], "vel", gs[i, 7:10])
