"""Step-by-step parity against the oracle at the bench's own batch size (1 048 576 envs per GPU), a one-off that is too slow for
the test suite's budget: every event byte and done flag (each mismatch is listed with the oracle's margin to the nearest decision
threshold), the statistics, and the state / obs / reward errors over the envs that have never mismatched.
usage: full_size_parity.py [n_envs] [episodes (SwingRacket-v0) | steps (Tennisbot-v0)] [env]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import binding
from tests.harness import reference_reset_params
from tennisbot_rl_b200.batch import TennisBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
count = int(sys.argv[2]) if len(sys.argv) > 2 else 4
env = sys.argv[3] if len(sys.argv) > 3 else "SwingRacket-v0"
swing = env == "SwingRacket-v0"
steps = 26 * count if swing else count
every = 13 if swing else 100
binding.build()
t0 = time.time()
b = TennisBatch(env, n, seed=101, precision="f64")
o = binding.OracleEnv(env, n, seed=101, threads=16)
rng = np.random.default_rng(77)
init = reference_reset_params(o.kind, n, rng)
np.testing.assert_array_equal(b.reset(init=init).cpu().numpy(), o.reset(init=init))
valid = np.ones(n, bool)
e_obs, e_rew, e_state = np.zeros(n), np.zeros(n), np.zeros(n)  # per env, over the steps before its first mismatch
mismatches = []
for t in range(steps):
    a = rng.uniform(-1, 1, (n, o.act_dim)).astype(np.float32)
    g = [x.cpu().numpy() for x in b.step(torch.from_numpy(a).to(b.device))]
    r = o.step(a, want_margin=True)
    disc = (g[2] != r["done"]) | (g[4] != r["events"])
    for i in np.nonzero(disc)[0][:20]:
        mismatches.append(dict(step=t, env=int(i), events_gpu=int(g[4][i]), events_oracle=int(r["events"][i]), done_gpu=int(g[2][i]),
                               done_oracle=int(r["done"][i]), oracle_margin=float(r["margin"][i]), counted=bool(valid[i])))
    valid &= ~disc
    v = valid
    e_obs[v] = np.maximum(e_obs[v], np.abs(g[0][v].astype(np.float64) - r["obs"][v]).max(1))
    e_rew[v] = np.maximum(e_rew[v], np.abs(g[1][v].astype(np.float64) - r["reward"][v]))
    d = v & (r["done"] != 0)
    if d.any():
        e_obs[d] = np.maximum(e_obs[d], np.abs(g[3][d].astype(np.float64) - r["terminal_obs"][d]).max(1))
    if (t + 1) % every == 0:
        e_state[v] = np.maximum(e_state[v], np.abs(b.get_state().cpu().numpy()[v] - o.get_state()[v]).max(1))
sg, so = b.read_stats(), o.read_stats()
print(f"{env} {n} envs x {steps} steps = {n * steps} env steps, {int(so[0])} episodes, {time.time() - t0:.0f} s")
print(f"  event / done mismatches: {len(mismatches)}; envs still compared at the end: {int(valid.sum())}")
for m in mismatches:
    print("   ", m)
v = valid
print(f"  max errors over the {int(v.sum())} envs without a mismatch: obs {e_obs[v].max():.3e}, reward {e_rew[v].max():.3e}, state {e_state[v].max():.3e}")
print("  envs (all) whose largest state error before any mismatch exceeds 1e-9 / 1e-8 / 1e-7 / 1e-6 / 1e-5 / 1e-4:",
      [int((e_state > x).sum()) for x in (1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4)])
print(f"  statistics  gpu    {sg.tolist()}\n              oracle {so.tolist()}  {'equal' if (sg == so).all() else 'DIFFERENT'}")
