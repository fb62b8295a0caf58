"""Distributional known-answer test: the saved PPO policy of the reference (backup_models/ppo_swing.zip, weights
in tests/golden/ppo_swing_policy.npz) rolled out through the oracle must reproduce the return statistics the same
zip recorded on REAL PyBullet (tests/golden/ppo_swing_monitor.json): mean 31.53, sigma 24.57, 27 % goals, 26-step
episodes.  This is the only PyBullet-produced number available offline; it pins the damping law, the recomputed
racket inertia and (weakly) the contact threshold - see SURVEY.md Appendix C for what it cannot discriminate."""
import json
from pathlib import Path

import numpy as np

GOLD = Path(__file__).parent / "golden"


def policy_mean(w, obs):
    h = obs
    for layer in (0, 2, 4):
        h = np.tanh(h @ w[f"mlp_extractor__policy_net__{layer}__weight"].T + w[f"mlp_extractor__policy_net__{layer}__bias"])
    return h @ w["action_net__weight"].T + w["action_net__bias"]


def test_recorded_monitor_buffer_is_what_survey_says():
    mon = json.loads((GOLD / "ppo_swing_monitor.json").read_text())
    r = np.array(mon["episode_returns"])
    assert len(r) == 100 and abs(r.mean() - 31.5256) < 1e-3 and abs(r.std() - 24.5676) < 1e-3
    assert (r > 50).sum() == 27 and r.min() == 4.0 and abs(r.max() - 88.562688) < 1e-6
    assert sorted(set(mon["episode_lengths"])) == [26, 38]
    assert mon["hyper"]["n_steps"] == 1100 and mon["hyper"]["batch_size"] == 1100


def test_ppo_policy_return_distribution(oracle_lib):
    w = np.load(GOLD / "ppo_swing_policy.npz")
    mon = json.loads((GOLD / "ppo_swing_monitor.json").read_text())
    rec = np.array(mon["episode_returns"])
    n = 3000
    env = oracle_lib.OracleEnv("SwingRacket-v0", n, seed=1, threads=8, auto_reset=False)
    obs = env.reset()
    rng = np.random.default_rng(0)
    std = np.exp(w["log_std"])
    ret = np.zeros(n)
    hit_steps = np.zeros(n, np.int64)
    for k in range(26):
        a = np.clip(policy_mean(w, obs) + std * rng.standard_normal((n, 6)), -1, 1).astype(np.float32)
        out = env.step(a)
        obs = out["obs"]
        ret += out["reward"]
        hit_steps += ((out["events"] & 1) != 0) & (k < 24)  # rewarded racket-ball contact steps (+2 each, swingracket_env.py:98-102)
        assert bool(out["done"].all()) == (k == 25)  # every episode is exactly 26 agent steps
    goal = (ret > 50).mean()
    print("oracle: mean %.2f std %.2f max %.2f goal %.3f | recorded: mean %.2f std %.2f max %.2f goal %.2f"
          % (ret.mean(), ret.std(), ret.max(), goal, rec.mean(), rec.std(), rec.max(), (rec > 50).mean()))
    # acceptance window of SURVEY Appendix C (recorded sample is only 100 episodes: sigma_mean ~ 2.5)
    assert 27 <= ret.mean() <= 35
    assert 21 <= ret.std() <= 27
    assert 0.20 <= goal <= 0.32
    assert ret.max() >= 85            # court + goal contact in the same step occurs occasionally (recorded max 88.56)
    # goal returns sit at 50 + moved() + 2 per contact step, like the recorded 70.0 .. 73.1 cluster
    g = ret[(ret > 50) & (ret < 80)]
    assert 69 < g.min() and g.max() < 75
    # --- the whole distribution against the 98 regular episodes of the recording (two are the 38-step EvalCallback artefact,
    # SURVEY 3.1): two-sample Kolmogorov-Smirnov distance (critical value at alpha = 0.05 for 98 vs 3000 samples: 0.14;
    # measured 0.084), quantiles of the court-landing body, and the mix of one / two rewarded contact steps
    rec26 = np.sort(rec[np.array(mon["episode_lengths"]) == 26])
    xs = np.sort(np.concatenate([ret, rec26]))
    ks = np.abs(np.searchsorted(np.sort(ret), xs, side="right") / n - np.searchsorted(rec26, xs, side="right") / len(rec26)).max()
    print("KS distance %.3f" % ks, "hit-step mix", np.bincount(hit_steps)[:4] / n)
    assert ks < 0.14
    for q, tol in ((0.1, 1.0), (0.25, 1.2), (0.5, 0.6)):
        assert abs(np.quantile(ret, q) - np.quantile(rec26, q)) < tol, q
    court, rcourt = ret[ret < 50], rec26[rec26 < 50]
    assert abs(court.mean() - rcourt.mean()) < 0.8 and abs(court.std() - rcourt.std()) < 0.8
    # goal returns = 50 + moved() + 2 per contact step: the recording has 24 in 70.0 .. 71.3 (one contact step) and 3 in
    # 71.9 .. 73.1 (two); a single +2 is the rule, two happen in a few per cent of episodes
    two = (hit_steps == 2).mean()
    assert (hit_steps == 1).mean() > 0.8 and 0.03 < two < 0.2 and (hit_steps > 2).mean() < 0.01
    # ~100 % of trained-policy episodes hit the ball; most end by landing on the court
    st = env.read_stats()
    assert st[2] >= 0.95 * n and st[5] < 0.02 * n
    assert 300 < st[8] / n < 420      # physics steps per episode (SURVEY: 361)
