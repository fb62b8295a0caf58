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
    for k in range(26):
        a = np.clip(policy_mean(w, obs) + std * rng.standard_normal((n, 6)), -1, 1).astype(np.float32)
        out = env.step(a)
        obs = out["obs"]
        ret += out["reward"]
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
    # ~100 % of trained-policy episodes hit the ball; most end by landing on the court
    st = env.read_stats()
    assert st[2] >= 0.95 * n and st[5] < 0.02 * n
    assert 300 < st[8] / n < 420      # physics steps per episode (SURVEY: 361)
