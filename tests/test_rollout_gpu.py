"""GPU tests of the in-kernel action sources: the policy rollout (tb_set_policy / tb_policy_rollout) against the
numpy policy on tests/golden/ppo_swing_policy.npz with the oracle as the env, and the scripted tracker of tb_rollout."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden" / "ppo_swing_policy.npz"


def _golden_policy():
    """The reference's saved policy (backup_models/ppo_swing.zip -> policy.pth, extracted by tools/extract_fixtures.py)."""
    w = np.load(GOLDEN)
    lin = lambda net: [(w[f"mlp_extractor__{net}__{l}__weight"], w[f"mlp_extractor__{net}__{l}__bias"]) for l in (0, 2, 4)]  # noqa: E731
    return dict(pi=lin("policy_net"), vf=lin("value_net"), mu=(w["action_net__weight"], w["action_net__bias"]),
                v=(w["value_net__weight"], w["value_net__bias"]), log_std=w["log_std"])


def _forward(p, obs):
    """float32 numpy forward of the SB3 MlpPolicy: mean action and value."""
    def tower(layers, x):
        for W, b in layers:
            x = np.tanh(x @ W.T.astype(np.float32) + b.astype(np.float32))
        return x
    obs = obs.astype(np.float32)
    mean = tower(p["pi"], obs) @ p["mu"][0].T.astype(np.float32) + p["mu"][1].astype(np.float32)
    value = (tower(p["vf"], obs) @ p["v"][0].T.astype(np.float32) + p["v"][1].astype(np.float32))[:, 0]
    return mean, value


def _pack(p):
    from tennisbot_rl_b200.ppo import pack_policy

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))  # noqa: E731
    tw = lambda layers: [(t(W), t(b)) for W, b in layers]  # noqa: E731
    return pack_policy(tw(p["pi"]), (t(p["mu"][0]), t(p["mu"][1])), tw(p["vf"]), (t(p["v"][0]), t(p["v"][1])), t(p["log_std"]))


@pytest.mark.parametrize("deterministic", [False, True])
def test_policy_rollout_matches_numpy_policy_and_oracle(oracle_lib, deterministic):
    """One whole episode (K = 26) per env from tb_policy_rollout: (1) the recorded mean / value / log-density agree with the
    numpy forward of the same parameters on the recorded observations, the implied noise is N(0, 1); (2) the oracle driven by
    the recorded (clipped) actions reproduces observations, rewards and done flags: the trajectory is the env's."""
    from tennisbot_rl_b200 import _lib
    from tennisbot_rl_b200.batch import TennisBatch

    n, K = 4096, 26
    p = _golden_policy()
    b = TennisBatch("SwingRacket-v0", n, seed=21, precision="f64")
    o = oracle_lib.OracleEnv("SwingRacket-v0", n, seed=21, threads=8)
    b.set_policy(_pack(p).cuda())
    dev = b.device
    obs = torch.zeros((K, n, 6), device=dev)
    act, logp, val = torch.zeros((K, n, 6), device=dev), torch.zeros((K, n), device=dev), torch.zeros((K, n), device=dev)
    rew, done = torch.zeros((K, n), device=dev), torch.zeros((K, n), dtype=torch.uint8, device=dev)
    last_obs, last_val = torch.zeros((n, 6), device=dev), torch.zeros(n, device=dev)
    obs[0].copy_(b.reset())
    np.testing.assert_array_equal(obs[0].cpu().numpy(), o.reset())
    b.policy_rollout(obs, rew, done, last_obs, actions=act, logp=logp, value=val, last_value=last_val,
                     deterministic=deterministic, noise_seed=5)
    torch.cuda.synchronize()
    obs_h, act_h, logp_h, val_h, rew_h, done_h = (x.cpu().numpy() for x in (obs, act, logp, val, rew, done))
    std = np.exp(p["log_std"]).astype(np.float32)
    eps_all = []
    for t in range(K):
        mean, value = _forward(p, obs_h[t])
        np.testing.assert_allclose(val_h[t], value, atol=2e-4, rtol=1e-4)
        eps = (act_h[t] - mean) / std
        if deterministic:
            assert np.abs(eps).max() < 1e-4
        eps_all.append(eps)
        ref_logp = (-0.5 * eps.astype(np.float64) ** 2 - p["log_std"] - 0.5 * np.log(2 * np.pi)).sum(1)
        np.testing.assert_allclose(logp_h[t], ref_logp, atol=2e-3)
        # the env side: the oracle stepped with the clipped action SB3 would pass to env.step
        ref = o.step(np.clip(act_h[t], -1, 1).astype(np.float32))
        nxt = obs_h[t + 1] if t + 1 < K else last_obs.cpu().numpy()
        np.testing.assert_array_equal(done_h[t], ref["done"])
        np.testing.assert_allclose(nxt, ref["obs"], atol=2e-6)
        np.testing.assert_allclose(rew_h[t], ref["reward"], atol=2e-6)
    assert done_h[:25].sum() == 0 and done_h[25].all()
    np.testing.assert_allclose(last_val.cpu().numpy(), _forward(p, last_obs.cpu().numpy())[1], atol=2e-4, rtol=1e-4)
    np.testing.assert_array_equal(b.read_stats(), o.read_stats())
    if not deterministic:
        e = np.concatenate(eps_all).ravel()
        assert abs(e.mean()) < 0.01 and abs(e.std() - 1) < 0.01 and abs((e ** 3).mean()) < 0.03
        # independent across components and steps: no pair of columns is correlated
        c = np.corrcoef(np.concatenate(eps_all).T)
        assert np.abs(c - np.eye(6)).max() < 0.02
        st = b.read_stats()
        assert st[2] > 0.5 * st[0]  # the trained policy hits the ball in most episodes (racket-ball contact steps rewarded)
    b.close()


def test_policy_rollout_noise_advances_under_graph_replay():
    """The rollout captured once in a CUDA graph draws fresh noise on every replay (the noise counter is a device word)."""
    from tennisbot_rl_b200.batch import TennisBatch

    n, K = 2048, 26
    b = TennisBatch("SwingRacket-v0", n, seed=2)
    b.set_policy(_pack(_golden_policy()).cuda())
    dev = b.device
    obs, act = torch.zeros((K, n, 6), device=dev), torch.zeros((K, n, 6), device=dev)
    rew, done, last = torch.zeros((K, n), device=dev), torch.zeros((K, n), dtype=torch.uint8, device=dev), torch.zeros((n, 6), device=dev)
    obs[0].copy_(b.reset())
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        b.policy_rollout(obs, rew, done, last, actions=act, noise_seed=1)
    torch.cuda.current_stream(dev).wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        obs[0].copy_(last)
        b.policy_rollout(obs, rew, done, last, actions=act, noise_seed=1)
    runs = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        runs.append((act.clone(), rew[25].clone()))
    assert not torch.equal(runs[0][0], runs[1][0]) and not torch.equal(runs[1][0], runs[2][0])
    assert b.read_stats()[0] == 4 * n  # four whole episodes per env: the warm-up and three replays
    b.close()


def test_tracker_rollout_matches_oracle(oracle_lib):
    """tb_rollout(TB_ACT_TRACK) on Tennisbot-v0 == the oracle's rollout with the same action law and Philox streams; the
    scripted tracker produces racket-ball contacts (BASELINE config 3)."""
    from tennisbot_rl_b200 import _lib
    from tennisbot_rl_b200.batch import TennisBatch

    n, K = 4096, 900
    b = TennisBatch("Tennisbot-v0", n, seed=13)
    o = oracle_lib.OracleEnv("Tennisbot-v0", n, seed=13, threads=8)
    np.testing.assert_array_equal(b.reset().cpu().numpy(), o.reset())
    gobs, grs, gdc = (x.cpu().numpy() for x in b.rollout(K, action_mode=_lib.ACT_TRACK))
    ref = o.rollout(K, action_mode=1)
    np.testing.assert_array_equal(gdc, ref["done_count"])
    np.testing.assert_allclose(gobs, ref["obs"], atol=2e-5)
    np.testing.assert_allclose(grs, ref["reward_sum"], atol=1e-4)
    st = b.read_stats()
    np.testing.assert_array_equal(st, o.read_stats())
    assert st[2] > 0  # racket hits
    b.close()
