"""Host-side logic that needs no GPU: spaces, VecEnv contract (over an oracle-backed fake batch), sharding maths."""
import numpy as np
import pytest

from tests.fake_batch import FakeBatch


def test_spaces_match_reference():
    from tennisbot_rl_b200 import spaces

    obs, act = spaces.swing_spaces()  # swingracket_env.py:29-39
    assert obs.shape == (6,) and act.shape == (6,)
    np.testing.assert_array_equal(obs.low, [-20, -10, -20, -10, -15, -5])
    np.testing.assert_array_equal(obs.high, [20, 10, 20, 10, 0, 5])
    np.testing.assert_array_equal(act.low, -np.ones(6))
    obs, act = spaces.hit_spaces()    # tennisbot_env.py:37-55
    assert obs.shape == (12,) and act.shape == (2,) and obs.dtype == np.float32
    np.testing.assert_array_equal(obs.low, [-20, -20, -5, -5, -5, -5, -20, -20, 0, -10, -10, -10])
    a = act.sample()
    assert act.contains(a) and a.dtype == np.float32
    # the saved PPO model stores the same bounds
    assert spaces.spaces_for("SwingRacket-v0")[0] == spaces.swing_spaces()[0]


def test_registry_shim():
    import tennisbot
    import tennisbot.envs as envs

    assert tennisbot.ENTRY_POINTS == {"Tennisbot-v0": "tennisbot.envs:TennisbotEnv",
                                      "SwingRacket-v0": "tennisbot.envs:SwingRacketEnv"}
    assert envs.SwingRacketEnv.metadata == {"render.modes": ["human"]}
    assert envs.TennisbotEnv._env_id == "Tennisbot-v0"


@pytest.fixture
def vec(monkeypatch):
    import tennisbot_rl_b200.vec_env as ve

    monkeypatch.setattr(ve, "TennisBatch", FakeBatch)
    return ve


@pytest.mark.parametrize("env_id", ["SwingRacket-v0", "Tennisbot-v0"])
def test_vecenv_contract(vec, env_id):
    n = 16
    env = vec.TennisVecEnv(env_id, n, seed=5, report_truncation=True)
    assert env.num_envs == n and env.observation_space.shape[0] == env.batch.obs_dim
    obs = env.reset()
    assert obs.shape == (n, env.batch.obs_dim) and obs.dtype == np.float32
    rng = np.random.default_rng(0)
    horizon = 60 if env_id == "SwingRacket-v0" else 1100
    prev_obs = obs
    ep_ret = np.zeros(n)
    seen_done = 0
    for t in range(horizon):
        a = rng.uniform(-1, 1, (n, env.batch.act_dim)).astype(np.float32)
        held = prev_obs.copy()
        obs, rew, dones, infos = env.step(a)
        np.testing.assert_array_equal(prev_obs, held)  # returned arrays are fresh copies (SB3 keeps _last_obs)
        assert rew.dtype == np.float32 and dones.dtype == bool and len(infos) == n
        ep_ret += rew
        for i in np.nonzero(dones)[0]:
            info = infos[i]
            assert info["terminal_observation"].shape == (env.batch.obs_dim,)
            assert info["episode"]["r"] == pytest.approx(ep_ret[i], rel=1e-6, abs=1e-6)
            if env_id == "SwingRacket-v0":
                assert info["episode"]["l"] == 26
                # auto-reset: the returned obs is the next episode's reset obs (ball 0.1 behind the racket base)
                assert obs[i, 2] == pytest.approx(obs[i, 0] - 0.5 * np.sin(0.5) - 0.1, abs=1e-5)
            assert info["TimeLimit.truncated"] == bool(info["events"] & 8 and not info["events"] & (2 | 4 | 16))
            ep_ret[i] = 0
            seen_done += 1
        for i in np.nonzero(~dones)[0]:
            assert infos[i] == {}
        prev_obs = obs
    assert seen_done >= n
    with pytest.raises(ValueError):
        env.step_async(np.zeros((n + 1, env.batch.act_dim), np.float32))
    assert env.env_is_wrapped(object) == [False] * n and env.get_attr("num_envs", [0, 1]) == [n, n]
    st = env.episode_statistics()
    assert st["episodes"] == seen_done and st["env_steps"] == n * horizon
    env.close()


def test_vecenv_reference_defaults(vec):
    """Reference behaviour by default: the envs are registered without max_episode_steps and return {} as info, so no
    TimeLimit.truncated key (SB3 would bootstrap from the terminal value otherwise); set_racket_scale takes effect at the
    next reset(), not under the episodes in flight (tennisbot_env.py:213-215,234)."""
    from tennisbot_rl_b200.playground import RacketScaleCurriculum, curriculum_scale

    env = vec.TennisVecEnv("Tennisbot-v0", 8, seed=1)
    env.reset()
    rng = np.random.default_rng(0)
    seen = 0
    for t in range(1010):
        _, _, dones, infos = env.step(rng.uniform(-1, 1, (8, 2)).astype(np.float32))
        for i in np.nonzero(dones)[0]:
            assert "TimeLimit.truncated" not in infos[i] and "events" in infos[i] and "episode" in infos[i]
            seen += 1
    assert seen >= 8
    cur = RacketScaleCurriculum(env, total_timesteps=1000)
    assert cur.on_rollout_start(0) == 3.0 and env.batch.get_param("racket_scale") == 1.0  # stored, not applied yet
    env.reset()
    assert env.batch.get_param("racket_scale") == 3.0
    env.set_racket_scale(1.3, apply_now=True)
    assert env.batch.get_param("racket_scale") == 1.3
    # train.py:166-176: percent_thresh [3, 5, 10, 15, 25, 45, 70, 101] -> scale [3, 2.6, 2.3, 2.1, 1.9, 1.7, 1.3, 1]
    assert [curriculum_scale(p * 10, 1000) for p in (0, 2, 3, 4, 5, 9, 10, 14, 15, 24, 25, 44, 45, 69, 70, 100)] == \
        [3, 3, 2.6, 2.6, 2.3, 2.3, 2.1, 2.1, 1.9, 1.9, 1.7, 1.7, 1.3, 1.3, 1, 1]
    env.close()


def test_stats_dict_and_shard_ranges():
    from tennisbot_rl_b200.sharding import shard_range
    from tennisbot_rl_b200.vec_env import stats_dict

    d = stats_dict([4, 104, 3, 1, 3, 0, int(40.0 * 2 ** 20), int((4 * 100.0 + 4 * 9.0) * 2 ** 10), 500, 104])
    assert d["mean_length"] == 26 and d["mean_return"] == 10.0 and d["std_return"] == pytest.approx(3.0)
    for total, world in ((8388608, 8), (1000, 3), (5, 8)):
        r = [shard_range(total, g, world) for g in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1
