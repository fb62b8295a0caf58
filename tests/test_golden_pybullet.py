"""Replay of recorded trajectories (tennisbot_rl_b200/trajectory.py) through the CPU oracle and through the CUDA path.

Two files can sit in tests/golden/:
  pybullet_traj.npz      recorded from the UNMODIFIED reference on real PyBullet by tools/record_golden_pybullet.py.  Cannot be
                         produced in the build image (no pybullet wheel, no network): its tests are skipped until it is
                         supplied.  It is what turns "parity unpinned" into pinned.
  selfrecorded_traj.npz  produced by the in-repo oracle (tools/make_selfrecorded_golden.py).  Pins nothing about Bullet; it
                         keeps this file's plumbing - placement from the recorded reset state, action tape, comparison of
                         every recorded quantity, both env kinds, both implementations - running in CI.
"""
from pathlib import Path

import numpy as np
import pytest

from tennisbot_rl_b200 import trajectory as tj

GOLD = Path(__file__).parent / "golden"
FILES = {"pybullet": GOLD / "pybullet_traj.npz", "selfrecorded": GOLD / "selfrecorded_traj.npz"}
# bars: a self-recording must come back to rounding (the CUDA path's straight-line substeps differ from the oracle's generic
# one at 1e-16 per substep); a PyBullet recording is compared at the parity bar of the north star
TOL = {"pybullet": dict(state=1e-6, out=1e-5), "selfrecorded": dict(state=5e-8, out=2e-6)}
EV_RACKET_LOW = 64


def _load(which):
    p = FILES[which]
    if not p.exists():
        pytest.skip(f"{p.name} is not available in this environment (parity against PyBullet stays unpinned)")
    return tj.load(p)


def _compare_step(which, env_id, rec, t, live, state, obs, reward, done, events, low_seen):
    """live: episodes that still run at step t.  The racket's pose is outside the parity horizon of a PyBullet recording once
    the hull has reached the floor (racket-court contact is a modelling decision)."""
    tol = TOL[which]
    np.testing.assert_array_equal(done[live] != 0, rec["done"][live, t])
    for bit, col in ((1, 0), (2, 1), (4, 2)):
        np.testing.assert_array_equal((events[live] & bit) != 0, rec["contact"][live, t, col])
    np.testing.assert_allclose(reward[live], rec["reward"][live, t], atol=tol["out"])
    cols = np.ones(22, bool)
    ok = live.copy()
    if which == "pybullet":
        ok &= ~low_seen
    np.testing.assert_allclose(state[ok][:, :22][:, cols], rec["state"][ok, t, :22], atol=tol["state"])
    np.testing.assert_allclose(state[live][:, 13:22], rec["state"][live, t, 13:22], atol=tol["state"])  # the ball, always
    np.testing.assert_allclose(obs[ok], rec["obs"][ok, t], atol=tol["out"])
    assert (state[live][:, 29] == rec["state"][live, t, 29]).all()  # step_count


@pytest.mark.parametrize("which", ["pybullet", "selfrecorded"])
@pytest.mark.parametrize("env_id", ["SwingRacket-v0", "Tennisbot-v0"])
def test_oracle_replays_recording(oracle_lib, which, env_id):
    meta, envs = _load(which)
    rec = envs[env_id]
    E, T = rec["action"].shape[:2]
    o = oracle_lib.OracleEnv(env_id, E, auto_reset=False, threads=4)
    if meta.get("racket_scale", 1.0) != 1.0:
        o.set_param("racket_scale", meta["racket_scale"])
    o.reset(init=rec["init"])
    s0 = o.get_state()
    np.testing.assert_allclose(s0[:, :28], rec["reset_state"][:, :28], atol=TOL[which]["state"])  # the placement record is complete
    low_seen = np.zeros(E, bool)
    for t in range(T):
        live = rec["length"] > t
        if not live.any():
            break
        out = o.step(rec["action"][:, t], want_obs64=True)
        low_seen |= (out["events"] & EV_RACKET_LOW) != 0
        _compare_step(which, env_id, rec, t, live, o.get_state(), out["obs64"], out["reward"], out["done"], out["events"], low_seen)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["pybullet", "selfrecorded"])
@pytest.mark.parametrize("env_id", ["SwingRacket-v0", "Tennisbot-v0"])
def test_cuda_path_replays_recording(which, env_id):
    import torch

    from tennisbot_rl_b200.batch import TennisBatch

    meta, envs = _load(which)
    rec = envs[env_id]
    E, T = rec["action"].shape[:2]
    b = TennisBatch(env_id, E, precision="f64", auto_reset=False)
    if meta.get("racket_scale", 1.0) != 1.0:
        b.set_param("racket_scale", meta["racket_scale"])
    b.reset(init=rec["init"])
    np.testing.assert_allclose(b.get_state().cpu().numpy()[:, :28], rec["reset_state"][:, :28], atol=TOL[which]["state"])
    low_seen = np.zeros(E, bool)
    acts = torch.from_numpy(rec["action"]).cuda()
    for t in range(T):
        live = rec["length"] > t
        if not live.any():
            break
        obs, rew, done, _, ev = b.step(acts[:, t].contiguous())
        ev = ev.cpu().numpy()
        low_seen |= (ev & EV_RACKET_LOW) != 0
        _compare_step(which, env_id, rec, t, live, b.get_state().cpu().numpy(), obs.cpu().numpy().astype(np.float64),
                      rew.cpu().numpy().astype(np.float64), done.cpu().numpy(), ev, low_seen)
    b.close()


def test_recording_format_round_trip(tmp_path):
    w = tj.EpisodeWriter("Tennisbot-v0")
    rng = np.random.default_rng(0)
    for n in (3, 5):
        s = rng.normal(size=32)
        w.begin(s, tj.init_from_reset_state("Tennisbot-v0", s))
        for _ in range(n):
            w.step(rng.normal(size=2), rng.normal(size=32), rng.normal(size=12), 1.5, False, [True, False, False])
    tj.save(tmp_path / "t.npz", [w], {"producer": "test", "engine": "none", "engine_params": {"fixedTimeStep": 1 / 240}, "racket_scale": 1.0})
    meta, envs = tj.load(tmp_path / "t.npz")
    r = envs["Tennisbot-v0"]
    assert meta["engine_params"]["fixedTimeStep"] == 1 / 240 and r["action"].shape == (2, 5, 2) and list(r["length"]) == [3, 5]
    assert r["contact"][1, 4, 0] and not r["contact"][0, 4, 0] and r["init"][0, 2] == pytest.approx(r["reset_state"][0, 2] - 0.5)
