"""The scripted scenes of the reference's playground.py (tennisbot_rl_b200/playground.py) on the CUDA path against the same
runner over the CPU oracle: first-contact frames identical, final poses within the parity bar."""
import numpy as np
import pytest

from tennisbot_rl_b200 import playground as pg


class OracleBackend:
    """The oracle behind the surface the scenario runners use (see playground.BatchBackend)."""

    def __init__(self, ob, env_id, n):
        self.o = ob.OracleEnv(env_id, n, auto_reset=False, threads=4)
        self.n = n

    def set_param(self, k, v):
        self.o.set_param(k, v)

    def set_control_mode(self, m):
        self.o.set_control_mode(m)

    def reset(self, init):
        self.o.reset(init=init)

    def get_state(self):
        return self.o.get_state()

    def set_state(self, s):
        self.o.set_state(s)

    def step(self, actions):
        return self.o.step(np.asarray(actions, np.float32))["events"]

    def close(self):
        self.o.close()


def _scenes(n):
    rng = np.random.default_rng(5)
    force = np.tile([-400.0, 50.0, 400.0], (n, 1))
    force[1:] *= rng.uniform(0.5, 1.0, (n - 1, 1))           # row 0 is playground.py's own swing
    torque = np.tile([0.0, 0.3, -0.2], (n, 1))
    torque[1:] += rng.uniform(-0.5, 0.5, (n - 1, 3))
    return force, torque


def test_swing_scene_on_the_oracle(oracle_lib):
    """playground.py --swing: the scripted swing (400 N for 20 frames) meets the ball at the end of the swing and sends it
    over the net - with the reference's own numbers far beyond the court's end at x = -14, past the goal at (-12, 0)."""
    n = 16
    force, torque = _scenes(n)
    r = pg.swing_scene(OracleBackend(oracle_lib, "SwingRacket-v0", n), force=force, torque=torque, frames=700)
    assert 15 <= r["first_racket_contact"][0] < 25          # the reference script hits the ball during its swing
    assert (r["first_racket_contact"] >= 0).all() and (r["ball_pos"][:, 0] < 0).all()  # every variation clears the net
    assert r["ball_pos"][0, 0] < -14 and r["first_court_contact"][0] == -1          # ... the reference's own leaves the court
    assert np.isfinite(r["ball_pos"]).all() and np.isfinite(r["racket_pos"]).all()


@pytest.mark.gpu
def test_swing_scene_matches_oracle(oracle_lib):
    n = 64
    force, torque = _scenes(n)
    ref = pg.swing_scene(OracleBackend(oracle_lib, "SwingRacket-v0", n), force=force, torque=torque, frames=600)
    be = pg.BatchBackend("SwingRacket-v0", n)
    got = pg.swing_scene(be, force=force, torque=torque, frames=600)
    be.close()
    for k in ("first_racket_contact", "first_court_contact", "first_goal_contact"):
        np.testing.assert_array_equal(got[k], ref[k])
    # a ball that has come to rest on the court amplifies rounding (rolling contact): bar of the gravity -25 parity scene
    np.testing.assert_allclose(got["ball_pos"], ref["ball_pos"], atol=2e-5)
    np.testing.assert_allclose(got["racket_pos"], ref["racket_pos"], atol=1e-6)


@pytest.mark.gpu
def test_pid_hold_scene_matches_oracle(oracle_lib):
    n = 64
    rng = np.random.default_rng(9)
    base = np.stack([rng.uniform(5, 12, n), rng.uniform(-3, 3, n), rng.uniform(0, 2, n)], 1)      # racket.random_pos, playground.py:84
    ball = np.stack([rng.uniform(-12, -3, n), rng.uniform(-4, 4, n), rng.uniform(0.1, 3, n)], 1)  # ball.random_pos, :81
    ref = pg.pid_hold_scene(OracleBackend(oracle_lib, "Tennisbot-v0", n), base, ball, frames=800)
    be = pg.BatchBackend("Tennisbot-v0", n)
    got = pg.pid_hold_scene(be, base, ball, frames=800)
    be.close()
    for k in ("first_racket_contact", "first_court_contact"):
        np.testing.assert_array_equal(got[k], ref[k])
    np.testing.assert_allclose(got["racket_pos"], ref["racket_pos"], atol=1e-6)
    np.testing.assert_allclose(got["ball_pos"], ref["ball_pos"], atol=2e-5)
    # the PIDs pull the racket towards targetPos = (min(13, ball_x + 20), ball_y, 0.5): y is reached, the +-10 N limit
    # against 39 N of weight leaves z to gravity (the reference's orientation PIDs and this limit are what playground.py shows)
    assert np.abs(ref["racket_pos"][:, 1] - ref["target"][:, 1]).mean() < np.abs(base[:, 1] - ref["target"][:, 1]).mean()
