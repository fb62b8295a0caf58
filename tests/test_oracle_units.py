"""CPU tests of the oracle against closed forms, known-answer vectors and the committed scene fixtures."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden"
SC = json.loads((GOLD / "scene_constants.json").read_text())


def test_philox_known_answers(oracle_lib):
    """Random123 kat_vectors for philox4x32-10: counter / key of zeros, of ones, and the digits of pi."""
    ph = oracle_lib.philox
    # (seed=key, env_id = c0 | c1<<32, episode = c2, word3 = c3)
    assert [hex(x) for x in ph(0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in ph(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == [
        "0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    key = 0xA4093822 | (0x299F31D0 << 32)
    ctr01 = 0x243F6A88 | (0x85A308D3 << 32)
    assert [hex(x) for x in ph(key, ctr01, 0x13198A2E, 0x03707344)] == [
        "0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_scene_constants_match_fixture(oracle_lib):
    """The numbers compiled into the oracle equal the ones parsed from the reference URDF / STL."""
    sc = oracle_lib.scene_constant
    assert sc("ball_radius") == SC["ball"]["radius"] and sc("ball_mass") == SC["ball"]["mass"]
    assert sc("racket_mass") == SC["racket"]["mass"] and sc("racket_com_z") == SC["racket"]["com_in_link"][2]
    assert sc("racket_half_x") == SC["racket"]["half_thickness_x"]
    assert [sc("floor_hx"), sc("floor_hy"), sc("floor_hz")] == [v / 2 for v in SC["court"]["floor_box_size"]]
    assert [sc("net_hx"), sc("net_hy"), sc("net_hz")] == [v / 2 for v in SC["court"]["net_box_size"]]
    assert sc("goal_radius") == SC["goal"]["radius"] and sc("goal_half_z") == SC["goal"]["length"] / 2
    n = int(sc("racket_outline_n"))
    assert n == len(SC["racket"]["outline_yz_link"]) == 38
    for i, (y, z) in enumerate(SC["racket"]["outline_yz_link"]):
        assert sc("racket_outline_y", i) == y and sc("racket_outline_z", i) == z
    np.testing.assert_allclose([sc("racket_inertia", i) for i in range(3)], SC["racket"]["inertia_aabb_box"], rtol=1e-12)
    # SURVEY Appendix A.3 values
    np.testing.assert_allclose(SC["racket"]["inertia_aabb_box"], [0.193656, 0.163256, 0.031040], atol=1e-6)
    assert abs(sc("contact_threshold") - 0.02 * math.sqrt(3) * 0.0335) < 1e-15
    # goal prism: PyBullet's 32-gon, first CCW vertex is (R sin(2 pi 31/32), R cos(2 pi 31/32))
    assert abs(sc("goal_vertex_x", 0) - 1.5 * math.sin(2 * math.pi * 31 / 32)) < 1e-15
    assert abs(math.hypot(sc("goal_vertex_x", 7), sc("goal_vertex_y", 7)) - 1.5) < 1e-14


def test_reset_geometry(oracle_lib):
    """reset obs = (rx + 0.5 sin 0.5, ry, rx - 0.1, ry, gx, gy); COM z = 0.6 + 0.5 cos 0.5."""
    o = oracle_lib.OracleEnv("SwingRacket-v0", 1, auto_reset=False)
    init = np.zeros((1, 8))
    init[0, :5] = [9.389706, -1.4842386, 0.6, -11.803556, 4.463079]
    obs = o.reset(init=init)[0]
    np.testing.assert_allclose(obs, [9.389706 + 0.5 * math.sin(0.5), -1.4842386, 9.289706, -1.4842386, -11.803556, 4.463079], rtol=1e-6)
    s = o.get_state()[0]
    assert abs(s[2] - (0.6 + 0.5 * math.cos(0.5))) < 1e-12          # COM z = 1.038791
    assert abs(s[oracle_lib.S_D0] - math.hypot(9.289706 + 11.803556, -1.4842386 - 4.463079)) < 1e-12
    np.testing.assert_allclose(s[oracle_lib.S_BP:oracle_lib.S_BP + 3], [9.289706, -1.4842386, 1.4], atol=1e-12)
    # ppo_swing.zip `_last_obs`: the recorded racket x (9.530986) lies between the spawn COM x and the ball x, which
    # only a COM-based observation can produce
    mon = json.loads((GOLD / "ppo_swing_monitor.json").read_text())
    assert obs[2] == pytest.approx(mon["last_obs"][2], abs=1e-6) and obs[0] > mon["last_obs"][0] > obs[2]


def test_hit_reset_geometry(oracle_lib):
    o = oracle_lib.OracleEnv("Tennisbot-v0", 1, auto_reset=False)
    init = np.array([[9.5, 1.0, 0.205, 30.0, -3.0, -9.0, 0.5, 1.2]])
    obs = o.reset(init=init)[0]
    np.testing.assert_allclose(obs, [9.5, 1.0, 0.705, 0, 0, 0, -9.0, 0.5, 1.2, 0, 0, 0], rtol=1e-6)
    s = o.get_state()[0]
    np.testing.assert_allclose(s[oracle_lib.S_AUX:oracle_lib.S_AUX + 3], [30.0, -3.0, 20.0])


def test_ballistic_step_with_damping(oracle_lib):
    """One isolated-ball step: v' = v + dt (g - v k (1 + |v|)), x' = x + dt v'  (Appendix A.2, k = 0.04)."""
    o = oracle_lib.OracleEnv("Tennisbot-v0", 1, auto_reset=False)
    s0 = np.zeros(32)
    s0[oracle_lib.S_RQ + 3] = 1
    s0[0:3] = [50, 50, 50]
    s0[oracle_lib.S_BP:oracle_lib.S_BP + 3] = [0, 0, 5]
    v = np.array([10.0, -3.0, 2.0])
    s0[oracle_lib.S_BV:oracle_lib.S_BV + 3] = v
    s1, bits = o.physics_step(s0)
    dt = 1 / 240
    k = 0.04 * (1 + np.linalg.norm(v))
    v1 = v + dt * (np.array([0, 0, -9.81]) - v * k)
    np.testing.assert_allclose(s1[oracle_lib.S_BV:oracle_lib.S_BV + 3], v1, rtol=1e-14)
    np.testing.assert_allclose(s1[oracle_lib.S_BP:oracle_lib.S_BP + 3], [0, 0, 5] + dt * v1, rtol=1e-14)
    assert bits == 0
    # at 18 m/s the drag deceleration (~13.7 m/s^2) exceeds gravity: SURVEY 0-6
    assert 18 * 0.04 * (1 + 18) > 9.81


def test_floor_contact_threshold_and_bounce(oracle_lib):
    """Court contact exists iff centre z <= 0.005 + 0.0335 + 1.1605e-3; the bounce follows A.6."""
    o = oracle_lib.OracleEnv("Tennisbot-v0", 1, auto_reset=False)
    thr = 0.005 + 0.0335 + 0.02 * math.sqrt(3) * 0.0335

    def drop(z, vz):
        s = np.zeros(32)
        s[oracle_lib.S_RQ + 3] = 1
        s[0:3] = [50, 50, 50]
        s[oracle_lib.S_BP:oracle_lib.S_BP + 3] = [3, 1, z]
        s[oracle_lib.S_BV + 2] = vz
        return o.physics_step(s)

    assert drop(thr + 1e-9, -1.0)[1] == 0
    s1, bits = drop(thr - 1e-9, -5.0)
    assert bits == oracle_lib.EV_COURT_BALL
    # normal row: v_n after = -0.81 v_n(before, incl. this step's gravity/drag) - (d + slop)/dt  with d = threshold gap
    dt = 1 / 240
    vz0 = -5.0 + dt * (-9.81 + 5.0 * 0.04 * 6.0)
    d = (thr - 1e-9) - 0.005 - 0.0335
    expect = 0.81 * (-vz0) - (d + 1e-5) / dt
    assert s1[oracle_lib.S_BV + 2] == pytest.approx(expect, rel=1e-9)
    # frictionless in the absence of tangential motion, no spin picked up
    np.testing.assert_allclose(s1[oracle_lib.S_BW:oracle_lib.S_BW + 3], 0, atol=1e-12)
    # slow contact (< 0.2 m/s): no restitution
    s2, _ = drop(0.0390, -0.1)
    assert s2[oracle_lib.S_BV + 2] < 0.81 * 0.15 + 0.2


def _poly_distance(poly, p):
    """Brute-force signed distance of 2-D point p to a CCW polygon (negative inside)."""
    poly = np.asarray(poly)
    best, inside = np.inf, True
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        e = b - a
        t = np.clip(np.dot(p - a, e) / np.dot(e, e), 0, 1)
        best = min(best, np.linalg.norm(p - (a + t * e)))
        if e[0] * (p[1] - a[1]) - e[1] * (p[0] - a[0]) < 0:
            inside = False
    return -best if inside else best


def test_racket_hull_distance_matches_bruteforce(oracle_lib):
    o = oracle_lib.OracleEnv("SwingRacket-v0", 1)
    poly = np.array(SC["racket"]["outline_yz_link"]) - [0, 0.5]
    hx = SC["racket"]["half_thickness_x"]
    rng = np.random.default_rng(0)
    for _ in range(400):
        p = rng.uniform([-0.1, -0.3, -0.7], [0.1, 0.3, 0.4])
        d, n, q = o.racket_core_distance(p)
        d2 = _poly_distance(poly, p[1:])
        ex = abs(p[0]) - hx
        if d2 > 0 and ex > 0:
            ref = math.hypot(d2, ex)
        elif d2 > 0:
            ref = d2
        elif ex > 0:
            ref = ex
        else:
            ref = max(d2, ex)  # inside: minimum translation distance, negative
        assert d == pytest.approx(ref, abs=1e-12)
        assert np.linalg.norm(n) == pytest.approx(1.0, abs=1e-12)
        if d > 0:  # outside: q is the closest point, p = q + d n
            np.testing.assert_allclose(q + d * n, p, atol=1e-12)
    # the ball starts 0.4223 m in front of the racket face
    c, s_ = math.cos(0.5), math.sin(0.5)
    rel_world = np.array([-0.1 - 0.5 * s_, 0.0, 0.8 - 0.5 * c])           # ball - COM at reset
    rel_local = np.array([c * rel_world[0] - s_ * rel_world[2], 0.0, s_ * rel_world[0] + c * rel_world[2]])
    d, n, _ = o.racket_core_distance(rel_local)
    assert d - 0.0335 - 0.001 == pytest.approx(0.4223, abs=2e-4) and n[0] == -1.0


def test_goal_prism_and_box_distance(oracle_lib):
    o = oracle_lib.OracleEnv("SwingRacket-v0", 1)
    # above the goal top: distance to the core = z - 0.125
    d, n, _ = o.goal_core_distance([0.3, -0.2, 0.4])
    assert d == pytest.approx(0.275) and list(n) == [0, 0, 1]
    # radially outside: apothem of the 32-gon is 1.5 cos(pi/32) = 1.49278
    ang = 2 * math.pi * (5.5 / 32)
    d, n, _ = o.goal_core_distance([2.0 * math.sin(ang), 2.0 * math.cos(ang), 0.0])
    assert d == pytest.approx(2.0 - 1.5 * math.cos(math.pi / 32), abs=1e-12)
    d, n, q = oracle_lib.box_core_distance([14, 7, 0.005], 0.001, [1.0, 2.0, 0.5])
    assert d == pytest.approx(0.5 - 0.004) and list(n) == [0, 0, 1]
    d, n, q = oracle_lib.box_core_distance([14, 7, 0.005], 0.001, [14.5, 0, 0.3])   # edge region
    assert d == pytest.approx(math.hypot(0.5 + 0.001, 0.3 - 0.004))


def test_zero_action_episode(oracle_lib):
    """Zero actions: the racket free-falls, the ball drops 1.4 m, grazes the handle and lands at physics step 133
    (SURVEY 8(c)); exactly 26 env steps; reward = moved() + contact bonus only before step 25."""
    o = oracle_lib.OracleEnv("SwingRacket-v0", 1, auto_reset=False)
    init = np.zeros((1, 8))
    init[0, :5] = [9.5, 0.0, 0.6, -8.0, 1.0]
    o.reset(init=init)
    for k in range(26):
        r = o.step(np.zeros((1, 6), np.float32))
        assert bool(r["done"][0]) == (k == 25)
    s = o.get_state()[0]
    assert s[oracle_lib.S_STEP] == 133
    assert r["events"][0] & oracle_lib.EV_COURT_BALL and not r["events"][0] & oracle_lib.EV_TIMEOUT
    assert 0 < r["reward"][0] < 1.0
    assert o.read_stats()[8] == 133 and o.read_stats()[9] == 26


def test_hit_env_episode_logic(oracle_lib):
    """No reward / done in the 5 shoot frames; episode ends when the ball passes the racket (x_b - x_r >= 0.5) or
    after 1000 steps; tier reward from the yz miss distance (tennisbot_env.py:90-102,138-203)."""
    n = 256
    o = oracle_lib.OracleEnv("Tennisbot-v0", n, seed=3, auto_reset=False, threads=4)
    o.reset()
    done_at = np.full(n, -1)
    for k in range(1001):
        r = o.step(np.zeros((n, 2), np.float32))
        if k < 4:
            assert not r["done"].any() and (r["reward"] == 0).all()
        newly = (r["done"] != 0) & (done_at < 0)
        if newly.any():
            ev = r["events"][newly]
            passed = (ev & oracle_lib.EV_BALL_PASSED) != 0
            assert (passed | ((ev & oracle_lib.EV_TIMEOUT) != 0)).all()
            assert set(np.unique(r["reward"][newly & (r["events"] & oracle_lib.EV_RACKET_BALL == 0)])) <= {0, 1, 5, 10, 15, 20}
            st = o.get_state()
            x_rel = st[newly, oracle_lib.S_BP] - st[newly, oracle_lib.S_RP]
            assert (x_rel[passed] >= 0.5).all()
        done_at[newly] = k
    assert (done_at >= 0).all() and done_at.max() == 1000 and 300 < np.median(done_at) < 900


def test_param_override_changes_dynamics(oracle_lib):
    """Every recalled Bullet constant is a named parameter: zero damping makes the ball fly further."""
    names = oracle_lib.OracleEnv.param_names()
    assert {"lin_damping", "contact_erp", "solver_residual", "racket_scale", "contact_threshold"} <= set(names)
    o = oracle_lib.OracleEnv("Tennisbot-v0", 1, auto_reset=False)
    s0 = np.zeros(32); s0[6] = 1; s0[0:3] = 50; s0[15] = 5; s0[16] = 10
    a, _ = o.physics_step(s0)
    o.set_param("lin_damping", 0.0)
    assert o.get_param("lin_damping") == 0.0
    b, _ = o.physics_step(s0)
    assert b[16] == 10.0 and a[16] < 10.0
    with pytest.raises(RuntimeError):
        o.set_param("no_such_parameter", 1.0)


class _SimplePID:
    """simple_pid.PID restated for the test (the package is not installed): proportional on error, derivative on
    measurement, integral and output clamped to the limits, first call has no derivative."""

    def __init__(self, kp, ki, kd, lim):
        self.kp, self.ki, self.kd, self.lim = kp, ki, kd, lim
        self.integral, self.last = 0.0, None

    def __call__(self, setpoint, x, dt):
        e = setpoint - x
        self.integral = float(np.clip(self.integral + self.ki * e * dt, -self.lim, self.lim))
        d = -self.kd * (x - self.last) / dt if self.last is not None else 0.0
        self.last = x
        return float(np.clip(self.kp * e + self.integral + d, -self.lim, self.lim))


def test_pid_control_mode(oracle_lib):
    """A9: Racket.apply_action (racket.py:66-89,103-122): force = (0,0,4) + PID(kp=3, ki=.01, kd=.1, +-10 N)(pos)."""
    o = oracle_lib.OracleEnv("Tennisbot-v0", 1, auto_reset=False)
    o.set_control_mode("pid")
    init = np.array([[9.5, 1.0, 0.205, 30.0, -3.0, -9.0, 0.5, 1.2]])
    o.reset(init=init)
    pids = [_SimplePID(3.0, 0.01, 0.1, 10.0) for _ in range(3)]
    dt, m, g = 1 / 240, 4.0, -9.81
    pos = np.array([9.5, 1.0, 0.705])
    vel = np.zeros(3)
    target = np.array([9.8, 0.7, 1.5], np.float32)
    for k in range(50):
        f = np.array([pids[i](float(target[i]), pos[i], dt) for i in range(3)]) + [0, 0, 4.0]
        kdamp = 0.04 * (1 + np.linalg.norm(vel))
        vel = vel + dt * (f / m + [0, 0, g] - vel * kdamp)
        pos = pos + dt * vel
        r = o.step(target[None, :2])
        np.testing.assert_allclose(r["obs"][0][:3], pos, rtol=0, atol=2e-6)
        np.testing.assert_allclose(r["obs"][0][3:6], vel, rtol=0, atol=2e-6)
    # gains are parameters (Racket.update_pid); reset clears the controller memory
    o.set_param("pid_kp", 10.0)
    assert o.get_param("pid_kp") == 10.0
