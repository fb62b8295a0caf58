"""The two scripted scenes of the reference's playground.py as batched scenario runners.

playground.py is the reference's physics sandbox: no gym env, no reward, just the scene objects driven by a script.
Both scripts are expressible through the env-step kernels - placements through the explicit-placement / state-injection entry
points, the scripts as action tapes - so N variations of a scene run as one batch:

  swing scene (playground.py:38-62)     racket at `position` tilted rpy (0, 0.5, 0), ball 0.8 m above it, goal at (-12, 0); a
      constant force / torque on the racket for SWING_FRAME_COUNT = 20 frames, gravity compensation only afterwards.  In
      SwingRacket-v0's terms: action = force / 400, torque / 5 for 20 control steps, then zero actions.  The env would fast-
      forward after its 25th control step; the runner keeps it in the control phase by rewinding the step counter.
  PID-hold scene (playground.py:64-116)  upright racket under the three position PIDs (Racket.apply_pid_force_torque,
      kp 10, ki 0.001, kd 2) holding a target behind the ball's line, ball shot with (5.5, 0, 12.1) N on frames 11 .. 49.
      In Tennisbot-v0's terms: TB_CONTROL_PID, action = target x, y (z = pid_hit_z), shoot window = parameters
      shoot_start / shoot_frames.

A runner talks to a `backend` with reset(init) / get_state() / set_state() / step(actions) -> event bytes / set_param /
set_control_mode: BatchBackend wraps TennisBatch (the CUDA path); the tests wrap the CPU oracle the same way and compare.
"""
import numpy as np

EV_RACKET_BALL, EV_COURT_BALL, EV_GOAL_BALL = 1, 2, 4
S_STEP, S_AUX = 29, 22


class BatchBackend:
    """TennisBatch (auto_reset off) behind the small surface the runners use."""

    def __init__(self, env_id, n, device=0, precision="f64"):
        import torch

        from .batch import TennisBatch

        self.torch = torch
        self.b = TennisBatch(env_id, n, device=device, precision=precision, auto_reset=False)
        self.n = n

    def set_param(self, k, v):
        self.b.set_param(k, v)

    def set_control_mode(self, m):
        self.b.set_control_mode(m)

    def reset(self, init):
        self.b.reset(init=init)

    def get_state(self):
        return self.b.get_state().cpu().numpy()

    def set_state(self, s):
        self.b.set_state(s)

    def step(self, actions):
        _, _, _, _, ev = self.b.step(self.torch.from_numpy(np.ascontiguousarray(actions, np.float32)).to(self.b.device))
        return ev.cpu().numpy()

    def close(self):
        self.b.close()


def _summarise(first_hit, first_court, first_goal, state):
    return dict(first_racket_contact=first_hit, first_court_contact=first_court, first_goal_contact=first_goal,
                racket_pos=state[:, 0:3].copy(), ball_pos=state[:, 13:16].copy(), ball_vel=state[:, 16:19].copy())


def _track(ev, t, first, bit):
    new = ((ev & bit) != 0) & (first < 0)
    first[new] = t


def swing_scene(backend, position=(3.0, 0.1, 0.5), goal=(-12.0, 0.0), force=(-400.0, 50.0, 400.0), torque=(0.0, 0.3, -0.2),
                swing_frames=20, frames=3000, rewind_every=20):
    """playground.py:38-62 for every row of `position` / `force` / `torque` (each broadcast to [N, 3]).  Returns the frame of
    the first racket-ball / court-ball / goal-ball contact per scene (-1: none) and the final racket / ball state."""
    n = backend.n
    pos, F, Tq = (np.broadcast_to(np.asarray(x, np.float64), (n, 3)) for x in (position, force, torque))
    init = np.zeros((n, 8))
    init[:, 0:3] = pos
    init[:, 3:5] = np.broadcast_to(np.asarray(goal, np.float64), (n, 2))
    backend.reset(init)
    swing = np.concatenate([F / 400.0, Tq / 5.0], 1).astype(np.float32)  # swingracket_env.py:76-79 read backwards
    if np.abs(swing).max() > 1:
        raise ValueError("force / torque outside what a SwingRacket-v0 action can express (|F| <= 400, |T| <= 5)")
    rest = np.zeros((n, 6), np.float32)  # racket.apply_target_action([0, 0, 4 * 9.81])
    first = [np.full(n, -1, np.int64) for _ in range(3)]
    for t in range(frames):
        if t and t % rewind_every == 0:  # stay in the env's control phase: one physics step per env step
            s = backend.get_state()
            s[:, S_STEP] = 0
            backend.set_state(s)
        ev = backend.step(swing if t < swing_frames else rest)
        for f, bit in zip(first, (EV_RACKET_BALL, EV_COURT_BALL, EV_GOAL_BALL)):
            _track(ev, t, f, bit)
    return _summarise(*first, backend.get_state())


def pid_hold_scene(backend, racket_base, ball_pos, kp=10.0, ki=0.001, kd=2.0, ball_force=5.5, shoot_start=11, shoot_frames=39,
                   frames=3000):
    """playground.py:64-116 for N (racket base, ball position) pairs (what Racket.random_pos / Ball.random_pos draw there).
    The racket holds targetPos = (min(13, ball_x + 20), ball_y, 0.5) under its position PIDs while the ball is shot at it."""
    n = backend.n
    base, ball = np.asarray(racket_base, np.float64).reshape(n, 3), np.asarray(ball_pos, np.float64).reshape(n, 3)
    for k, v in (("pid_kp", kp), ("pid_ki", ki), ("pid_kd", kd), ("pid_hit_z", 0.5), ("shoot_start", shoot_start),
                 ("shoot_frames", shoot_frames)):
        backend.set_param(k, v)
    backend.set_control_mode("pid")
    init = np.zeros((n, 8))
    init[:, 0:3] = base
    init[:, 3:5] = [ball_force, 0.0]       # ball.apply_force([BALL_FORCE, 0, BALL_FORCE * 2.2]), playground.py:99-100
    init[:, 5:8] = ball
    backend.reset(init)
    s = backend.get_state()
    s[:, S_AUX + 2] = ball_force * 2.2
    backend.set_state(s)
    target = np.stack([np.minimum(13.0, ball[:, 0] + 20.0), ball[:, 1]], 1).astype(np.float32)  # playground.py:87-90
    first = [np.full(n, -1, np.int64) for _ in range(3)]
    for t in range(frames):
        ev = backend.step(target)
        for f, bit in zip(first, (EV_RACKET_BALL, EV_COURT_BALL, EV_GOAL_BALL)):
            _track(ev, t, f, bit)
    out = _summarise(*first, backend.get_state())
    out["target"] = target
    return out


# ------------------------------------------------------------------------------------------------ curriculum (train.py:155-176)
PERCENT_THRESH = (3, 5, 10, 15, 25, 45, 70, 101)
SCALE_THRESH = (3, 2.6, 2.3, 2.1, 1.9, 1.7, 1.3, 1)


def curriculum_scale(num_timesteps, total_timesteps):
    """Racket scale of train.py's ProgressCallback for the given progress: the hull shrinks from 3x to 1x in eight stages."""
    progress = int(num_timesteps / total_timesteps * 100)
    for pct, scale in zip(PERCENT_THRESH, SCALE_THRESH):
        if progress < pct:
            return float(scale)
    return 1.0


class RacketScaleCurriculum:
    """train.py:155-176 for a TennisVecEnv: call on_rollout_start(num_timesteps) where SB3 would call the callback's
    _on_rollout_start; the scale takes effect at the env's next reset(), as in the reference (tennisbot_env.py:213-215,234)."""

    def __init__(self, env, total_timesteps):
        self.env, self.total = env, total_timesteps

    def on_rollout_start(self, num_timesteps):
        scale = curriculum_scale(num_timesteps, self.total)
        self.env.set_racket_scale(scale)
        return scale
