"""Parity harness shared by the GPU tests: replays identical initial states and action sequences through the
CUDA path (via the C ABI) and the CPU oracle, and compares events / done exactly and poses / rewards within a
stated tolerance."""
import numpy as np

from oracle import binding as ob


def reference_reset_params(kind, n, rng):
    """Initial placements drawn from the reference's reset() ranges (swingracket_env.py:161-173,
    tennisbot_env.py:227-246) with a harness-owned generator (the reference's own seeding is non-functional)."""
    init = np.zeros((n, 8))
    if kind == ob.ENV_SWING:
        init[:, 0] = rng.uniform(5.5, 11, n)
        init[:, 1] = rng.uniform(-4, 4, n)
        init[:, 2] = 0.6
        init[:, 3] = -3 - 9 * rng.uniform(0, 1, n)
        init[:, 4] = rng.uniform(-5, 5, n)
    else:
        init[:, 0] = rng.uniform(7.5, 12.5, n)
        init[:, 1] = rng.uniform(-5, 5, n)
        init[:, 2] = rng.uniform(0.2, 0.21, n)
        init[:, 3] = rng.uniform(25, 37.5, n)
        init[:, 4] = rng.uniform(-10, 10, n)
        init[:, 5] = rng.uniform(-12, -6, n)
        init[:, 6] = rng.uniform(-1, 1, n)
        init[:, 7] = rng.uniform(1, 1.5, n)
    return init


class ParityReport:
    def __init__(self):
        self.steps = 0
        self.max_obs_err = 0.0
        self.max_reward_err = 0.0
        self.max_state_err = 0.0
        self.event_mismatch_hard = 0   # discrete mismatch with the oracle margin above the band: a bug
        self.event_mismatch_near = 0   # discrete mismatch inside the band: fp32 near-threshold flip
        self.dropped = 0
        self.compared = 0

    def __repr__(self):
        return ("ParityReport(steps=%d compared=%d obs_err=%.3e reward_err=%.3e state_err=%.3e hard=%d near=%d dropped=%d)"
                % (self.steps, self.compared, self.max_obs_err, self.max_reward_err, self.max_state_err,
                   self.event_mismatch_hard, self.event_mismatch_near, self.dropped))


def run_parity(batch, oracle, steps, action_fn, band=0.0, check_state_every=0, obs0=None, event_mask=0xff,
               drop_after=None):
    """Step `batch` (TennisBatch) and `oracle` (OracleEnv) with the same actions = action_fn(t, last_oracle_obs).
    band: oracle margin (metres) under which a discrete mismatch counts as a near-threshold flip; such envs are
    dropped from the comparison from then on (their trajectories legitimately diverge).
    event_mask: event bits that are compared.  drop_after: {event bit: k} - an env leaves the comparison for good
    once the oracle has reported that bit on more than k steps (float32 only: the contact model turns metres of
    penetration into m/s with a factor 1/dt = 240, so every impact amplifies float32 position rounding; after a
    racket impact or a couple of floor bounces no useful pose tolerance is left)."""
    import torch

    n = batch.num_envs
    rep = ParityReport()
    valid = np.ones(n, bool)
    last_obs = obs0
    seen = {bit: np.zeros(n, np.int64) for bit in (drop_after or {})}
    for t in range(steps):
        a = action_fn(t, last_obs).astype(np.float32)
        g_obs, g_rew, g_done, g_term, g_ev = batch.step(torch.from_numpy(a).to(batch.device))
        torch.cuda.synchronize()
        o = oracle.step(a, want_margin=True)
        last_obs = o["obs"]
        g_obs, g_rew, g_done, g_term, g_ev = (x.cpu().numpy() for x in (g_obs, g_rew, g_done, g_term, g_ev))
        disc = (g_done != o["done"]) | ((g_ev & event_mask) != (o["events"] & event_mask))
        # an env whose oracle margin |quantity - threshold| fell inside the band this step took a discrete
        # decision the float32 path may legitimately take one substep apart: it is "near", counted, and left
        # out of the value comparison from here on (band = 0 for the float64 path: nothing is excused)
        near = valid & (o["margin"] <= band) if band > 0 else np.zeros(n, bool)
        rep.event_mismatch_near += int((near & disc).sum())
        rep.event_mismatch_hard += int((valid & ~near & disc).sum())
        valid &= ~(near | disc)
        for bit, k in (drop_after or {}).items():
            seen[bit] += (o["events"] & bit) != 0
            valid &= seen[bit] <= k
        rep.dropped = int((~valid).sum())
        v = valid
        rep.compared += int(v.sum())
        if v.any():
            rep.max_obs_err = max(rep.max_obs_err, float(np.abs(g_obs[v].astype(np.float64) - o["obs"][v]).max()))
            rep.max_reward_err = max(rep.max_reward_err, float(np.abs(g_rew[v].astype(np.float64) - o["reward"][v]).max()))
            d = v & (o["done"] != 0)
            if d.any():
                rep.max_obs_err = max(rep.max_obs_err, float(np.abs(g_term[d].astype(np.float64) - o["terminal_obs"][d]).max()))
        if check_state_every and (t + 1) % check_state_every == 0 and v.any():
            gs = batch.get_state().cpu().numpy()
            os_ = oracle.get_state()
            rep.max_state_err = max(rep.max_state_err, float(np.abs(gs[v] - os_[v]).max()))
        rep.steps += 1
    return rep, valid
