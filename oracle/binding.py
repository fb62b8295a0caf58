"""ctypes binding of the CPU oracle (oracle/tb_oracle.h).  TEST INFRASTRUCTURE.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package `tennisbot_rl_b200` never imports this module.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libtb_oracle.so"

ENV_SWING, ENV_HIT = 0, 1
STATE_WORDS, INIT_WORDS, NUM_STATS = 32, 8, 10
EV_RACKET_BALL, EV_COURT_BALL, EV_GOAL_BALL, EV_TIMEOUT, EV_BALL_PASSED, EV_NET_BALL, EV_RACKET_LOW = 1, 2, 4, 8, 16, 32, 64
ENV_KINDS = {"SwingRacket-v0": ENV_SWING, "Tennisbot-v0": ENV_HIT, "swing": ENV_SWING, "hit": ENV_HIT}

# canonical state record offsets (tb_oracle.c enum)
S_RP, S_RQ, S_RV, S_RW, S_BP, S_BV, S_BW, S_AUX, S_GOAL, S_D0, S_RET, S_STEP, S_FLAGS, S_EPISODE = (
    0, 3, 7, 10, 13, 16, 19, 22, 25, 27, 28, 29, 30, 31)


def build(force=False):
    src = [HERE / "tb_oracle.c", HERE / "tb_oracle.h", HERE / "tbo_scene_data.h"]
    if force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in src):
        cc = os.environ.get("TB_ORACLE_CC", "gcc")
        subprocess.check_call([cc, "-O2", "-std=c11", "-fPIC", "-pthread", "-ffp-contract=off", "-shared",
                               "-o", str(LIB_PATH), str(HERE / "tb_oracle.c"), "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
        L.tbo_last_error.restype = C.c_char_p
        L.tbo_create.argtypes = [i32, i64, i64, u64, i32, C.POINTER(vp)]
        L.tbo_destroy.argtypes = [vp]
        L.tbo_destroy.restype = None
        L.tbo_set_threads.argtypes = [vp, i32]
        L.tbo_set_control_mode.argtypes = [vp, i32]
        L.tbo_set_param.argtypes = [vp, C.c_char_p, dbl]
        L.tbo_get_param.argtypes = [vp, C.c_char_p, C.POINTER(dbl)]
        L.tbo_param_name.restype = C.c_char_p
        L.tbo_param_name.argtypes = [i32]
        L.tbo_scene_constant.argtypes = [C.c_char_p, i32, C.POINTER(dbl)]
        L.tbo_reset.argtypes = [vp, vp, vp]
        L.tbo_reset_from.argtypes = [vp, vp, vp, vp]
        L.tbo_step.argtypes = [vp] + [vp] * 8
        L.tbo_rollout.argtypes = [vp, i32, i32, vp, vp, vp]
        L.tbo_get_state.argtypes = [vp, vp]
        L.tbo_set_state.argtypes = [vp, vp]
        L.tbo_read_stats.argtypes = [vp, vp, i32]
        L.tbo_physics_steps.argtypes = [vp]
        L.tbo_physics_steps.restype = i64
        L.tbo_philox4x32.argtypes = [u64, u64, C.c_uint32, C.c_uint32, vp]
        L.tbo_philox4x32.restype = None
        L.tbo_racket_core_distance.argtypes = [vp, vp, vp, vp]
        L.tbo_racket_core_distance.restype = dbl
        L.tbo_goal_core_distance.argtypes = [vp, vp, vp, vp]
        L.tbo_goal_core_distance.restype = dbl
        L.tbo_box_core_distance.argtypes = [vp, dbl, vp, vp, vp]
        L.tbo_box_core_distance.restype = dbl
        L.tbo_physics_step.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i32)]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().tbo_last_error().decode())


def philox(seed, env_id, episode, word3):
    out = np.zeros(4, np.uint32)
    lib().tbo_philox4x32(seed, env_id, episode, word3, _ptr(out))
    return out


def scene_constant(name, index=0):
    v = C.c_double()
    _check(lib().tbo_scene_constant(name.encode(), index, C.byref(v)))
    return v.value


def box_core_distance(half_ext, margin, p):
    h = np.asarray(half_ext, np.float64)
    p = np.asarray(p, np.float64)
    n, q = np.zeros(3), np.zeros(3)
    d = lib().tbo_box_core_distance(_ptr(h), margin, _ptr(p), _ptr(n), _ptr(q))
    return d, n, q


class OracleEnv:
    """Batch of N envs stepped by the double-precision CPU oracle."""

    def __init__(self, env="SwingRacket-v0", num_envs=1, env_id_offset=0, seed=0, auto_reset=True, threads=1):
        self.kind = ENV_KINDS[env] if isinstance(env, str) else int(env)
        self.n = int(num_envs)
        h = C.c_void_p()
        _check(lib().tbo_create(self.kind, self.n, env_id_offset, seed, int(auto_reset), C.byref(h)))
        self.h = h
        self.obs_dim = lib().tbo_obs_dim(self.kind)
        self.act_dim = lib().tbo_act_dim(self.kind)
        if threads != 1:
            self.set_threads(threads)

    def close(self):
        if getattr(self, "h", None):
            lib().tbo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass  # interpreter shutdown

    def set_threads(self, n):
        _check(lib().tbo_set_threads(self.h, int(n)))

    def set_control_mode(self, mode):
        _check(lib().tbo_set_control_mode(self.h, {"force": 0, "pid": 1}.get(mode, mode)))

    def set_param(self, name, value):
        _check(lib().tbo_set_param(self.h, name.encode(), float(value)))

    def get_param(self, name):
        v = C.c_double()
        _check(lib().tbo_get_param(self.h, name.encode(), C.byref(v)))
        return v.value

    @staticmethod
    def param_names():
        L = lib()
        return [L.tbo_param_name(i).decode() for i in range(L.tbo_num_params())]

    def reset(self, mask=None, init=None):
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        if init is None:
            _check(lib().tbo_reset(self.h, _ptr(m), _ptr(obs)))
        else:
            init = np.ascontiguousarray(init, np.float64).reshape(self.n, INIT_WORDS)
            _check(lib().tbo_reset_from(self.h, _ptr(init), _ptr(m), _ptr(obs)))
        return obs

    def step(self, actions, want_margin=False, want_obs64=False):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, self.act_dim)
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.uint8)
        term = np.zeros((self.n, self.obs_dim), np.float32)
        ev = np.zeros(self.n, np.uint8)
        margin = np.zeros(self.n, np.float64) if want_margin else None
        obs64 = np.zeros((self.n, self.obs_dim), np.float64) if want_obs64 else None
        _check(lib().tbo_step(self.h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done), _ptr(term), _ptr(ev), _ptr(margin), _ptr(obs64)))
        out = dict(obs=obs, reward=rew, done=done, terminal_obs=term, events=ev)
        if want_margin:
            out["margin"] = margin
        if want_obs64:
            out["obs64"] = obs64
        return out

    def rollout(self, k_steps, action_mode=0):
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        rs = np.zeros(self.n, np.float32)
        dc = np.zeros(self.n, np.int32)
        _check(lib().tbo_rollout(self.h, action_mode, int(k_steps), _ptr(obs), _ptr(rs), _ptr(dc)))
        return dict(obs=obs, reward_sum=rs, done_count=dc)

    def get_state(self):
        s = np.zeros((self.n, STATE_WORDS), np.float64)
        _check(lib().tbo_get_state(self.h, _ptr(s)))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float64).reshape(self.n, STATE_WORDS)
        _check(lib().tbo_set_state(self.h, _ptr(s)))

    def read_stats(self, clear=False):
        st = np.zeros(NUM_STATS, np.int64)
        _check(lib().tbo_read_stats(self.h, _ptr(st), int(clear)))
        return st

    def physics_steps(self):
        return int(lib().tbo_physics_steps(self.h))

    def racket_core_distance(self, p_local):
        p = np.asarray(p_local, np.float64)
        n, q = np.zeros(3), np.zeros(3)
        d = lib().tbo_racket_core_distance(self.h, _ptr(p), _ptr(n), _ptr(q))
        return d, n, q

    def goal_core_distance(self, p_rel):
        p = np.asarray(p_rel, np.float64)
        n, q = np.zeros(3), np.zeros(3)
        d = lib().tbo_goal_core_distance(self.h, _ptr(p), _ptr(n), _ptr(q))
        return d, n, q

    def physics_step(self, state32, f_racket=(0, 0, 0), t_racket=(0, 0, 0), f_ball=(0, 0, 0)):
        s = np.array(state32, np.float64).reshape(STATE_WORDS).copy()
        fr, tr, fb = (np.asarray(v, np.float64) for v in (f_racket, t_racket, f_ball))
        bits = C.c_int()
        _check(lib().tbo_physics_step(self.h, _ptr(s), _ptr(fr), _ptr(tr), _ptr(fb), C.byref(bits)))
        return s, bits.value
