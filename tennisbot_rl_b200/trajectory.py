"""On-disk trajectory format of the parity harness: one .npz holding, per env id, E episodes of at most T
steps each, with everything needed to REPLAY them - the state right after reset() (before step 1), the explicit placement record
the C ABI's tb_reset_from takes, the action tape - and everything to COMPARE - per-step canonical state records, observations,
rewards, done flags and the three contact predicates the reference envs query (swingracket_env.py:99,111,119,
tennisbot_env.py:170).  Producers: tools/record_golden_pybullet.py (the unmodified reference on real PyBullet; where it exists)
and tools/make_selfrecorded_golden.py (the in-repo oracle; plumbing only, labelled as such).  Consumer:
tests/test_golden_pybullet.py, which drives the oracle AND the CUDA path from the file.

Keys, prefixed "<env id>/":
  reset_state float64 [E, 32]   canonical record (include/tennisbot_b200.h TB_S_*) after reset(), velocities zero
  init        float64 [E, 8]    placement for tb_reset_from: swing = racket base x,y,z, goal x,y, 0,0,0;
                                hit = racket base x,y,z, shoot force x,y, ball x,y,z
  action      float32 [E, T, A]
  state       float64 [E, T, 32] record AFTER step t          obs float64 [E, T, O]     reward float64 [E, T]
  done        bool    [E, T]                                   contact bool [E, T, 3]  (racket-ball, court-ball, goal-ball)
  length      int64   [E]        steps recorded for the episode (the rest of the T axis is zero)
plus the un-prefixed "meta" (JSON string): producer, engine (pybullet version or "oracle"), engine_params (structured:
fixedTimeStep, numSolverIterations, erp, contactERP, ... as returned by getPhysicsEngineParameters()), racket_scale.
"""
import json

import numpy as np

STATE_WORDS, INIT_WORDS = 32, 8
ACT_DIM = {"SwingRacket-v0": 6, "Tennisbot-v0": 2}
OBS_DIM = {"SwingRacket-v0": 6, "Tennisbot-v0": 12}
MAX_STEPS = {"SwingRacket-v0": 26, "Tennisbot-v0": 1001}
RACKET_COM_Z = 0.5  # racket.urdf:18-19: the inertial origin, i.e. the COM, sits 0.5 m up the link's z axis


class EpisodeWriter:
    """Accumulates episodes of one env id; `arrays()` returns the padded [E, T, ...] blocks."""

    def __init__(self, env_id):
        self.env_id = env_id
        self.eps = []

    def begin(self, reset_state, init):
        self.eps.append(dict(reset_state=np.asarray(reset_state, np.float64), init=np.asarray(init, np.float64), action=[], state=[],
                             obs=[], reward=[], done=[], contact=[]))

    def step(self, action, state, obs, reward, done, contact):
        e = self.eps[-1]
        e["action"].append(np.asarray(action, np.float32)); e["state"].append(np.asarray(state, np.float64))
        e["obs"].append(np.asarray(obs, np.float64)); e["reward"].append(float(reward)); e["done"].append(bool(done))
        e["contact"].append(np.asarray(contact, bool))

    def arrays(self):
        E = len(self.eps)
        T = max(len(e["action"]) for e in self.eps)
        A, O = ACT_DIM[self.env_id], OBS_DIM[self.env_id]
        out = dict(reset_state=np.zeros((E, STATE_WORDS)), init=np.zeros((E, INIT_WORDS)), action=np.zeros((E, T, A), np.float32),
                   state=np.zeros((E, T, STATE_WORDS)), obs=np.zeros((E, T, O)), reward=np.zeros((E, T)), done=np.zeros((E, T), bool),
                   contact=np.zeros((E, T, 3), bool), length=np.zeros(E, np.int64))
        for i, e in enumerate(self.eps):
            n = len(e["action"])
            out["reset_state"][i], out["init"][i], out["length"][i] = e["reset_state"], e["init"], n
            for k in ("action", "state", "obs", "reward", "done", "contact"):
                out[k][i, :n] = np.asarray(e[k])
        return out


def save(path, writers, meta):
    data = {"meta": np.array(json.dumps(meta))}
    for w in writers:
        for k, v in w.arrays().items():
            data[f"{w.env_id}/{k}"] = v
    np.savez_compressed(path, **data)


def load(path):
    """-> (meta dict, {env id: {key: array}})"""
    d = np.load(path, allow_pickle=False)
    meta = json.loads(str(d["meta"]))
    envs = {}
    for k in d.files:
        if "/" in k:
            env_id, key = k.split("/", 1)
            envs.setdefault(env_id, {})[key] = d[k]
    return meta, envs


def init_from_reset_state(env_id, s, racket_scale=1.0):
    """Placement record of tb_reset_from from the canonical state right after reset() (what a recorder can read back from
    the simulator): swing: base = spawn_pos (aux), goal; hit: base = COM - (0, 0, com_z) since the racket spawns upright
    (tennisbot_env.py:230-234), shoot force (aux x, y), ball position."""
    init = np.zeros(INIT_WORDS)
    if env_id == "SwingRacket-v0":
        init[0:3] = s[22:25]
        init[3:5] = s[25:27]
    else:
        init[0:3] = [s[0], s[1], s[2] - RACKET_COM_Z * racket_scale]
        init[3:5] = s[22:24]
        init[5:8] = s[13:16]
    return init
