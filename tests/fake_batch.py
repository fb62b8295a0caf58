"""Oracle-backed stand-in for TennisBatch's host-buffer interface, so the VecEnv adapter's host logic (auto-reset
infos, episode accounting, truncation flags, copies) can be tested on a machine without a GPU."""
import numpy as np

from oracle import binding as ob


class FakeBatch:
    def __init__(self, env_id="SwingRacket-v0", num_envs=8, device=0, seed=0, precision="f64", auto_reset=True,
                 env_id_offset=0):
        self.o = ob.OracleEnv(env_id, num_envs, env_id_offset=env_id_offset, seed=seed, auto_reset=auto_reset)
        self.num_envs, self.obs_dim, self.act_dim = num_envs, self.o.obs_dim, self.o.act_dim
        self._hb = None

    def reset_host(self, mask=None):
        self._hb = dict(obs=self.o.reset(mask))
        return self._hb["obs"]

    def step_host(self, actions=None, want_terminal=True, want_events=True):
        r = self.o.step(actions)
        # like the real pinned buffers: the SAME arrays are overwritten by every call
        if self._hb is None or "reward" not in self._hb:
            self._hb = {k: v.copy() for k, v in r.items()}
        else:
            for k, v in r.items():
                self._hb[k][...] = v
        return self._hb

    def read_stats(self, clear=False):
        return self.o.read_stats(clear)

    def set_param(self, name, value):
        self.o.set_param(name, value)

    def get_param(self, name):
        return self.o.get_param(name)

    def close(self):
        self.o.close()
