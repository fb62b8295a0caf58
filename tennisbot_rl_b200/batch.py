"""TennisBatch: N lock-step envs on one B200, stepped by the CUDA kernels behind the C ABI.

PyTorch is used for device memory and streams only; all arithmetic of the env step happens in
libtennisbot_b200.so (csrc/tb_kernels.cu).  Replaces, for a batch, what one reference env object does with
p.connect / loadURDF / stepSimulation / getContactPoints (tennisbot/envs/*.py, tennisbot/resources/*.py).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

ENV_IDS = {"SwingRacket-v0": _lib.ENV_SWING, "Tennisbot-v0": _lib.ENV_HIT, "swing": _lib.ENV_SWING, "hit": _lib.ENV_HIT}
PRECISIONS = {"f32": _lib.F32, "f64": _lib.F64, "float32": _lib.F32, "float64": _lib.F64}


def _dptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _hptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class TennisBatch:
    def __init__(self, env_id="SwingRacket-v0", num_envs=4096, device=0, seed=0, precision="f64", auto_reset=True,
                 env_id_offset=0):
        if not torch.cuda.is_available():
            raise _lib.TennisbotLibraryError("TennisBatch needs a CUDA device: the env step has no CPU fallback")
        self.lib = _lib.load()
        self.kind = ENV_IDS[env_id] if isinstance(env_id, str) else int(env_id)
        self.precision = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.num_envs = int(num_envs)
        self.obs_dim = self.lib.tb_obs_dim(self.kind)
        self.act_dim = self.lib.tb_act_dim(self.kind)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists before the library uses it
        cfg = _lib.TbConfig(C.sizeof(_lib.TbConfig), self.kind, self.precision, self.device.index, self.num_envs,
                            int(env_id_offset), int(seed) & (2 ** 64 - 1), int(bool(auto_reset)), 0)
        h = C.c_void_p()
        _lib.check(self.lib.tb_create(C.byref(cfg), C.byref(h)))
        self.h = h
        n, od = self.num_envs, self.obs_dim
        dev = self.device
        self.obs = torch.zeros((n, od), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.terminal_obs = torch.zeros((n, od), dtype=torch.float32, device=dev)
        self.events = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._host = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "h", None):
            self.lib.tb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ parameters
    def set_param(self, name, value):
        _lib.check(self.lib.tb_set_param(self.h, name.encode(), float(value)))

    def set_control_mode(self, mode):
        """'force' (both gym envs) or 'pid' (Racket.apply_action: the action is a target position)."""
        m = {"force": _lib.CONTROL_FORCE, "pid": _lib.CONTROL_PID}.get(mode, mode)
        _lib.check(self.lib.tb_set_control_mode(self.h, int(m)))

    def get_param(self, name):
        v = C.c_double()
        _lib.check(self.lib.tb_get_param(self.h, name.encode(), C.byref(v)))
        return v.value

    # ------------------------------------------------------------------ device-resident API (zero copy)
    def reset(self, mask=None, init=None):
        """New episode for masked envs (all if None). `init`: [N, 8] explicit placement. Returns obs (device)."""
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        if init is None:
            _lib.check(self.lib.tb_reset(self.h, _dptr(m), _dptr(self.obs), self._stream()))
        else:
            i = torch.as_tensor(init, dtype=torch.float64, device=self.device).reshape(self.num_envs, _lib.INIT_WORDS).contiguous()
            _lib.check(self.lib.tb_reset_from(self.h, _dptr(i), _dptr(m), _dptr(self.obs), self._stream()))
        return self.obs

    def step(self, actions, obs=None, reward=None, done=None, terminal_obs=None, events=None):
        """One env step for all N envs (step_kernel + ff_kernel on the current stream). actions: float32 CUDA tensor [N, act_dim].
        Returns (obs, reward, done, terminal_obs, events) device tensors (the caller's, or reused internal ones)."""
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.numel() != self.num_envs * self.act_dim:
            raise ValueError(f"actions must have shape ({self.num_envs}, {self.act_dim})")
        obs = self.obs if obs is None else obs
        reward = self.reward if reward is None else reward
        done = self.done if done is None else done
        terminal_obs = self.terminal_obs if terminal_obs is None else terminal_obs
        events = self.events if events is None else events
        _lib.check(self.lib.tb_step(self.h, _dptr(actions), _dptr(obs), _dptr(reward), _dptr(done), _dptr(terminal_obs),
                                    _dptr(events), self._stream()))
        return obs, reward, done, terminal_obs, events

    def rollout(self, k_steps, action_mode=_lib.ACT_RANDOM, want_outputs=True):
        """K env steps fused in one launch with in-kernel actions. Returns (obs, reward_sum, done_count)."""
        if want_outputs:
            if not hasattr(self, "_rsum"):
                self._rsum = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
                self._dcount = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            args = (_dptr(self.obs), _dptr(self._rsum), _dptr(self._dcount))
        else:
            args = (None, None, None)
        _lib.check(self.lib.tb_rollout(self.h, int(action_mode), int(k_steps), *args, self._stream()))
        return (self.obs, self._rsum, self._dcount) if want_outputs else None

    # ------------------------------------------------------------------ policy rollout (tb_set_policy / tb_policy_rollout)
    def set_policy(self, params):
        """params: float32 CUDA tensor [POLICY_FLOATS] in the layout of include/tennisbot_b200.h (pack_policy builds it)."""
        p = params.to(device=self.device, dtype=torch.float32).contiguous()
        if p.numel() != _lib.POLICY_FLOATS:
            raise ValueError(f"policy parameter vector must have {_lib.POLICY_FLOATS} floats")
        _lib.check(self.lib.tb_set_policy(self.h, _dptr(p), p.numel(), self._stream()))

    def policy_rollout(self, obs, reward, done, last_obs, actions=None, logp=None, value=None, last_value=None,
                       deterministic=False, noise_seed=0):
        """K = obs.shape[0] env steps with actions sampled in-kernel from the policy set by set_policy.  obs[0] must hold
        the current observation.  All buffers are float32 CUDA tensors ([K, N, 6] / [K, N]; done uint8)."""
        k = int(obs.shape[0])
        _lib.check(self.lib.tb_policy_rollout(self.h, k, int(bool(deterministic)), int(noise_seed) & (2 ** 64 - 1), _dptr(obs),
                                              _dptr(actions), _dptr(logp), _dptr(value), _dptr(reward), _dptr(done), _dptr(last_obs),
                                              _dptr(last_value), self._stream()))

    def get_state(self):
        s = torch.empty((self.num_envs, _lib.STATE_WORDS), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.tb_get_state(self.h, _dptr(s), self._stream()))
        return s

    def set_state(self, state):
        s = torch.as_tensor(state, dtype=torch.float64, device=self.device).reshape(self.num_envs, _lib.STATE_WORDS).contiguous()
        _lib.check(self.lib.tb_set_state(self.h, _dptr(s), self._stream()))

    def stats_tensor(self):
        """int64[10] statistics vector living in HBM, as a torch view (all-reduce it over NCCL)."""
        p = C.c_void_p()
        _lib.check(self.lib.tb_stats_device_ptr(self.h, C.byref(p)))
        return _wrap_device_int64(p.value, _lib.NUM_STATS, self.device, owner=self)

    def read_stats(self, clear=False):
        out = np.zeros(_lib.NUM_STATS, np.int64)
        _lib.check(self.lib.tb_read_stats(self.h, _hptr(out), int(clear), self._stream()))
        return out

    def launch_count(self):
        v = C.c_int64()
        _lib.check(self.lib.tb_launch_count(self.h, C.byref(v)))
        return v.value

    def ff_diagnostics(self):
        """int64[16] record of the most recent fast-forward launch (see tb_ff_diagnostics)."""
        import numpy as np

        out = np.zeros(16, np.int64)
        _lib.check(self.lib.tb_ff_diagnostics(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def set_kernel_timing(self, enabled):
        _lib.check(self.lib.tb_set_kernel_timing(self.h, int(bool(enabled))))

    def kernel_timing(self):
        """(ms in step_kernel, ms in ff_kernel, steps covered) accumulated since the last call."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(self.lib.tb_get_kernel_timing(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    # ------------------------------------------------------------------ host-buffer API (pinned numpy in/out)
    def host_buffers(self):
        """Pinned host arrays reused by reset_host/step_host (views of torch pinned tensors)."""
        if self._host is None:
            n, od, ad = self.num_envs, self.obs_dim, self.act_dim

            def pin(shape, dt):
                return torch.zeros(shape, dtype=dt).pin_memory()

            t = dict(actions=pin((n, ad), torch.float32), obs=pin((n, od), torch.float32), reward=pin((n,), torch.float32),
                     done=pin((n,), torch.uint8), terminal_obs=pin((n, od), torch.float32), events=pin((n,), torch.uint8))
            self._host_t = t
            self._host = {k: v.numpy() for k, v in t.items()}
        return self._host

    def reset_host(self, mask=None):
        hb = self.host_buffers()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        _lib.check(self.lib.tb_reset_host(self.h, _hptr(m), _hptr(hb["obs"])))
        return hb["obs"]

    def step_host(self, actions=None, want_terminal=True, want_events=True):
        """actions: float32 [N, act_dim] host array (None = already written into host_buffers()['actions']).
        One env step inside the C call, synchronised on return.  The buffers are pinned, so by default the kernels read the
        actions from and write the results to host memory themselves (zero copy; TB_HOST_MODE selects the sliced
        copy-engine pipeline or plain staging instead).  Returns the pinned result arrays (overwritten by the next call)."""
        hb = self.host_buffers()
        if actions is not None:
            np.copyto(hb["actions"], np.asarray(actions, np.float32).reshape(self.num_envs, self.act_dim))
        _lib.check(self.lib.tb_step_host(self.h, _hptr(hb["actions"]), _hptr(hb["obs"]), _hptr(hb["reward"]),
                                         _hptr(hb["done"]), _hptr(hb["terminal_obs"]) if want_terminal else None,
                                         _hptr(hb["events"]) if want_events else None))
        return hb


def _wrap_device_int64(ptr, n, device, owner):
    """View `n` int64 at raw device address `ptr` as a torch tensor through __cuda_array_interface__."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}
    h._owner = owner
    return torch.as_tensor(h, device=device)
