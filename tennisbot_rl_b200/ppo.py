"""GPU-resident PPO on SwingRacket-v0 (BASELINE config 4: 16 384 envs on one B200) - SURVEY 8(f)-1.

The reference trains with SB3 PPO (`train_swing.py:80-91`: MlpPolicy, net_arch pi/vf [32,64,32], tanh, ent_coef
0.002, gamma .99, gae_lambda .95, clip .2, n_epochs 10, vf_coef .5, max_grad_norm .5) on ONE env, 1100 steps per
update.  SB3 is not installed here, so this is a plain-torch PPO with the same policy architecture and
hyper-parameters; only the batch geometry changes (N envs x 26 steps = one whole episode per env per update) and
the learning rate is the reference's 3e-4 default.  Observations and actions never leave HBM, and the rollout runs no
torch code at all: tb_policy_rollout evaluates the policy in a CUDA kernel of the library (forward, Philox Gaussian noise,
log-density, value), so the 26-step episode is 26 x (policy_kernel, step_kernel, ff_kernel) issued by one C call; the call
is captured once as a CUDA graph and replayed per update (nothing in it synchronises; the queue tags of ff_kernel advance
on the device).  `fused_policy=False` keeps the eager torch policy of round 1 for comparison.  Target: the mean episodic
return stored in backup_models/ppo_swing.zip, 31.5.
"""
import math
import time

import torch
import torch.nn as nn

from . import _lib
from .batch import TennisBatch

EPISODE = 26  # agent steps per SwingRacket-v0 episode (swingracket_env.py:85-86: 25 control steps + the fast-forward step)


class ActorCritic(nn.Module):
    def __init__(self, obs_dim=6, act_dim=6):
        super().__init__()

        def mlp():
            return nn.Sequential(nn.Linear(obs_dim, 32), nn.Tanh(), nn.Linear(32, 64), nn.Tanh(), nn.Linear(64, 32), nn.Tanh())

        self.pi, self.vf = mlp(), mlp()
        self.mu = nn.Linear(32, act_dim)
        self.v = nn.Linear(32, 1)
        self.log_std = nn.Parameter(torch.zeros(act_dim))
        for m in list(self.pi) + list(self.vf):
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, 2 ** 0.5)
                nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.mu.weight, 0.01)
        nn.init.orthogonal_(self.v.weight, 1.0)

    def mean(self, obs):
        return self.mu(self.pi(obs))

    def log_prob(self, obs, act):
        mu, ls = self.mean(obs), self.log_std
        return (-0.5 * ((act - mu) / ls.exp()) ** 2 - ls - 0.5 * math.log(2 * math.pi)).sum(-1)

    def entropy(self):
        return (0.5 + 0.5 * math.log(2 * math.pi) + self.log_std).sum()

    def value(self, obs):
        return self.v(self.vf(obs)).squeeze(-1)

    def packed(self):
        lin = lambda seq: [(m.weight, m.bias) for m in seq if isinstance(m, nn.Linear)]  # noqa: E731
        return pack_policy(lin(self.pi), (self.mu.weight, self.mu.bias), lin(self.vf), (self.v.weight, self.v.bias), self.log_std)


def pack_policy(pi, mu, vf, v, log_std):
    """Flat float32 parameter vector in tb_set_policy's layout from (weight, bias) lists of the two towers: pi, vf = three
    (W [out, in], b) pairs each; mu, v = the heads; every tensor padded to a multiple of 4 floats."""
    def pad(t):
        t = t.detach().reshape(-1).to(torch.float32)
        r = (-t.numel()) % 4
        return t if r == 0 else torch.cat([t, t.new_zeros(r)])
    parts = []
    for tower, head in ((pi, mu), (vf, v)):
        for W, b in tower:
            parts += [pad(W), pad(b)]
        parts += [pad(head[0]), pad(head[1])]
    parts.append(pad(log_std))
    out = torch.cat(parts)
    assert out.numel() == _lib.POLICY_FLOATS, out.numel()
    return out


class SwingPPO:
    """PPO learner whose rollouts run on the B200 env batch.  use_graph: replay the rollout as one CUDA graph."""

    def __init__(self, num_envs=16384, precision="f64", seed=0, lr=3e-4, epochs=10, minibatches=8, device=0, use_graph=True,
                 fused_policy=True):
        torch.manual_seed(seed)
        self.fused_policy = fused_policy
        self.seed = seed
        self.dev = torch.device("cuda", device)
        self.n = n = int(num_envs)
        self.env = TennisBatch("SwingRacket-v0", n, device=device, seed=seed, precision=precision)
        self.ac = ActorCritic().to(self.dev)
        self.opt = torch.optim.Adam(self.ac.parameters(), lr=lr, eps=1e-5, capturable=use_graph)
        self.epochs, self.minibatches = epochs, minibatches
        self.gamma, self.lam, self.clip, self.ent_coef, self.vf_coef, self.max_norm = 0.99, 0.95, 0.2, 0.002, 0.5, 0.5
        z = lambda *s: torch.zeros(s, device=self.dev)  # noqa: E731
        self.obs_buf, self.act_buf = z(EPISODE, n, 6), z(EPISODE, n, 6)
        self.logp_buf, self.rew_buf, self.done_buf, self.val_buf = z(EPISODE, n), z(EPISODE, n), z(EPISODE, n), z(EPISODE + 1, n)
        self.done_u8 = torch.zeros((EPISODE, n), dtype=torch.uint8, device=self.dev)
        self.params = z(_lib.POLICY_FLOATS)
        self.obs = self.env.reset().clone()
        self.graph = None
        self.use_graph = use_graph
        self.rollout_s = 0.0
        # one minibatch update as a CUDA graph over static buffers (the update is ~40 small kernels, 80 times per iteration)
        self.mb = EPISODE * n // minibatches
        self.upd_graph = None
        self.mb_obs, self.mb_act = z(self.mb, 6), z(self.mb, 6)
        self.mb_logp, self.mb_adv, self.mb_ret = z(self.mb), z(self.mb), z(self.mb)

    # ------------------------------------------------------------------ rollout
    def _rollout_body(self):
        ac, env = self.ac, self.env
        if self.fused_policy:
            # parameters -> the library's buffer (device to device, inside the graph: every replay sees the current weights)
            self.params.copy_(ac.packed())
            env.set_policy(self.params)
            self.obs_buf[0].copy_(self.obs)
            env.policy_rollout(self.obs_buf, self.rew_buf, self.done_u8, self.obs, actions=self.act_buf, logp=self.logp_buf,
                               value=self.val_buf[:EPISODE], last_value=self.val_buf[EPISODE], noise_seed=self.seed + 1)
            self.done_buf.copy_(self.done_u8)
            return
        for t in range(EPISODE):
            mu = ac.mean(self.obs)
            a = mu + ac.log_std.exp() * torch.randn_like(mu)
            self.obs_buf[t].copy_(self.obs)
            self.act_buf[t].copy_(a)
            self.logp_buf[t].copy_(ac.log_prob(self.obs, a))
            self.val_buf[t].copy_(ac.value(self.obs))
            o, r, dn, _, _ = env.step(a.clamp(-1, 1))  # SB3 clips to the Box before env.step
            self.rew_buf[t].copy_(r)
            self.done_buf[t].copy_(dn)
            self.obs.copy_(o)
        self.val_buf[EPISODE].copy_(ac.value(self.obs))

    @torch.no_grad()
    def rollout(self):
        """One whole episode per env into the buffers; returns the statistics vector of those episodes."""
        self.env.read_stats(clear=True)
        torch.cuda.synchronize()
        t0 = time.time()
        if not self.use_graph:
            self._rollout_body()
        else:
            if self.graph is None:  # warm up on a side stream (library state, cuBLAS handles), then capture
                s = torch.cuda.Stream(self.dev)
                s.wait_stream(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(s):
                    self._rollout_body()
                torch.cuda.current_stream(self.dev).wait_stream(s)
                torch.cuda.synchronize()
                self.env.read_stats(clear=True)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._rollout_body()
                torch.cuda.synchronize()
                t0 = time.time()
            self.graph.replay()
        torch.cuda.synchronize()
        self.rollout_s += time.time() - t0
        return self.env.read_stats()

    # ------------------------------------------------------------------ update
    def update(self):
        n, ac = self.n, self.ac
        adv = torch.zeros_like(self.rew_buf)
        last = torch.zeros(n, device=self.dev)
        for t in reversed(range(EPISODE)):
            nonterm = 1.0 - self.done_buf[t]
            delta = self.rew_buf[t] + self.gamma * self.val_buf[t + 1] * nonterm - self.val_buf[t]
            last = delta + self.gamma * self.lam * nonterm * last
            adv[t] = last
        ret = adv + self.val_buf[:EPISODE]
        B = EPISODE * n
        fo, fa, fl = self.obs_buf.reshape(B, 6), self.act_buf.reshape(B, 6), self.logp_buf.reshape(B)
        fadv, fret = adv.reshape(B), ret.reshape(B)
        mb = self.mb
        for _ in range(self.epochs):
            perm = torch.randperm(B, device=self.dev)
            for k in range(self.minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                torch.index_select(fo, 0, idx, out=self.mb_obs)
                torch.index_select(fa, 0, idx, out=self.mb_act)
                torch.index_select(fl, 0, idx, out=self.mb_logp)
                torch.index_select(fadv, 0, idx, out=self.mb_adv)
                torch.index_select(fret, 0, idx, out=self.mb_ret)
                if not self.use_graph:
                    self._minibatch_step()
                    continue
                if self.upd_graph is None:
                    # warm-up steps are real Adam updates: take them on a snapshot and put model and optimiser back, so
                    # that the graph path trains exactly like the eager one (capture itself executes nothing)
                    snap_m = {k: v.clone() for k, v in self.ac.state_dict().items()}
                    s = torch.cuda.Stream(self.dev)
                    s.wait_stream(torch.cuda.current_stream(self.dev))
                    with torch.cuda.stream(s):
                        for _w in range(3):
                            self._minibatch_step()
                    torch.cuda.current_stream(self.dev).wait_stream(s)
                    self.upd_graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.upd_graph):
                        self._minibatch_step()
                    with torch.no_grad():  # back to "no update taken yet": weights, Adam moments and step counts
                        for k, v in self.ac.state_dict().items():
                            v.copy_(snap_m[k])
                        for st in self.opt.state.values():
                            for v in st.values():
                                if torch.is_tensor(v):
                                    v.zero_()
                self.upd_graph.replay()

    def _minibatch_step(self):
        ac = self.ac
        ratio = (ac.log_prob(self.mb_obs, self.mb_act) - self.mb_logp).exp()
        a_ = self.mb_adv
        a_ = (a_ - a_.mean()) / (a_.std() + 1e-8)
        pg = -torch.min(ratio * a_, ratio.clamp(1 - self.clip, 1 + self.clip) * a_).mean()
        vloss = 0.5 * (ac.value(self.mb_obs) - self.mb_ret).pow(2).mean()
        loss = pg + self.vf_coef * vloss - self.ent_coef * ac.entropy()
        self.opt.zero_grad(set_to_none=False)  # grads stay in place: the captured graphs refer to their storage
        loss.backward()
        nn.utils.clip_grad_norm_(ac.parameters(), self.max_norm)
        self.opt.step()

    def train(self, iters=150, target=31.5, log=None):
        history, reached, t0 = [], None, time.time()
        for it in range(iters):
            st = self.rollout()
            eps = max(int(st[0]), 1)
            h = {"iter": it, "env_steps": (it + 1) * EPISODE * self.n, "mean_return": float(st[6] / 2 ** 20 / eps),
                 "goal_fraction": float(st[3] / eps), "hits_per_episode": float(st[2] / eps)}
            history.append(h)
            if reached is None and h["mean_return"] >= target:
                reached = h
            self.update()
            if log and (it % 10 == 0 or it == iters - 1):
                log(f"iter {it:4d} env-steps {h['env_steps']:>10d} return {h['mean_return']:7.2f} goals {h['goal_fraction']:5.3f} "
                    f"hits/ep {h['hits_per_episode']:.2f} wall {time.time() - t0:6.1f}s")
        wall = time.time() - t0
        return {"envs": self.n, "iters": iters, "cuda_graph_rollout": bool(self.use_graph),
                "final_mean_return": history[-1]["mean_return"], "best_mean_return": max(h["mean_return"] for h in history),
                "reference_target": target, "first_reached": reached, "wall_s": wall, "rollout_s": self.rollout_s,
                "rollout_env_steps_per_s": iters * EPISODE * self.n / max(self.rollout_s, 1e-9), "history": history}
