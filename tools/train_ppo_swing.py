#!/usr/bin/env python3
"""GPU-resident PPO on SwingRacket-v0 (BASELINE config 4: 16 384 envs on one B200) - SURVEY 8(f)-1.

The reference trains with SB3 PPO (`train_swing.py:80-91`: MlpPolicy, net_arch pi/vf [32,64,32], tanh, ent_coef
0.002, gamma .99, gae_lambda .95, clip .2, n_epochs 10, vf_coef .5, max_grad_norm .5) on ONE env, 1100 steps per
update.  SB3 is not installed here, so this is a plain-torch PPO with the same policy architecture and
hyper-parameters; only the batch geometry changes (16 384 envs x 26 steps = one whole episode per env per update)
and the learning rate is the reference's 3e-4 default.  Observations and actions never leave HBM: the rollout calls
TennisBatch.step with CUDA tensors.  Target: the mean episodic return stored in backup_models/ppo_swing.zip, 31.5.

    python tools/train_ppo_swing.py --iters 150 --out gpurun_out/ppo_swing_b200.json
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tennisbot_rl_b200.batch import TennisBatch  # noqa: E402

EPISODE = 26


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

    def dist(self, obs):
        return torch.distributions.Normal(self.mu(self.pi(obs)), self.log_std.exp())

    def value(self, obs):
        return self.v(self.vf(obs)).squeeze(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=150)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--minibatches", type=int, default=8)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--target", type=float, default=31.5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    torch.manual_seed(args.seed)
    dev = torch.device("cuda", 0)
    n = args.envs
    env = TennisBatch("SwingRacket-v0", n, device=0, seed=args.seed, precision=args.precision)
    ac = ActorCritic().to(dev)
    opt = torch.optim.Adam(ac.parameters(), lr=args.lr, eps=1e-5)
    gamma, lam, clip, ent_coef, vf_coef, max_norm = 0.99, 0.95, 0.2, 0.002, 0.5, 0.5

    obs_buf = torch.zeros((EPISODE, n, 6), device=dev)
    act_buf = torch.zeros((EPISODE, n, 6), device=dev)
    logp_buf = torch.zeros((EPISODE, n), device=dev)
    rew_buf = torch.zeros((EPISODE, n), device=dev)
    val_buf = torch.zeros((EPISODE + 1, n), device=dev)
    done_buf = torch.zeros((EPISODE, n), device=dev)

    obs = env.reset().clone()
    history, t_env, t0 = [], 0.0, time.time()
    reached = None
    for it in range(args.iters):
        env.read_stats(clear=True)
        torch.cuda.synchronize()
        te = time.time()
        with torch.no_grad():
            for t in range(EPISODE):
                d = ac.dist(obs)
                a = d.sample()
                obs_buf[t], act_buf[t], logp_buf[t], val_buf[t] = obs, a, d.log_prob(a).sum(-1), ac.value(obs)
                o, r, dn, _, _ = env.step(a.clamp(-1, 1))  # SB3 clips to the Box before env.step
                rew_buf[t], done_buf[t] = r, dn.float()
                obs = o.clone()
            val_buf[EPISODE] = ac.value(obs)
        torch.cuda.synchronize()
        t_env += time.time() - te
        st = env.read_stats()
        mean_ret = st[6] / 2 ** 20 / max(int(st[0]), 1)
        # GAE
        adv = torch.zeros_like(rew_buf)
        last = torch.zeros(n, device=dev)
        for t in reversed(range(EPISODE)):
            nonterm = 1.0 - done_buf[t]
            delta = rew_buf[t] + gamma * val_buf[t + 1] * nonterm - val_buf[t]
            last = delta + gamma * lam * nonterm * last
            adv[t] = last
        ret = adv + val_buf[:EPISODE]
        B = EPISODE * n
        fo, fa, fl, fadv, fret = obs_buf.reshape(B, 6), act_buf.reshape(B, 6), logp_buf.reshape(B), adv.reshape(B), ret.reshape(B)
        mb = B // args.minibatches
        for _ in range(args.epochs):
            perm = torch.randperm(B, device=dev)
            for k in range(args.minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                d = ac.dist(fo[idx])
                logp = d.log_prob(fa[idx]).sum(-1)
                ratio = (logp - fl[idx]).exp()
                a_ = fadv[idx]
                a_ = (a_ - a_.mean()) / (a_.std() + 1e-8)
                pg = -torch.min(ratio * a_, ratio.clamp(1 - clip, 1 + clip) * a_).mean()
                vloss = 0.5 * (ac.value(fo[idx]) - fret[idx]).pow(2).mean()
                loss = pg + vf_coef * vloss - ent_coef * d.entropy().sum(-1).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(ac.parameters(), max_norm)
                opt.step()
        goals = st[3] / max(int(st[0]), 1)
        history.append({"iter": it, "env_steps": (it + 1) * B, "mean_return": float(mean_ret), "goal_fraction": float(goals),
                        "hits_per_episode": float(st[2] / max(int(st[0]), 1))})
        if reached is None and mean_ret >= args.target:
            reached = history[-1]
        if it % 10 == 0 or it == args.iters - 1:
            print(f"iter {it:4d} env-steps {history[-1]['env_steps']:>10d} return {mean_ret:7.2f} goals {goals:5.3f} "
                  f"hits/ep {history[-1]['hits_per_episode']:.2f} wall {time.time() - t0:6.1f}s", flush=True)
    wall = time.time() - t0
    summary = {"envs": n, "iters": args.iters, "precision": args.precision, "final_mean_return": history[-1]["mean_return"],
               "best_mean_return": max(h["mean_return"] for h in history), "reference_target": args.target,
               "first_reached": reached, "wall_s": wall, "rollout_s": t_env,
               "rollout_env_steps_per_s": args.iters * EPISODE * n / t_env, "history": history}
    print(json.dumps({k: v for k, v in summary.items() if k != "history"}))
    if args.out:
        Path(args.out).write_text(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
