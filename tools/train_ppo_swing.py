#!/usr/bin/env python3
"""CLI of tennisbot_rl_b200.ppo (GPU-resident PPO on SwingRacket-v0, BASELINE config 4).

    python tools/train_ppo_swing.py --iters 150 --out gpurun_out/ppo_swing_b200.json
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tennisbot_rl_b200.ppo import SwingPPO  # noqa: E402


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
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--eager-policy", action="store_true", help="torch policy in the rollout instead of tb_policy_rollout")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ppo = SwingPPO(args.envs, args.precision, args.seed, args.lr, args.epochs, args.minibatches, use_graph=not args.no_graph,
                   fused_policy=not args.eager_policy)
    summary = ppo.train(args.iters, args.target, log=lambda s: print(s, flush=True))
    print(json.dumps({k: v for k, v in summary.items() if k != "history"}))
    if args.out:
        Path(args.out).write_text(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
