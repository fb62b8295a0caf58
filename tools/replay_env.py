"""Replay single envs of a tools/full_size_parity.py run (same seeds, placements and action streams) on the GPU and on the oracle and
report the first step at which their states differ by more than `tol`, with both states before and after that step.
usage: replay_env.py <env name> <n_envs of the original run> <steps> <tol> <env index> [<env index> ...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import binding
from tests.harness import reference_reset_params
from tennisbot_rl_b200.batch import TennisBatch

env, n, steps, tol = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
ids = [int(x) for x in sys.argv[5:]]
binding.build()
probe = binding.OracleEnv(env, 1, seed=101)
rng = np.random.default_rng(77)
init = reference_reset_params(probe.kind, n, rng)
acts = np.empty((steps, len(ids), probe.act_dim), np.float32)
for t in range(steps):
    a = rng.uniform(-1, 1, (n, probe.act_dim)).astype(np.float32)
    acts[t] = a[ids]
np.set_printoptions(precision=17, linewidth=200)
for k, i in enumerate(ids):
    b = TennisBatch(env, 1, seed=101, precision="f64", env_id_offset=i)
    o = binding.OracleEnv(env, 1, seed=101, env_id_offset=i)
    np.testing.assert_array_equal(b.reset(init=init[i:i + 1]).cpu().numpy(), o.reset(init=init[i:i + 1]))
    prev_g, prev_o = b.get_state().cpu().numpy(), o.get_state().copy()
    for t in range(steps):
        a = acts[t, k:k + 1]
        g = [x.cpu().numpy() for x in b.step(torch.from_numpy(a).to(b.device))]
        r = o.step(a, want_margin=True)
        gs, os_ = b.get_state().cpu().numpy(), o.get_state().copy()
        d = np.abs(gs - os_).max()
        if d > tol or g[4][0] != r["events"][0] or g[2][0] != r["done"][0]:
            print(f"env {i}: first difference at step {t}: max state diff {d:.3e} (entry {int(np.abs(gs - os_).argmax())}), events gpu {int(g[4][0])} oracle "
                  f"{int(r['events'][0])}, done {int(g[2][0])}/{int(r['done'][0])}, oracle margin {float(r['margin'][0]):.6g}, action {a[0].tolist()}")
            print("  state before the step (gpu)   ", prev_g[0].tolist())
            print("  state before the step (oracle)", prev_o[0].tolist())
            print("  state after (gpu)   ", gs[0].tolist())
            print("  state after (oracle)", os_[0].tolist())
            np.savez(f"gpurun_out/replay_env_{i}.npz", before_gpu=prev_g, before_oracle=prev_o, after_gpu=gs, after_oracle=os_, action=a, step=t)
            break
        prev_g, prev_o = gs, os_
    else:
        print(f"env {i}: no difference above {tol} in {steps} steps")
