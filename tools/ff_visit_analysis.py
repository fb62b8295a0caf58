"""Which envs of a fast-forward launch go through the servers, and which work-list rule predicts it (the "last-to-start" class of
step_kernel's lists).  Needs a library built with -DTB_FF_DIAG_VISITS (bit 7 of the events byte = the flight was
parked with the servers at least once):

    python tools/build_variants.py diagv:-DTB_FF_DIAG_VISITS
    TB_LIB_PATH=build/variants/lib_diagv.so python tools/ff_visit_analysis.py [n_envs]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from tennisbot_rl_b200.batch import TennisBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
b = TennisBatch("SwingRacket-v0", n, seed=0, precision="f64", auto_reset=False)
b.reset()
g = torch.Generator(device="cuda")
g.manual_seed(1)
for t in range(25):
    b.step(torch.empty((n, 6), device="cuda").uniform_(-1, 1, generator=g))
s = b.get_state().cpu().numpy()
ev = b.step(torch.empty((n, 6), device="cuda").uniform_(-1, 1, generator=g))[4].cpu().numpy()
vis = (ev & 128) != 0
if not vis.any():
    sys.exit("no visit bits: load a -DTB_FF_DIAG_VISITS build through TB_LIB_PATH")
rp, rq, rv, bp, bv = s[:, 0:3], s[:, 3:7], s[:, 7:10], s[:, 13:16], s[:, 16:19]
x, y, z, w = rq.T
nrm = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w)], 1)  # the racket face's normal
d = (nrm * (bp - rp)).sum(1)
vn = (nrm * (bv - rv)).sum(1)
closing = (d * vn < 0) & (np.abs(d) < 0.6 * np.abs(vn))
front = ((bv ** 2).sum(1) > 9) | closing
print(f"{n} envs: {vis.mean():.3f} visit the servers; front list {front.mean():.3f} of the envs ({vis[front].mean():.3f} of them visit), "
      f"the others {vis[~front].mean():.4f}")
print("free-falling balls that move away from the racket's plane at >= V m/s: share of all envs, visitors among them")
for V in (0.0, 0.2, 0.4, 0.5, 0.6, 0.8):
    safe = ~front & (d * vn > 0) & (np.abs(vn) >= V)
    print(f"  V = {V:.1f}: {safe.mean():.3f}  {int((safe & vis).sum())}")
