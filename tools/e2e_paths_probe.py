"""Where a zero-copy env step spends its time: tb_step with the actions and / or the results in pinned host memory (the kernels
address it over PCIe) against the same step on device buffers.  Light steps only (steps 1..20 of an episode).
usage: e2e_paths_probe.py [n_envs] [env]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from tennisbot_rl_b200 import _lib
from tennisbot_rl_b200.batch import TennisBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = sys.argv[2] if len(sys.argv) > 2 else "SwingRacket-v0"
b = TennisBatch(env, n, seed=0)
od, ad = b.obs_dim, b.act_dim
dev = b.device


def bufs(host):
    """host: bool, or a string of the arrays that live in host memory ('o', 'r', 'd')"""
    host = "ord" if host is True else (host or "")
    kw = lambda c: dict(pin_memory=True) if c in host else dict(device=dev)  # noqa: E731
    return dict(obs=torch.zeros((n, od), dtype=torch.float32, **kw("o")), reward=torch.zeros(n, dtype=torch.float32, **kw("r")),
                done=torch.zeros(n, dtype=torch.uint8, **kw("d")))


def acts(host):
    a = torch.empty((n, ad), dtype=torch.float32).uniform_(-1, 1)
    return a.pin_memory() if host else a.to(dev)


p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
for name, ah, oh in (("device in, device out", False, False), ("host in, device out", True, False), ("device in, host out", False, True),
                     ("host in, host out", True, True), ("host in, obs to host", True, "o"), ("host in, obs+reward to host", True, "or"),
                     ("device in, done to host", False, "d"), ("device in, reward to host", False, "r"), ("device in, obs to host", False, "o")):
    a, o = acts(ah), bufs(oh)
    ts = []
    for rep in range(3):
        b.reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(20):
            _lib.check(b.lib.tb_step(b.h, p(a), p(o["obs"]), p(o["reward"]), p(o["done"]), None, None, b._stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    mb_in = n * ad * 4 / 1e6
    mb_out = n * (od * 4 + 5) / 1e6 if oh is True else sum(n * {"o": od * 4, "r": 4, "d": 1}[ch] for ch in (oh or "")) / 1e6
    t = min(ts)
    print(f"{name:24s} {t:.4f} ms/step   in {mb_in:.1f} MB, out {mb_out:.1f} MB" +
          (f"   -> {((mb_in if ah else 0) + (mb_out if oh else 0)) / t:.1f} GB/s over the link" if ah or oh else ""), flush=True)
