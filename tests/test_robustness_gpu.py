"""GPU tests of ff_kernel's behaviour next to other work on the device and of the fault word: the persistent
fast-forward kernel has no grid barrier and claims all its work from counters and queues, so it must complete - with results
equal to the oracle's - whatever else occupies SMs while it runs; and a launch that gives up on a wait must make the next call
on the context fail loudly."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _spin_kernel_stream(dev, seconds):
    """Keep every SM busy on a side stream: a long chain of large matmuls (torch kernels, not ours)."""
    s = torch.cuda.Stream(dev)
    a = torch.randn((8192, 8192), device=dev, dtype=torch.float32)
    with torch.cuda.stream(s):
        c = a
        for _ in range(int(seconds * 40)):
            c = (c @ a).mul_(1e-4)
    return s, c


def test_env_steps_complete_next_to_a_busy_stream(oracle_lib):
    from tennisbot_rl_b200.batch import TennisBatch

    n = 8192
    dev = torch.device("cuda", 0)
    b = TennisBatch("SwingRacket-v0", n, seed=41)
    o = oracle_lib.OracleEnv("SwingRacket-v0", n, seed=41, threads=8)
    np.testing.assert_array_equal(b.reset().cpu().numpy(), o.reset())
    rng = np.random.default_rng(3)
    acts = rng.uniform(-1, 1, (52, n, 6)).astype(np.float32)
    dacts = torch.from_numpy(acts).to(dev)
    side, keep = _spin_kernel_stream(dev, 1.0)  # ~1 s of matmuls in flight while two whole episodes are stepped
    outs = []
    for t in range(52):
        obs, rew, done, _, ev = b.step(dacts[t])
        outs.append((obs.clone(), rew.clone(), done.clone(), ev.clone()))
    busy_during = not side.query()
    torch.cuda.synchronize()
    for t in range(52):
        ref = o.step(acts[t])
        g = [x.cpu().numpy() for x in outs[t]]
        np.testing.assert_array_equal(g[2], ref["done"])
        np.testing.assert_array_equal(g[3], ref["events"])
        np.testing.assert_allclose(g[0], ref["obs"], atol=2e-6)
        np.testing.assert_allclose(g[1], ref["reward"], atol=2e-6)
    np.testing.assert_array_equal(b.read_stats(), o.read_stats())
    assert busy_during, "the side stream finished before the env steps were issued: the test did not overlap anything"
    assert b.ff_diagnostics()[15] == 0
    b.close()


def test_env_steps_next_to_an_nccl_all_reduce(oracle_lib):
    """Single-rank NCCL group: the collective's kernel holds SM slots on its own stream while the env steps."""
    import torch.distributed as dist

    from tennisbot_rl_b200.batch import TennisBatch

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        n = 8192
        dev = torch.device("cuda", 0)
        b = TennisBatch("SwingRacket-v0", n, seed=43)
        o = oracle_lib.OracleEnv("SwingRacket-v0", n, seed=43, threads=8)
        np.testing.assert_array_equal(b.reset().cpu().numpy(), o.reset())
        rng = np.random.default_rng(4)
        acts = rng.uniform(-1, 1, (26, n, 6)).astype(np.float32)
        dacts = torch.from_numpy(acts).to(dev)
        big = torch.ones(64 << 20, device=dev)
        works = []
        outs = []
        for t in range(26):
            works.append(dist.all_reduce(big, async_op=True))
            obs, rew, done, _, ev = b.step(dacts[t])
            outs.append((obs.clone(), rew.clone(), done.clone(), ev.clone()))
        for w in works:
            w.wait()
        torch.cuda.synchronize()
        for t in range(26):
            ref = o.step(acts[t])
            g = [x.cpu().numpy() for x in outs[t]]
            np.testing.assert_array_equal(g[2], ref["done"])
            np.testing.assert_array_equal(g[3], ref["events"])
            np.testing.assert_allclose(g[0], ref["obs"], atol=2e-6)
        np.testing.assert_array_equal(b.read_stats(), o.read_stats())
        b.close()
    finally:
        if created:
            dist.destroy_process_group()


def test_a_timed_out_fast_forward_fails_the_next_call():
    """TB_FF_SPIN_LIMIT_MS far below what any wait inside ff_kernel needs: the launch gives up, and the context refuses
    every later call instead of handing stale observations on (tb_step / tb_step_host / tb_reset / tb_read_stats)."""
    from tennisbot_rl_b200 import _lib
    from tennisbot_rl_b200.batch import TennisBatch

    old = os.environ.get("TB_FF_SPIN_LIMIT_MS")
    os.environ["TB_FF_SPIN_LIMIT_MS"] = "0.0005"  # ~1000 clock cycles
    try:
        b = TennisBatch("SwingRacket-v0", 65536, seed=1)
    finally:
        if old is None:
            del os.environ["TB_FF_SPIN_LIMIT_MS"]
        else:
            os.environ["TB_FF_SPIN_LIMIT_MS"] = old
    b.reset()
    a = torch.zeros((65536, 6), device="cuda")
    with pytest.raises(_lib.TennisbotLibraryError, match="timed out"):
        for t in range(30):  # the 26th step runs the fast-forward; the call after it sees the fault word
            b.step(a)
            torch.cuda.synchronize()
    with pytest.raises(_lib.TennisbotLibraryError, match="timed out"):
        b.reset()
    with pytest.raises(_lib.TennisbotLibraryError):
        b.read_stats()
    b.close()
