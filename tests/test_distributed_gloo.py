"""N > 1 path on CPU: two gloo ranks each step their contiguous shard (global-id-keyed RNG streams) and all-reduce the
int64 statistics vector; the result equals the single-process full batch.  The shards are stepped by the oracle
here (no GPU in this container); the GPU version of the same check is tests/test_api_gpu.py::test_shard_invariance."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
TOTAL, STEPS, SEED = 96, 60, 13


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    from tennisbot_rl_b200.sharding import all_reduce_stats, shard_range

    lo, hi = shard_range(TOTAL, rank, world)
    env = ob.OracleEnv("SwingRacket-v0", hi - lo, env_id_offset=lo, seed=SEED)
    env.reset()
    res = env.rollout(STEPS)
    stats = torch.from_numpy(env.read_stats().copy())
    all_reduce_stats(stats)
    np.save(Path(out_dir) / f"obs{rank}.npy", res["obs"])
    if rank == 0:
        np.save(Path(out_dir) / "stats.npy", stats.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_batch(tmp_path, oracle_lib):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = oracle_lib.OracleEnv("SwingRacket-v0", TOTAL, seed=SEED)
    full.reset()
    ref = full.rollout(STEPS)
    obs = np.concatenate([np.load(tmp_path / f"obs{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(obs, ref["obs"])                      # shard union == full batch, bit for bit
    np.testing.assert_array_equal(np.load(tmp_path / "stats.npy"), full.read_stats())  # integer sums: order independent
