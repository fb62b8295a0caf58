"""Step-by-step parity against the oracle at the bench's own batch size (1 048 576 SwingRacket-v0 envs per GPU), a one-off that is
too slow for the test suite's budget: every event byte and done flag, the statistics, and the state / obs / reward bars of
tests/test_parity_gpu.py.  usage: full_size_parity.py [n_envs] [episodes]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import binding
from tests.harness import reference_reset_params, run_parity
from tennisbot_rl_b200.batch import TennisBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
episodes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
binding.build()
t0 = time.time()
b = TennisBatch("SwingRacket-v0", n, seed=101, precision="f64")
o = binding.OracleEnv("SwingRacket-v0", n, seed=101, threads=16)
rng = np.random.default_rng(77)
init = reference_reset_params(o.kind, n, rng)
np.testing.assert_array_equal(b.reset(init=init).cpu().numpy(), o.reset(init=init))
rep, valid = run_parity(b, o, 26 * episodes, lambda t, _obs: rng.uniform(-1, 1, (n, o.act_dim)), band=0.0, check_state_every=13)
np.testing.assert_array_equal(b.read_stats(), o.read_stats())
assert rep.event_mismatch_hard == 0 and rep.event_mismatch_near == 0 and rep.dropped == 0
assert rep.max_state_err < 5e-8 and rep.max_obs_err < 2e-6 and rep.max_reward_err < 2e-6
print(f"{n} envs x {episodes} episodes ({n * episodes} episodes, {n * 26 * episodes} env steps): {rep}; statistics equal {b.read_stats().tolist()}; "
      f"{time.time() - t0:.0f} s")
