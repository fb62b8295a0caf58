"""One-off large parity run (a checker like the rest of tests/, but not collected by pytest: run `python tests/stress_parity.py` on a B200):
CUDA path vs the CPU oracle on 262 144 SwingRacket envs x 78 steps
and 65 536 Tennisbot envs x 1100 steps, random actions, f64.  Prints the parity report of tests/harness.py."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import binding as ob
from tests.harness import run_parity, reference_reset_params
from tennisbot_rl_b200.batch import TennisBatch
for env, n, steps in (("SwingRacket-v0", 262144, 78), ("Tennisbot-v0", 65536, 1100)):
    b = TennisBatch(env, n, device=0, seed=101, precision="f64")
    o = ob.OracleEnv(env, n, seed=101, threads=16)
    rng = np.random.default_rng(77)
    init = reference_reset_params(o.kind, n, rng)
    g0 = b.reset(init=init).cpu().numpy(); o0 = o.reset(init=init)
    assert (g0 == o0).all()
    t0 = time.time()
    rep, valid = run_parity(b, o, steps, lambda t, _o: rng.uniform(-1, 1, (n, o.act_dim)), band=0.0, check_state_every=13 if env.startswith("Swing") else 100)
    print(env, n, steps, rep, "stats equal:", bool((b.read_stats() == o.read_stats()).all()), "%.1fs" % (time.time() - t0), flush=True)
