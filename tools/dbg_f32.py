import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import binding as ob
from tennisbot_rl_b200.batch import TennisBatch
from tests.harness import reference_reset_params
env = sys.argv[1]; steps = int(sys.argv[2]); n = 4096
b = TennisBatch(env, n, precision='f32', seed=21); o = ob.OracleEnv(env, n, seed=21, threads=8)
rng = np.random.default_rng(6)
init = reference_reset_params(o.kind, n, rng)
b.reset(init=init); o.reset(init=init)
for t in range(steps):
    a = rng.uniform(-1, 1, (n, o.act_dim)).astype(np.float32)
    b.step(torch.from_numpy(a).cuda()); r = o.step(a)
    gs = b.get_state().cpu().numpy(); os_ = o.get_state()
    err = np.abs(gs - os_)[:, :28]
    i, j = np.unravel_index(err.argmax(), err.shape)
    if err.max() > 1e-4 or t % 10 == 0:
        print(t, 'max state err %.3e env %d word %d' % (err.max(), i, j), 'g', gs[i, j], 'o', os_[i, j], 'ev', r['events'][i], 'step', os_[i, 29])
        if err.max() > 1e-3:
            print(' gpu', np.round(gs[i, :22], 5)); print(' ora', np.round(os_[i, :22], 5)); break
