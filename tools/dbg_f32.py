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
valid = np.ones(n, bool); last = 0
for t in range(steps):
    a = rng.uniform(-1, 1, (n, o.act_dim)).astype(np.float32)
    g = b.step(torch.from_numpy(a).cuda()); r = o.step(a, want_margin=True)
    gev = g[4].cpu().numpy(); gdone = g[2].cpu().numpy()
    valid &= (r['events'] & 1) == 0
    valid &= r['margin'] > 2e-4
    gs = b.get_state().cpu().numpy(); os_ = o.get_state()
    err = np.abs(gs - os_)[:, :22] * valid[:, None]
    i, j = np.unravel_index(err.argmax(), err.shape)
    mism = valid & (((gev & 191) != (r["events"] & 191)) | (gdone != r["done"]))
    if mism.any():
        k = np.nonzero(mism)[0][0]
        print(t, 'MISMATCH env', k, 'gpu ev', gev[k], gdone[k], 'ora ev', r['events'][k], r['done'][k], 'margin', r['margin'][k]); print(' gpu', np.round(gs[k, :22], 5)); print(' ora', np.round(os_[k, :22], 5)); valid &= ~mism
    if err.max() > 3 * last and err.max() > 1e-4:
        last = err.max()
        print(t, 'max state err %.3e env %d word %d' % (err.max(), i, j), 'ev', r['events'][i], 'step', os_[i, 29], 'margin', r['margin'][i])
        print(' gpu', np.round(gs[i, :22], 5)); print(' ora', np.round(os_[i, :22], 5))
