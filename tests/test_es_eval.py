"""ES evaluation adapter: the per-env-weight GatedCNN and normaliser maths on CPU, the batched fitness on a GPU."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden"


def _reference_arch(input_c, action_dim):
    """The architecture of tennisbot/ES/policies.py:59-128, rebuilt from torch layers for the comparison."""
    import torch.nn as nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_0 = nn.Conv1d(input_c, 8, 2, dilation=1)
            self.conv_gate_0 = nn.Conv1d(input_c, 8, 2, dilation=1)
            self.conv_1 = nn.Conv1d(8, 12, 2, dilation=2)
            self.conv_gate_1 = nn.Conv1d(8, 12, 2, dilation=2)
            self.conv_2 = nn.Conv1d(12, action_dim, 2, dilation=4)

        def forward(self, x):
            h = torch.tanh(self.conv_0(x)) * torch.sigmoid(self.conv_gate_0(x))
            h = torch.tanh(self.conv_1(h)) * torch.sigmoid(self.conv_gate_1(h))
            return self.conv_2(h).squeeze()

    return Net()


def test_gated_cnn_per_env_weights_match_module():
    from tennisbot_rl_b200 import es_eval

    assert es_eval.num_params(6, 6) == 766
    saved = np.load(GOLD / "es_swing_weights.npz")["weights"]
    assert saved.shape == (766,) and abs(np.abs(saved).sum() - 208.02) < 0.01
    torch.manual_seed(0)
    ws = [torch.from_numpy(saved)] + [torch.randn(766) * 0.3 for _ in range(4)]
    hist = torch.randn(len(ws), 6, 8)
    out = es_eval.gated_cnn_forward(torch.stack(ws), hist, 6)
    for i, w in enumerate(ws):
        net = _reference_arch(6, 6)
        torch.nn.utils.vector_to_parameters(w.clone(), net.parameters())
        np.testing.assert_allclose(out[i].numpy(), net(hist[i:i + 1]).detach().numpy(), rtol=1e-5, atol=1e-6)


def test_batched_normalizer_matches_scalar_recurrence():
    from tennisbot_rl_b200 import es_eval

    rng = np.random.default_rng(1)
    xs = rng.normal(2.0, 3.0, (20, 4, 6))
    bn = es_eval.BatchedNormalizer(4, 6, "cpu")
    n, mean, md = np.zeros((4, 6)), np.zeros((4, 6)), np.zeros((4, 6))
    for x in xs:
        bn.observe(torch.from_numpy(x))
        n += 1
        last = mean.copy()
        mean += (x - mean) / n
        md += (x - last) * (x - mean)
        var = (md / n).clip(min=1e-2)
        np.testing.assert_allclose(bn.normalize(torch.from_numpy(x)).numpy(), (x - mean) / np.sqrt(var), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_batched_fitness_matches_sequential_oracle_evaluation(oracle_lib):
    """Same semantics as fitness_static: evaluate 3 individuals x 4 repeats in one CUDA batch and, separately, one
    (individual, repeat) at a time through the CPU oracle with numpy policy maths; returns must agree."""
    from tennisbot_rl_b200 import es_eval

    saved = np.load(GOLD / "es_swing_weights.npz")["weights"]
    rng = np.random.default_rng(0)
    pop = np.stack([saved, saved + 0.05 * rng.standard_normal(766).astype(np.float32),
                    0.3 * rng.standard_normal(766).astype(np.float32)])
    repeats, seed = 4, 11
    fit, per_ep = es_eval.batched_fitness_static(pop, "SwingRacket-v0", repeats=repeats, seed=seed)
    per_ep = per_ep.cpu().numpy()
    assert fit.shape == (3,) and per_ep.shape == (3, repeats)
    for e in range(3 * repeats):
        o = oracle_lib.OracleEnv("SwingRacket-v0", 1, env_id_offset=e, seed=seed, auto_reset=False)
        net = _reference_arch(6, 6)
        torch.nn.utils.vector_to_parameters(torch.from_numpy(pop[e // repeats].copy()), net.parameters())
        bn = es_eval.BatchedNormalizer(1, 6, "cpu")
        obs = torch.from_numpy(o.reset())
        bn.observe(obs)
        hist = bn.normalize(obs)[:, :, None].repeat(1, 1, 8)
        total = 0.0
        for _ in range(26):
            with torch.no_grad():
                a = net(hist).clamp(-1, 1).numpy().astype(np.float32)
            r = o.step(a[None])
            ob = torch.from_numpy(r["obs"])
            bn.observe(ob)
            hist = torch.cat([hist[:, :, 1:], bn.normalize(ob)[:, :, None]], dim=2)
            total += float(r["reward"][0])
            if r["done"][0]:
                break
        assert per_ep[e // repeats, e % repeats] == pytest.approx(total, abs=2e-3), e
    assert fit[0] > fit[2]  # the saved ES policy beats random weights
