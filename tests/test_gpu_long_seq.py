"""Parity at the long sequences of BASELINE config 5 (reference experiments/scaling.py:387-422: `medium` fits of one
(layer, head) at 4096-32768 tokens): final per-fit CosSim and the loss curve against the oracle (reference
siren.py:70-149 restated) run in torch eager fp32 on the same GPU, TF32 off -- fp32 mode within 1e-3, bf16 mode within
5e-3 (BASELINE.json north_star), through the C ABI.  300 epochs: the gate is on the trajectory so far, the
2000-epoch gate at the benched length is tests/test_gpu_full_length.py.
"""

import pytest
import torch

import nerf_attention as na
from oracle import siren_oracle as orc
from gpu_util import model_from_state

pytestmark = pytest.mark.gpu

LAYER, HEAD, D, EPOCHS = 5, 3, 128, 300
COS_ATOL = {'fp32': 1e-3, 'bf16': 5e-3}


@pytest.fixture(scope='module', params=[4096, 32768])
def runs(request, cuda_device):
    from nerf_attention.extract import synthetic_head
    n = request.param
    keys, values = synthetic_head(LAYER, HEAD, n, 32, 8, D)
    cfg = next(c for c in na.CONFIGS_FULL if c.name == 'medium')
    specs = []
    for is_value, kv in enumerate((keys, values)):
        torch.manual_seed(1000 * LAYER + 100 * HEAD + 10 * is_value + 2)
        specs.append({'kv': kv, 'state': orc.init_state(cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, D)})
    out = {'n': n, 'cfg': cfg}
    for prec in ('fp32', 'bf16'):
        jobs = [na.FitJob(s['kv'], cfg, model_from_state(cfg, D, s['state'])) for s in specs]
        out[prec] = na.fit_many(jobs, epochs=EPOCHS, device='cuda', verbose=False, precision=prec)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False               # the reference never enables TF32
    try:
        out['oracle'] = [orc.fit(s['kv'], cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, epochs=EPOCHS, lr=1e-4,
                                 device='cuda', log_every=10 ** 9, init={k: v.clone() for k, v in s['state'].items()})
                         for s in specs]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_long_sequence_fit_matches_oracle(runs, prec):
    for name, got, ref in zip(('key', 'value'), runs[prec], runs['oracle']):
        diff = abs(got.final_cosine_mean - ref.final_cosine_mean)
        print(f"\nN={runs['n']} medium {name} {prec}: oracle {ref.final_cosine_mean:.6f} got {got.final_cosine_mean:.6f} "
              f"|diff| {diff:.2e} loss {got.losses[-1]:.5f} / {ref.losses[-1]:.5f}")
        assert got.seq_len == runs['n'] and len(got.cosine_sims) == runs['n']
        assert diff <= COS_ATOL[prec]
        assert got.losses[-1] == pytest.approx(ref.losses[-1], rel=0.05)
        assert got.losses[0] == pytest.approx(ref.losses[0], rel=2e-3 if prec == 'fp32' else 2e-2)
        assert got.compression_ratio == pytest.approx(ref.compression_ratio)
