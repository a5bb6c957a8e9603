"""Decode-side kernels: SIREN-evaluated q.K (output layer folded into the query) and the
bandwidth-bound fp16 KV-read baseline, through the C ABI, against the oracle."""

import json

import numpy as np
import pytest
import torch

import nerf_attention as na
from nerf_attention.evaluate import PackedModels, kvread_qk, profile_decode, profile_latency
from nerf_attention.fit import _result_to_record, _save_model
from oracle import siren_oracle as orc
from gpu_util import gpu_fit, model_from_state, rel_err, seeded_state, smooth_tensor

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,heads,d', [(512, 3, 128), (2048, 8, 128), (300, 2, 64), (1, 1, 256)])
def test_kvread_qk_matches_torch(cuda_device, n, heads, d):
    g = torch.Generator(device='cuda').manual_seed(n)
    k = torch.randn(heads, n, d, device='cuda', generator=g).half()
    q = torch.randn(heads, d, device='cuda', generator=g).half()
    out = kvread_qk(k, q)
    ref = torch.stack([orc.kvread_scores(k[i].cpu(), q[i].cpu()) for i in range(heads)])
    assert rel_err(out.cpu(), ref) <= 1e-5          # fp32 accumulate of exact fp16 products


@pytest.mark.parametrize('name,n', [('medium', 512), ('tiny', 2048), ('large', 256), ('deep', 1024), ('medium', 333)])
@pytest.mark.parametrize('precision,tol', [('fp32', 2e-5), ('bf16', 3e-2)])
def test_decode_qk_matches_oracle(cuda_device, name, n, precision, tol):
    cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
    heads = 3
    states = [seeded_state(cfg, 128, 70 + i) for i in range(heads)]
    g = torch.Generator().manual_seed(1)
    means = [torch.randn(1, 128, generator=g) * 0.1 for _ in range(heads)]
    stds = [torch.rand(1, 128, generator=g) + 0.5 for _ in range(heads)]
    q = torch.randn(heads, 128, generator=g).half()
    packed = PackedModels([model_from_state(cfg, 128, s) for s in states], n, means, stds)
    out = packed.decode_qk(q.cuda(), precision).clone()
    again = packed.decode_qk(q.cuda(), precision, reuse_setup=True)
    assert torch.equal(out, again)
    for i in range(heads):
        ref = orc.decode_scores(states[i], cfg.omega_0, means[i], stds[i], q[i], n)
        exact = orc.decode_scores(states[i], cfg.omega_0, means[i], stds[i], q[i], n, dtype=torch.float64)
        # the kernel against the exact function of the same weights, and against the reference's fp32
        # arithmetic up to that arithmetic's own rounding error (oracle/siren_oracle.py::forward_exact)
        assert rel_err(out[i].cpu(), exact) <= tol, i
        assert rel_err(out[i].cpu(), ref) <= tol + 2 * rel_err(ref, exact), i
    # a new query with the cached set-up is the per-token path
    q2 = torch.randn(heads, 128, generator=g).half()
    out2 = packed.decode_qk(q2.cuda(), precision, reuse_setup=True)
    with pytest.raises(ValueError):
        packed.decode_qk(q2.cuda(), precision, out=torch.empty_like(out2), reuse_setup=True)
    ref2 = orc.decode_scores(states[1], cfg.omega_0, means[1], stds[1], q2[1], n, dtype=torch.float64)
    assert rel_err(out2[1].cpu(), ref2) <= tol


def test_decode_without_hidden_layers(cuda_device):
    cfg = na.SIRENConfig(64, 0, 30.0, 'flat')
    state = seeded_state(cfg, 128, 3)
    q = torch.randn(1, 128).half()
    packed = PackedModels([model_from_state(cfg, 128, state)], 200)
    out = packed.decode_qk(q.cuda(), 'fp32')
    ref = orc.decode_scores(state, 30.0, torch.zeros(1, 128), torch.ones(1, 128), q[0], 200)
    assert rel_err(out[0].cpu(), ref) <= 2e-5


def test_profile_latency_outputs(cuda_device, tmp_path):
    cfg = na.CONFIGS_FULL[2]
    kv = smooth_tensor(4, 512, 128)
    res = gpu_fit(kv, cfg, 5, 'fp32', seeded_state(cfg, 128, 9))
    rec = _result_to_record('L0_H0_key_medium', 0, 0, 'key', res)
    _save_model(tmp_path, rec['name'], res, rec)
    rows = profile_latency(tmp_path, tmp_path / 'fig', device='cuda')
    saved = json.loads((tmp_path / 'fig' / 'latency_results.json').read_text())
    assert saved == rows and len(rows) == 1
    reference_keys = ['name', 'config', 'siren_time_ms', 'hbm_time_4060_ms', 'hbm_time_h100_ms',
                      'speedup_vs_4060', 'speedup_vs_h100', 'num_params']
    assert list(rows[0])[:8] == reference_keys                   # reference evaluate.py:206-215
    assert rows[0]['num_params'] == 164992 and rows[0]['hbm_time_4060_ms'] == pytest.approx(512 * 128 * 2 / 272e9 * 1e3)
    assert rows[0]['siren_time_ms'] > 0 and rows[0]['hbm_time_b200_measured_ms'] > 0


def test_profile_decode_table(cuda_device):
    cfg = na.CONFIGS_FULL[2]
    models = [model_from_state(cfg, 128, seeded_state(cfg, 128, i)) for i in range(2)]
    table = profile_decode(models, [512, 1024], heads_per_launch=8, warmup=2, runs=3)
    assert [r['seq_len'] for r in table] == [512, 1024]
    for r in table:
        assert r['kvread_us'] > 0 and r['siren_fp32_us'] > 0 and r['siren_bf16_us'] > 0


def test_scaling_experiment_and_layer_profile_outputs(cuda_device, tmp_path):
    """experiments/scaling.py on the batched path: same files and keys as the reference
    (scaling.py:124-222, 387-422), plus the measured-B200 columns."""
    from nerf_attention.experiments import run_full_layer_profile, run_scaling_experiment
    rows = run_scaling_experiment('synthetic', [256, 512], tmp_path / 'scaling', epochs=40, precision='bf16',
                                  num_layers=3, num_kv_heads=2, head_dim=128, heads_per_launch=4)
    saved = json.loads((tmp_path / 'scaling' / 'scaling_results.json').read_text())
    assert sorted(saved) == ['256', '512'] and sorted(rows) == [256, 512]
    reference_keys = ['seq_len', 'actual_tokens', 'autocorr_keys', 'autocorr_values', 'spectral_keys', 'spectral_values',
                      'avg_cossim_keys', 'avg_cossim_values', 'avg_compression', 'siren_time_ms', 'hbm_4060_ms',
                      'hbm_h100_ms', 'num_experiments']
    for n in (256, 512):
        r = saved[str(n)]
        assert list(r)[:13] == reference_keys                       # reference scaling.py:200-214
        assert r['num_experiments'] == 6 and r['siren_time_ms'] > 0
        assert r['hbm_4060_ms'] == pytest.approx(n * 128 * 2 / 272e9 * 1000)
        assert r['hbm_b200_measured_ms'] > 0 and r['siren_decode_qk_ms'] > 0
        assert 0.0 < r['avg_cossim_keys'] <= 1.0
        assert 0.5 < r['autocorr_keys'] < 1.0 and 0.5 < r['autocorr_values'] < 1.0       # synthetic KV is smooth (analyze.py)
        assert 0.0 < r['spectral_keys'] <= 1.0 and 0.0 < r['spectral_values'] <= 1.0
        assert (tmp_path / 'scaling' / f'seq_{n}' / 'analysis' / 'analysis_results.json').exists()
        assert len(list((tmp_path / 'scaling' / f'seq_{n}' / 'fits').glob('*_model.pt'))) == 6
    ckpt = torch.load(tmp_path / 'scaling' / 'seq_256' / 'fits' / 'L0_H0_key_medium_model.pt', weights_only=True)
    assert sorted(ckpt) == ['config', 'metrics', 'model_state', 'target_mean', 'target_std']   # scaling.py:175-187
    assert ckpt['metrics'] == {'name': 'L0_H0_key_medium', 'config_name': 'medium', 'seq_len': 256,
                               'raw_size_bytes': 256 * 128 * 2}
    cross = json.loads((tmp_path / 'scaling' / 'crossover_data.json').read_text())
    assert 'siren_fit_log_slope' in cross and 'b200_decode_fit_log_slope' in cross
    # the batched call gives the same fits as one fit_siren per tensor would (order / batch independent)
    data = torch.load(tmp_path / 'scaling' / 'seq_256' / 'kv_cache' / 'layer_00.pt', weights_only=True)
    torch.manual_seed(123)
    prof = run_full_layer_profile(tmp_path / 'scaling' / 'seq_256' / 'kv_cache', tmp_path / 'profile', epochs=20,
                                  precision='bf16')
    assert [(r['layer'], r['kv_type']) for r in prof] == [(0, 'key'), (0, 'value'), (1, 'key'), (1, 'value'),
                                                          (2, 'key'), (2, 'value')]
    assert json.loads((tmp_path / 'profile' / 'full_layer_profile.json').read_text()) == prof
    assert data['keys'].shape == (2, 256, 128)


@pytest.mark.parametrize('name,n', [('medium', 512), ('tiny', 1024), ('large', 256), ('deep', 384), ('small', 777)])
@pytest.mark.parametrize('precision,tol', [('fp32', 5e-5), ('bf16', 3e-2)])
def test_siren_attention_matches_oracle(cuda_device, name, n, precision, tol):
    """softmax(scale * q.K_hat) @ V_hat from key and value SIRENs against the CPU restatement."""
    from nerf_attention.evaluate import siren_attention
    cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
    heads = 3
    ks = [seeded_state(cfg, 128, 170 + i) for i in range(heads)]
    vs = [seeded_state(cfg, 128, 270 + i) for i in range(heads)]
    g = torch.Generator().manual_seed(2)
    mk = [torch.randn(1, 128, generator=g) * 0.1 for _ in range(heads)]
    sk = [torch.rand(1, 128, generator=g) + 0.5 for _ in range(heads)]
    mv = [torch.randn(1, 128, generator=g) * 0.1 for _ in range(heads)]
    sv = [torch.rand(1, 128, generator=g) + 0.5 for _ in range(heads)]
    q = (torch.randn(heads, 128, generator=g) * 3).half()           # sharp enough that the softmax matters
    keys = PackedModels([model_from_state(cfg, 128, s) for s in ks], n, mk, sk)
    values = PackedModels([model_from_state(cfg, 128, s) for s in vs], n, mv, sv)
    scale = 128 ** -0.5
    out = siren_attention(keys, values, q.cuda(), scale, precision)
    again = siren_attention(keys, values, q.cuda(), scale, precision)
    assert torch.equal(out, again)                                   # deterministic reductions
    for i in range(heads):
        args = (ks[i], vs[i], cfg.omega_0, cfg.omega_0, mk[i], sk[i], mv[i], sv[i], q[i], n, scale)
        ref, exact = orc.decode_attention(*args), orc.decode_attention(*args, dtype=torch.float64)
        assert rel_err(out[i].cpu(), exact) <= tol, i
        assert rel_err(out[i].cpu(), ref) <= tol + 2 * rel_err(ref, exact), i


@pytest.mark.parametrize('n,heads,d', [(512, 3, 128), (2048, 8, 128), (300, 2, 64), (5, 1, 256)])
def test_kvread_attention_matches_torch(cuda_device, n, heads, d):
    from nerf_attention.evaluate import kvread_attention
    g = torch.Generator(device='cuda').manual_seed(n)
    k = torch.randn(heads, n, d, device='cuda', generator=g).half()
    v = torch.randn(heads, n, d, device='cuda', generator=g).half()
    q = torch.randn(heads, d, device='cuda', generator=g).half()
    out = kvread_attention(k, v, q)
    ref = torch.stack([orc.kvread_attention(k[i].cpu(), v[i].cpu(), q[i].cpu(), d ** -0.5) for i in range(heads)])
    assert rel_err(out.cpu(), ref) <= 2e-5
