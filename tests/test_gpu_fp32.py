"""fp32 mode (NA_PREC_FP32) of the CUDA path against the oracle and the reference's golden
vectors, through the C ABI.  Gates (BASELINE.json north_star): forward within 1e-5 relative,
final per-fit CosSim within 1e-3."""

import json

import numpy as np
import pytest
import torch

import nerf_attention as na
from nerf_attention import _native
from nerf_attention.evaluate import PackedModels, _load_model_from_checkpoint
from oracle import siren_oracle as orc
from gpu_util import flat, gpu_fit, model_from_state, oracle_fit, rel_err, seeded_state, smooth_tensor

pytestmark = pytest.mark.gpu

FWD_RTOL = 1e-5        # north_star: fp32 forward within 1e-5 relative error
COS_ATOL = 1e-3        # north_star: final per-fit CosSim within 1e-3


def cfg_by_name(name):
    return next(c for c in na.CONFIGS_FULL if c.name == name)


@pytest.mark.parametrize('name', ['tiny', 'medium', 'hifreq'])
def test_forward_matches_reference_golden(golden, cuda_device, name):
    cfg = cfg_by_name(name)
    torch.manual_seed(77)
    model = na.SIREN(cfg, out_features=128)
    out = PackedModels([model], 96).forward()[0].cpu().numpy()
    ref = golden['forward'][f'fwd_{name}']
    assert rel_err(out, ref) <= FWD_RTOL


@pytest.mark.parametrize('name,n', [('small', 2048), ('large', 512), ('deep', 1000), ('lofreq', 4096)])
def test_forward_matches_oracle_at_size(cuda_device, name, n):
    cfg = cfg_by_name(name)
    state = seeded_state(cfg, 128, 5)
    mean, std = torch.randn(1, 128), torch.rand(1, 128) + 0.5
    packed = PackedModels([model_from_state(cfg, 128, state)] * 3, n, [mean] * 3, [std] * 3)
    ref = orc.forward(state, cfg.omega_0, orc.positions_for(n))
    exact = orc.forward_exact(state, cfg.omega_0, orc.positions_for(n))
    ref_err = rel_err(ref, exact)                    # the reference arithmetic's own rounding error
    out = packed.forward().cpu()
    assert rel_err(out[0], exact) <= FWD_RTOL and torch.equal(out[0], out[2])
    assert rel_err(out[0], ref) <= FWD_RTOL + 2 * ref_err
    out = packed.forward(denormalise=True).cpu()
    assert rel_err(out[1], exact * std.double() + mean.double()) <= FWD_RTOL
    assert rel_err(out[1], ref * std + mean) <= FWD_RTOL + 2 * ref_err


@pytest.mark.parametrize('h,l,w,n,d', [(64, 1, 30.0, 192, 32), (256, 2, 60.0, 300, 16), (128, 3, 15.0, 128, 128)])
def test_one_step_loss_gradients_and_adam(cuda_device, h, l, w, n, d):
    cfg = na.SIRENConfig(h, l, w, 'kat')
    state = seeded_state(cfg, d, 21)
    kv = smooth_tensor(9, n, d)
    res = gpu_fit(kv, cfg, 1, 'fp32', state, keep_optimizer_state=True)
    _, _, t_norm = orc.normalise(kv)
    loss, grads = orc.loss_and_grads(state, w, orc.positions_for(n), t_norm)
    assert res.losses[0] == pytest.approx(loss, rel=2e-5)
    g_ref = flat(grads)
    m, v = (t.cpu() for t in res.model.adam_state)
    g_gpu = m / 0.1                                                  # m_1 = (1 - beta1) g
    assert rel_err(g_gpu, g_ref) <= 2e-4
    cos = torch.nn.functional.cosine_similarity(g_gpu, g_ref, dim=0).item()
    assert cos > 1 - 1e-6
    # one Adam step from the oracle's gradient
    p = flat(state).clone(); mm = torch.zeros_like(p); vv = torch.zeros_like(p)
    orc.adam_reference_step(p, g_ref, mm, vv, 1, 1e-4)
    p_gpu = flat(res.model.state_dict()).cpu()
    assert (p_gpu - p).abs().max().item() <= 2.5e-6                  # steps are +-1e-4; sign flips only where g ~ 0
    assert torch.allclose(v, vv, rtol=1e-3, atol=1e-12)


@pytest.mark.parametrize('name', ['tiny', 'small', 'medium', 'hifreq', 'deep'])
def test_fit_matches_reference_golden_trajectory(golden, cuda_device, name):
    spec = golden['meta'][f'fit_{name}']
    h, l, w, n, d, epochs, _dseed, mseed = spec['spec']
    cfg = na.SIRENConfig(int(h), int(l), float(w), name)
    kv = torch.from_numpy(golden['fits'][f'{name}_kv'])
    torch.manual_seed(int(mseed))                                   # same CPU stream as the reference run
    res = na.fit_siren(kv, cfg, epochs=int(epochs), device='cuda', verbose=False, precision='fp32')
    np.testing.assert_allclose(res.losses, golden['fits'][f'{name}_losses'], rtol=2e-3)
    assert abs(res.final_cosine_mean - spec['final_cosine_mean']) <= COS_ATOL
    assert abs(res.final_cosine_min - spec['final_cosine_min']) <= 5 * COS_ATOL
    assert res.final_mse == pytest.approx(spec['final_mse'], rel=5e-3)
    assert res.final_cosine_std == pytest.approx(spec['final_cosine_std'], rel=2e-2, abs=1e-4)
    np.testing.assert_allclose(res.cosine_sims, golden['fits'][f'{name}_cos'], atol=5e-3)
    np.testing.assert_allclose(res.target_mean.numpy(), golden['fits'][f'{name}_mean'], atol=1e-6)
    np.testing.assert_allclose(res.target_std.numpy(), golden['fits'][f'{name}_std'], rtol=1e-5)
    for key in ('compression_ratio', 'raw_size_bytes', 'siren_size_bytes', 'num_parameters'):
        assert getattr(res, key) == spec[key]
    assert res.seq_len == n and res.d_head == d and len(res.losses) == epochs
    wf = res.model.state_dict()[f'network.{int(l) + 1}.weight'].cpu().numpy()
    assert rel_err(wf, golden['fits'][f'{name}_wf']) <= 5e-3


@pytest.mark.parametrize('name,epochs', [('medium', 150), ('large', 40), ('tiny', 300)])
def test_fit_at_baseline_shape_matches_oracle(cuda_device, name, epochs):
    """N=2048, D=128 synthetic keys (layer 16, head 0) as in BASELINE configs 2-3."""
    from nerf_attention.extract import synthetic_head
    kv, _ = synthetic_head(16, 0, 2048, 32, 8, 128)
    cfg = cfg_by_name(name)
    state = seeded_state(cfg, 128, 16010)
    ref = oracle_fit(kv, cfg, epochs, state)
    res = gpu_fit(kv, cfg, epochs, 'fp32', state)
    np.testing.assert_allclose(res.losses, ref.losses, rtol=2e-3)
    assert abs(res.final_cosine_mean - ref.final_cosine_mean) <= COS_ATOL
    assert res.final_mse == pytest.approx(ref.final_mse, rel=5e-3)
    assert np.abs(res.cosine_sims - ref.cosine_sims).max() <= 5e-3


def test_batch_is_deterministic_and_order_independent(cuda_device):
    cfgs = [na.SIRENConfig(64, 1, 30.0, 'a'), na.SIRENConfig(128, 2, 30.0, 'b'), na.SIRENConfig(64, 1, 60.0, 'c')]
    tensors = [smooth_tensor(i, 256, 32) for i in range(2)]
    spec = [(t, c, seeded_state(c, 32, 100 + 10 * ti + ci)) for ti, t in enumerate(tensors) for ci, c in enumerate(cfgs)]

    def run(order):
        jobs = [na.FitJob(spec[i][0], spec[i][1], model_from_state(spec[i][1], 32, spec[i][2])) for i in order]
        return na.fit_many(jobs, epochs=30, device='cuda', verbose=False, precision='fp32')
    a = run(range(6))
    b = run(range(6))
    c = run([5, 3, 1, 4, 2, 0])
    single = gpu_fit(spec[4][0], spec[4][1], 30, 'fp32', spec[4][2])
    for i in range(6):
        assert a[i].losses == b[i].losses                                     # bitwise run-to-run
        assert a[i].losses == c[[5, 3, 1, 4, 2, 0].index(i)].losses           # grouping-independent
    assert a[4].losses == single.losses and a[4].final_cosine_mean == single.final_cosine_mean
    assert all(x.losses[-1] < x.losses[0] for x in a)


def test_edge_cases(cuda_device):
    cfg = na.SIRENConfig(64, 1, 30.0, 'edge')
    kv = smooth_tensor(2, 100, 8)
    kv[:, 3] = 0.25                                        # constant dimension -> std clamp 1e-3
    state = seeded_state(cfg, 8, 1)
    ref = oracle_fit(kv, cfg, 20, state)
    kv_before = kv.clone()
    res = gpu_fit(kv, cfg, 20, 'fp32', state)
    assert torch.equal(kv, kv_before)                      # caller's tensor untouched
    assert res.target_std[0, 3].item() == pytest.approx(1e-3)
    np.testing.assert_allclose(res.losses, ref.losses, rtol=2e-3)
    assert abs(res.final_cosine_mean - ref.final_cosine_mean) <= COS_ATOL
    # epochs = 0: metrics of the untouched initial model
    res0 = gpu_fit(kv, cfg, 0, 'fp32', state)
    ref0 = oracle_fit(kv, cfg, 0, state)
    assert res0.losses == [] and abs(res0.final_cosine_mean - ref0.final_cosine_mean) <= 1e-5
    assert flat(res0.model.state_dict()).cpu().equal(flat(state))
    # a CUDA-resident input is accepted as well
    res_dev = gpu_fit(kv.cuda(), cfg, 20, 'fp32', state)
    assert res_dev.losses == res.losses
    with pytest.raises(_native.NativeError, match='multiple of'):
        gpu_fit(smooth_tensor(2, 64, 6), cfg, 1, 'fp32', seeded_state(cfg, 6, 1))
    with pytest.raises(ValueError):
        na.fit_many([na.FitJob(torch.zeros(4), cfg)], epochs=1)


def test_fit_kv_cache_quick_end_to_end(cuda_device, tmp_path, capsys):
    kv_dir, out_dir = tmp_path / 'kv', tmp_path / 'fits'
    na.extract_kv_cache_synthetic(seq_len=128, num_layers=4, num_kv_heads=2, head_dim=16, output_dir=kv_dir)
    records = na.fit_kv_cache(kv_dir, out_dir, epochs=60, device='cuda', quick=True, precision='fp32',
                              seed_fn=lambda j: 1000 * j['layer'] + 100 * j['head'] + 10 * (j['kv_type'] == 'value')
                              + j['config_index'])
    assert len(records) == 3 * 1 * 2 * 2
    assert json.loads((out_dir / 'fit_results.json').read_text()) == records
    printed = capsys.readouterr().out
    assert 'RESULTS SUMMARY' in printed and '[12/12] L3_H0_value_medium' in printed
    ckpts = sorted(out_dir.glob('*_model.pt'))
    assert [p.name for p in ckpts] == sorted(f'L{l}_H0_{kv}_medium_model.pt' for l in (0, 2, 3) for kv in ('key', 'value'))
    # the checkpoint reproduces its own metrics through plain torch on the CPU (reference evaluate.py:148-153)
    ckpt = torch.load(ckpts[0], map_location='cpu', weights_only=True)
    model = _load_model_from_checkpoint(ckpt, 'cpu')
    m = ckpt['metrics']
    blob = torch.load(kv_dir / f"layer_{m['layer']:02d}.pt", weights_only=True)
    original = blob['keys' if m['kv_type'] == 'key' else 'values'][m['head']]
    with torch.no_grad():
        pred = model(torch.linspace(0, 1, 128).unsqueeze(1)) * ckpt['target_std'] + ckpt['target_mean']
    cos = torch.nn.functional.cosine_similarity(pred, original, dim=1)
    assert cos.mean().item() == pytest.approx(m['final_cosine_mean'], abs=1e-5)
    # and against the oracle run with the same seed convention
    rec = next(r for r in records if r['name'] == 'L2_H0_value_small')
    torch.manual_seed(1000 * 2 + 10 + 0)
    ref = orc.fit(blob_for(kv_dir, 2)['values'][0], 128, 1, 30.0, epochs=60, device='cpu', log_every=10 ** 9)
    assert abs(rec['final_cosine_mean'] - ref.final_cosine_mean) <= COS_ATOL
    assert rec['compression_ratio'] == ref.compression_ratio


def blob_for(kv_dir, layer):
    return torch.load(kv_dir / f'layer_{layer:02d}.pt', weights_only=True)


@pytest.mark.parametrize('precision,rtol,cos_atol', [('fp32', 2e-3, 1e-3), ('bf16', 2e-2, 5e-3)])
def test_progress_metrics_match_reference_logging(cuda_device, capsys, precision, rtol, cos_atol):
    """siren.py:107-115: every log_every epochs the reference prints NormMSE, RealMSE and CosSim of the
    prediction made with the weights that epoch starts from.  The batched path evaluates them on the
    device between graph replays (nerfattn_fit_batched_ex) and prints the same line."""
    cfg = na.SIRENConfig(128, 2, 30.0, 'kat')
    kv = smooth_tensor(21, 512, 128)
    state = seeded_state(cfg, 128, 5)
    ref = orc.fit(kv, 128, 2, 30.0, epochs=60, lr=1e-4, device='cpu', log_every=20,
                  init={k: v.clone() for k, v in state.items()})
    job = na.FitJob(kv, cfg, model_from_state(cfg, 128, state))
    res = na.fit_many([job], epochs=60, device='cuda', verbose=True, log_every=20, precision=precision)[0]
    assert [p[0] for p in res.progress] == [p[0] for p in ref.progress] == [20, 40, 60]
    for got, want in zip(res.progress, ref.progress):
        assert got[1] == pytest.approx(want[1], rel=rtol)            # NormMSE
        assert got[2] == pytest.approx(want[2], rel=rtol)            # RealMSE
        assert abs(got[3] - want[3]) <= cos_atol                     # CosSim
    out = capsys.readouterr().out
    assert f'  Epoch 40/60 | NormMSE: {res.progress[1][1]:.6f} | RealMSE: {res.progress[1][2]:.6f} | CosSim: {res.progress[1][3]:.4f}' in out
    # quiet runs do not pay for the evaluations
    quiet = na.fit_many([na.FitJob(kv, cfg, model_from_state(cfg, 128, state))], epochs=60, device='cuda',
                        verbose=False, precision=precision)[0]
    assert quiet.progress == [] and quiet.losses == res.losses


def test_prenormalised_targets_give_the_same_fit(cuda_device):
    """NA_FIT_TARGETS_PRENORMALISED: handing in (t - mean) / std with the statistics must train the same model and
    report the same de-normalised metrics as handing in the raw tensor (siren.py:85-87 done by the caller)."""
    cfg = na.SIRENConfig(64, 1, 30.0, 'kat')
    kv = smooth_tensor(11, 200, 64) * 3.0 + 1.5
    state = seeded_state(cfg, 64, 5)
    mean, std, t_norm = orc.normalise(kv)
    raw = gpu_fit(kv, cfg, 30, 'fp32', state)
    job = na.FitJob(t_norm.contiguous(), cfg, model_from_state(cfg, 64, state), target_mean=mean, target_std=std)
    pre = na.fit_many([job], epochs=30, device='cuda', verbose=False, precision='fp32')[0]
    assert np.allclose(pre.losses, raw.losses, rtol=1e-5)
    assert pre.final_mse == pytest.approx(raw.final_mse, rel=1e-4)
    assert pre.final_cosine_mean == pytest.approx(raw.final_cosine_mean, abs=1e-6)
    assert np.allclose(pre.cosine_sims, raw.cosine_sims, atol=1e-5)
    assert torch.allclose(pre.target_mean, mean, atol=0) and torch.allclose(pre.target_std, std, atol=0)


def test_per_position_cosine_matches_the_fit_records(cuda_device, tmp_path):
    """evaluate.per_position_cosine (reference evaluate.py:148-153) from the saved checkpoints and layer files equals
    the cosine_sims the fit itself reported, and the checkpoints are the reference's size."""
    from nerf_attention.evaluate import per_position_cosine, plot_per_position_error
    kv_dir, out_dir = tmp_path / 'kv', tmp_path / 'fits'
    na.extract_kv_cache_synthetic(seq_len=192, num_layers=4, num_kv_heads=2, head_dim=64, output_dir=kv_dir)
    records = na.fit_kv_cache(kv_dir, out_dir, epochs=25, device='cuda', quick=True, precision='fp32')
    curves = per_position_cosine(out_dir, kv_dir, device='cuda')
    assert len(curves) == 4 and all(c.shape == (192,) for c in curves.values())
    by_name = {r['name']: r for r in records}
    for name, curve in curves.items():
        assert float(curve.mean()) == pytest.approx(by_name[name]['final_cosine_mean'], abs=2e-5)
        assert float(curve.min()) == pytest.approx(by_name[name]['final_cosine_min'], abs=2e-5)
        size = (out_dir / f'{name}_model.pt').stat().st_size
        assert size <= 4 * by_name[name]['num_parameters'] + 16384          # its own weights only, not the batch buffer
    plot_per_position_error(out_dir, kv_dir, tmp_path / 'figs', device='cuda')
    assert (tmp_path / 'figs' / 'per_position_cosine.json').exists()
