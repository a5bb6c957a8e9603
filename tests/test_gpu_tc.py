"""BF16 tensor-core mode (NA_PREC_BF16: TMA + tcgen05 + TMEM) through the C ABI.
Gate (BASELINE.json north_star): final per-fit CosSim within 5e-3 of the reference."""

import ctypes
import os

import numpy as np
import pytest
import torch

import nerf_attention as na
from nerf_attention import _native
from oracle import siren_oracle as orc
from gpu_util import flat, gpu_fit, model_from_state, oracle_fit, rel_err, seeded_state, smooth_tensor

pytestmark = pytest.mark.gpu

COS_ATOL_BF16 = 5e-3      # north_star: TF32/BF16 mode final CosSim within 5e-3


def debug_gemm(a, b, m, n, k, batch, a_mn, b_mn):
    c = torch.full((batch, m, n), float('nan'), device='cuda')
    _native.check(_native.lib().nerfattn_debug_gemm_bf16(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, batch,
                                                         int(a_mn), int(b_mn), _native.stream_handle()),
                  'nerfattn_debug_gemm_bf16')
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize('m,n,k,batch', [(128, 64, 64, 1), (256, 256, 256, 2), (384, 128, 512, 3),
                                         (64, 128, 2048, 2), (128, 512, 128, 1), (2048, 256, 256, 5)])
def test_tcgen05_gemm_against_matmul(cuda_device, a_mn, b_mn, m, n, k, batch):
    """Pins the smem descriptors (K-major and MN-major, 128B swizzle), the TMA boxes, the
    instruction descriptor and the TMEM epilogue mapping against a plain matmul."""
    if a_mn == 0 and m % 128:
        pytest.skip('K-major A tiles are full 128-row boxes in the fit')
    g = torch.Generator(device='cuda').manual_seed(m + n + k)
    a = torch.randn(batch, m, k, device='cuda', generator=g).bfloat16()
    b = torch.randn(batch, k, n, device='cuda', generator=g).bfloat16()
    ref = torch.bmm(a.float(), b.float())
    a_store = a.transpose(1, 2).contiguous() if a_mn else a.contiguous()        # [K,M] if MN-major
    b_store = b.contiguous() if b_mn else b.transpose(1, 2).contiguous()        # [N,K] if K-major
    out = debug_gemm(a_store, b_store, m, n, k, batch, a_mn, b_mn)
    assert not torch.isnan(out).any()
    assert rel_err(out.cpu(), ref.cpu()) <= 1e-5


@pytest.mark.parametrize('h,l,w,n,d', [(64, 1, 30.0, 256, 64), (256, 2, 30.0, 384, 128), (128, 3, 60.0, 128, 128),
                                       (512, 1, 30.0, 256, 256),
                                       # ragged sequence lengths: the last row tile is masked
                                       (256, 2, 30.0, 300, 128), (128, 1, 30.0, 100, 64), (512, 2, 30.0, 1000, 128),
                                       (64, 1, 60.0, 129, 128)])
def test_one_step_gradients(cuda_device, h, l, w, n, d):
    cfg = na.SIRENConfig(h, l, w, 'kat')
    state = seeded_state(cfg, d, 31)
    kv = smooth_tensor(8, n, d)
    res = gpu_fit(kv, cfg, 1, 'bf16', state, keep_optimizer_state=True)
    _, _, t_norm = orc.normalise(kv)
    loss, grads = orc.loss_and_grads(state, w, orc.positions_for(n), t_norm)
    assert res.losses[0] == pytest.approx(loss, rel=2e-3)
    g_gpu = res.model.adam_state[0].cpu() / 0.1
    g_ref = flat(grads)
    # per-layer: direction and norm of the gradient (bf16 operands: ~3 significant digits)
    off = 0
    for key, ref in grads.items():
        cnt = ref.numel()
        got = g_gpu[off:off + cnt]
        off += cnt
        cos = torch.nn.functional.cosine_similarity(got, ref.reshape(-1), dim=0).item()
        assert cos > 0.999, (key, cos)
        assert got.norm().item() == pytest.approx(ref.norm().item(), rel=2e-2), key
    assert torch.nn.functional.cosine_similarity(g_gpu, g_ref, dim=0).item() > 0.9995


@pytest.mark.parametrize('name,n,epochs', [('tiny', 1024, 400), ('small', 1024, 400), ('medium', 1024, 400),
                                           ('deep', 512, 300), ('large', 512, 150), ('hifreq', 512, 300),
                                           ('medium', 1000, 300), ('large', 333, 150)])      # ragged lengths
def test_fit_cossim_within_tolerance(cuda_device, name, n, epochs):
    from nerf_attention.extract import synthetic_head
    cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
    keys, values = synthetic_head(8, 1, n, 32, 8, 128)
    for kv in (keys, values):
        state = seeded_state(cfg, 128, 8110)
        ref = oracle_fit(kv, cfg, epochs, state)
        res = gpu_fit(kv, cfg, epochs, 'bf16', state)
        assert abs(res.final_cosine_mean - ref.final_cosine_mean) <= COS_ATOL_BF16
        assert res.losses[-1] == pytest.approx(ref.losses[-1], rel=3e-2)
        assert res.losses[-1] < res.losses[0]
        # the metrics are an fp32 evaluation of the returned weights: check them with torch
        with torch.no_grad():
            pred = res.model.cpu()(torch.linspace(0, 1, n).unsqueeze(1)) * res.target_std + res.target_mean
        cos = torch.nn.functional.cosine_similarity(pred, kv, dim=1)
        assert cos.mean().item() == pytest.approx(res.final_cosine_mean, abs=2e-5)


def test_sweep_groups_mix_and_match_fp32(cuda_device):
    """A miniature 7-architecture sweep in bf16 stays within tolerance of the same sweep in fp32."""
    from nerf_attention.extract import synthetic_head
    keys, values = synthetic_head(0, 0, 256, 32, 8, 128)
    spec = [(t, c, seeded_state(c, 128, 40 + i)) for t in (keys, values) for i, c in enumerate(na.CONFIGS_FULL)]

    def run(prec):
        jobs = [na.FitJob(t, c, model_from_state(c, 128, s)) for t, c, s in spec]
        return na.fit_many(jobs, epochs=120, device='cuda', verbose=False, precision=prec)
    a, b = run('bf16'), run('fp32')
    for x, y in zip(a, b):
        assert abs(x.final_cosine_mean - y.final_cosine_mean) <= COS_ATOL_BF16, x.config.name
        assert x.num_parameters == y.num_parameters
    again = run('bf16')
    assert all(x.losses == y.losses for x, y in zip(a, again))                  # deterministic


def test_unsupported_shapes_fail_loudly(cuda_device):
    cfg = na.SIRENConfig(64, 1, 30.0, 'x')
    with pytest.raises(_native.NativeError, match='bf16 path needs'):
        gpu_fit(smooth_tensor(1, 128, 16), cfg, 1, 'bf16', seeded_state(cfg, 16, 1))
    with pytest.raises(ValueError, match='precision must be one of'):                 # only fp32 and bf16 exist
        gpu_fit(smooth_tensor(1, 128, 128), cfg, 1, 'tf32', seeded_state(cfg, 128, 1))
    with pytest.raises(_native.NativeError, match='not implemented'):                 # precision code 1 is unassigned
        gpu_fit(smooth_tensor(1, 128, 128), cfg, 1, 1, seeded_state(cfg, 128, 1))


@pytest.mark.parametrize('mode,tol', [(0, 1.5e-7), (1, 6e-7)])
def test_sincos_accuracy(cuda_device, mode, tol):
    """Device sin/cos against float64.  mode 0 (polynomial: fp32 path and layer 0) stays within fp32
    rounding; mode 1 (exact Cody-Waite reduction + SFU core, hidden layers of the BF16 path) within
    6e-7 absolute for every |x| a SIREN produces (omega_0 = 60 puts arguments near +-120) -- four
    orders of magnitude below the bf16 rounding applied to its result."""
    g = torch.Generator().manual_seed(5)
    x = torch.cat([torch.linspace(-130, 130, 400001), (torch.rand(200000, generator=g) - 0.5) * 2000,
                   (torch.rand(100000, generator=g) - 0.5) * 0.02,
                   torch.arange(-40, 41, dtype=torch.float32) * (np.pi / 2)]).float()
    xd = x.cuda()
    s, c = torch.empty_like(xd), torch.empty_like(xd)
    _native.check(_native.lib().nerfattn_debug_sincos(xd.data_ptr(), s.data_ptr(), c.data_ptr(), xd.numel(), mode,
                                                      _native.stream_handle()), 'nerfattn_debug_sincos')
    torch.cuda.synchronize()
    x64 = x.double()
    es = (s.cpu().double() - torch.sin(x64)).abs().max().item()
    ec = (c.cpu().double() - torch.cos(x64)).abs().max().item()
    assert es <= tol and ec <= tol, (mode, es, ec)
    # beyond the fast range both fall back to libdevice
    big = torch.tensor([1e4, -3.3e5, 1.2345e7], device='cuda')
    sb, cb = torch.empty_like(big), torch.empty_like(big)
    _native.check(_native.lib().nerfattn_debug_sincos(big.data_ptr(), sb.data_ptr(), cb.data_ptr(), 3, mode,
                                                      _native.stream_handle()), 'nerfattn_debug_sincos')
    torch.cuda.synchronize()
    assert (sb.cpu().double() - torch.sin(big.cpu().double())).abs().max().item() <= 2e-7


@pytest.mark.parametrize('h,l,w,n,d', [(64, 1, 30.0, 256, 64), (256, 2, 60.0, 512, 128), (128, 3, 30.0, 384, 128),
                                       (512, 2, 30.0, 256, 128)])
def test_chain_kernel_agrees_with_unfused_path(cuda_device, monkeypatch, h, l, w, n, d):
    """The fused row-tile chain (siren_chain.cuh) and the per-layer grouped GEMM kernels (siren_tc.cuh)
    are two implementations of the same bf16 training step: one-step gradients and a short
    trajectory must agree to bf16 rounding, for one and two tiles in flight per CTA and for the
    polynomial and the SFU-core sine."""
    cfg = na.SIRENConfig(h, l, w, 'kat')
    state = seeded_state(cfg, d, 77)
    kv = smooth_tensor(5, n, d)

    def run(epochs, **env):
        for k in ('NERFATTN_NO_CHAIN', 'NERFATTN_CHAIN_SLOTS', 'NERFATTN_SINCOS', 'NERFATTN_CLUSTER'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        return gpu_fit(kv, cfg, epochs, 'bf16', state, keep_optimizer_state=True)

    ref1 = run(1, NERFATTN_NO_CHAIN='1')
    g_ref = ref1.model.adam_state[0].cpu()
    variants = [{}, {"NERFATTN_CHAIN_SLOTS": "1"}, {"NERFATTN_SINCOS": "0"}, {"NERFATTN_CLUSTER": "1"}, {"NERFATTN_CLUSTER": "0"}]
    for env in variants:
        got = run(1, **env)
        assert got.losses[0] == pytest.approx(ref1.losses[0], rel=1e-3), env
        g = got.model.adam_state[0].cpu()
        assert torch.nn.functional.cosine_similarity(g, g_ref, dim=0).item() > 0.9999, env
        assert g.norm().item() == pytest.approx(g_ref.norm().item(), rel=5e-3), env
    ref = run(60, NERFATTN_NO_CHAIN='1')
    for env in variants:
        got = run(60, **env)
        assert abs(got.final_cosine_mean - ref.final_cosine_mean) <= 1e-3, env
        assert np.allclose(got.losses, ref.losses, rtol=5e-3), env
    # bitwise run-to-run determinism of the fused path
    a, b = run(20), run(20)
    assert a.losses == b.losses and torch.equal(flat(a.model.state_dict()), flat(b.model.state_dict()))


@pytest.mark.parametrize('h,w,n,d', [(64, 30.0, 256, 128), (128, 30.0, 384, 128), (64, 60.0, 200, 64), (128, 15.0, 1000, 128)])
def test_resident_kernel_agrees_with_chain_path(cuda_device, monkeypatch, h, w, n, d):
    """The fit-resident kernel (siren_resident.cuh: one persistent CTA per fit, all epochs in one launch) and the row-tile
    chain + grouped dW / Adam kernels are two implementations of the same bf16 training step for narrow one-hidden-layer
    SIRENs: one-step gradients and a short trajectory must agree to bf16 rounding, ragged lengths included."""
    cfg = na.SIRENConfig(h, 1, w, 'kat')
    state = seeded_state(cfg, d, 123)
    kv = smooth_tensor(9, n, d)

    def run(epochs, resident):
        monkeypatch.delenv('NERFATTN_NO_RESIDENT', raising=False)
        monkeypatch.delenv('NERFATTN_RESIDENT', raising=False)
        if resident:
            monkeypatch.setenv('NERFATTN_RESIDENT', '1')     # a call this small would take the chain path by itself
        else:
            monkeypatch.setenv('NERFATTN_NO_RESIDENT', '1')
        return gpu_fit(kv, cfg, epochs, 'bf16', state, keep_optimizer_state=True)

    ref1, got1 = run(1, False), run(1, True)
    assert got1.losses[0] == pytest.approx(ref1.losses[0], rel=1e-4)
    g_ref, g = ref1.model.adam_state[0].cpu(), got1.model.adam_state[0].cpu()
    assert torch.nn.functional.cosine_similarity(g, g_ref, dim=0).item() > 0.99999
    assert g.norm().item() == pytest.approx(g_ref.norm().item(), rel=1e-3)
    ref, got = run(80, False), run(80, True)
    assert np.allclose(got.losses, ref.losses, rtol=2e-3)
    assert abs(got.final_cosine_mean - ref.final_cosine_mean) <= 5e-4
    # against the oracle, and deterministic
    orc_fit = oracle_fit(kv, cfg, 80, state)
    assert abs(got.final_cosine_mean - orc_fit.final_cosine_mean) <= COS_ATOL_BF16
    again = run(80, True)
    assert again.losses == got.losses and torch.equal(flat(again.model.state_dict()), flat(got.model.state_dict()))


def test_resident_groups_beside_epoch_graphs_and_progress(cuda_device, monkeypatch):
    """A mixed sweep: tiny / small fits train in the fit-resident kernel on a side stream while the other architectures
    replay their epoch graphs; progress evaluations (siren.py:107-115) split both into the same stretches."""
    from nerf_attention.extract import synthetic_head
    monkeypatch.setenv('NERFATTN_RESIDENT', '1')             # 14 short fits: the planner alone would not pick the kernel
    keys, values = synthetic_head(3, 1, 256, 32, 8, 128)
    spec = [(t, c, seeded_state(c, 128, 70 + i)) for t in (keys, values) for i, c in enumerate(na.CONFIGS_FULL)]
    jobs = [na.FitJob(t, c, model_from_state(c, 128, s)) for t, c, s in spec]
    res = na.fit_many(jobs, epochs=90, device='cuda', verbose=False, precision='bf16', log_every=40, progress=True)
    ref = na.fit_many([na.FitJob(t, c, model_from_state(c, 128, s)) for t, c, s in spec], epochs=90, device='cuda',
                      verbose=False, precision='fp32', log_every=40, progress=True)
    for x, y in zip(res, ref):
        assert abs(x.final_cosine_mean - y.final_cosine_mean) <= COS_ATOL_BF16, x.config.name
        assert [p[0] for p in x.progress] == [40, 80] and len(x.losses) == 90
        for (e, nm, rm, cs), (_, nm2, rm2, cs2) in zip(x.progress, y.progress):
            assert nm == pytest.approx(nm2, rel=3e-2) and cs == pytest.approx(cs2, abs=COS_ATOL_BF16)
