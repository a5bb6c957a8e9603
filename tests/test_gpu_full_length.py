"""Parity at the benched configuration (BASELINE configs 2-3), un-gated: every architecture of the sweep x
{key, value} on synthetic layer 16 / head 0, [2048 x 128], **2000 epochs**, both precisions, through the C ABI.

Gates (BASELINE.json north_star): final per-fit CosSim of the fp32 mode within 1e-3 and of the bf16 mode within
5e-3 of the reference.  The reference here is the oracle (reference siren.py:70-149 restated) run in torch eager
fp32 on the same GPU -- the reference's own default device path (siren.py:76,89), TF32 off -- because 14 full-length
fits on the CPU would take minutes; two cases are anchored against the CPU oracle as well.
Initial weights follow the bench's seed convention: manual_seed(1000 layer + 100 head + 10 is_value + ci),
so these are the same 14 fits bench.py samples for `quality.max_abs_cos_diff`.
"""

import pytest
import torch

import nerf_attention as na
from oracle import siren_oracle as orc
from gpu_util import model_from_state

pytestmark = pytest.mark.gpu

LAYER, HEAD, N, D, EPOCHS = 16, 0, 2048, 128, 2000
COS_ATOL = {'fp32': 1e-3, 'bf16': 5e-3}
CPU_ANCHORS = (('key', 'tiny'), ('key', 'hifreq'))            # hifreq: omega_0 = 60, the MUFU-core sine's worst case


def _specs():
    from nerf_attention.extract import synthetic_head
    keys, values = synthetic_head(LAYER, HEAD, N, 32, 8, D)
    out = []
    for is_value, (kv_name, kv) in enumerate((('key', keys), ('value', values))):
        for ci, cfg in enumerate(na.CONFIGS_FULL):
            torch.manual_seed(1000 * LAYER + 100 * HEAD + 10 * is_value + ci)
            out.append({'kv_name': kv_name, 'kv': kv, 'cfg': cfg,
                        'state': orc.init_state(cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, D)})
    return out


def _oracle(spec, device):
    cfg = spec['cfg']
    return orc.fit(spec['kv'], cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, epochs=EPOCHS, lr=1e-4,
                   device=device, log_every=10 ** 9, init={k: v.clone() for k, v in spec['state'].items()})


@pytest.fixture(scope='module')
def runs(cuda_device):
    specs = _specs()
    out = {'specs': specs}
    # the benched sweep trains `tiny` / `small` in the fit-resident kernel; a 14-fit call is too small for the planner to
    # pick it by itself (nerfattn.cu, resident_pays), so it is forced here: this is the kernel mix of the timed run
    import os
    saved = os.environ.get('NERFATTN_RESIDENT')
    os.environ['NERFATTN_RESIDENT'] = '1'
    try:
        for prec in ('fp32', 'bf16'):                           # all 14 fits in one batched call per precision
            jobs = [na.FitJob(s['kv'], s['cfg'], model_from_state(s['cfg'], D, s['state'])) for s in specs]
            out[prec] = na.fit_many(jobs, epochs=EPOCHS, device='cuda', verbose=False, precision=prec)
    finally:
        if saved is None:
            os.environ.pop('NERFATTN_RESIDENT', None)
        else:
            os.environ['NERFATTN_RESIDENT'] = saved
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False               # the reference never enables TF32
    try:
        out['oracle_cuda'] = [_oracle(s, 'cuda') for s in specs]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_final_cossim_matches_reference_at_benched_config(runs, prec):
    rows, worst = [], 0.0
    for spec, got, ref in zip(runs['specs'], runs[prec], runs['oracle_cuda']):
        diff = abs(got.final_cosine_mean - ref.final_cosine_mean)
        worst = max(worst, diff)
        rows.append(f"  {spec['kv_name']:<5} {spec['cfg'].name:<7} oracle {ref.final_cosine_mean:.6f} "
                    f"{prec} {got.final_cosine_mean:.6f} |diff| {diff:.2e} loss {got.losses[-1]:.5f} / {ref.losses[-1]:.5f}")
    print(f'\nfull-length parity ({prec}, N={N}, {EPOCHS} epochs), max |dCosSim| = {worst:.2e}\n' + '\n'.join(rows))
    for spec, got, ref in zip(runs['specs'], runs[prec], runs['oracle_cuda']):
        assert abs(got.final_cosine_mean - ref.final_cosine_mean) <= COS_ATOL[prec], (spec['kv_name'], spec['cfg'].name)
        assert got.losses[-1] == pytest.approx(ref.losses[-1], rel=0.1), (spec['kv_name'], spec['cfg'].name)
        assert len(got.losses) == EPOCHS and got.losses[-1] < got.losses[0]


@pytest.mark.parametrize('kv_name,arch', CPU_ANCHORS)
def test_cpu_oracle_anchor(runs, kv_name, arch):
    """The same gates against the oracle on the CPU (what the reference computes with --device cpu), and the two
    oracle devices against each other: the GPU-run oracle used above is the same function."""
    i = next(k for k, s in enumerate(runs['specs']) if s['kv_name'] == kv_name and s['cfg'].name == arch)
    cpu = _oracle(runs['specs'][i], 'cpu')
    print(f"\n{kv_name} {arch}: oracle cpu {cpu.final_cosine_mean:.6f} cuda {runs['oracle_cuda'][i].final_cosine_mean:.6f} "
          f"fp32 {runs['fp32'][i].final_cosine_mean:.6f} bf16 {runs['bf16'][i].final_cosine_mean:.6f}")
    assert abs(runs['oracle_cuda'][i].final_cosine_mean - cpu.final_cosine_mean) <= 1e-3
    for prec in ('fp32', 'bf16'):
        assert abs(runs[prec][i].final_cosine_mean - cpu.final_cosine_mean) <= COS_ATOL[prec], prec
