"""Structure analysis and SVD baseline against the reference's own outputs
(tests/golden/analysis.json, written by tests/golden/make_golden_analysis.py from /root/reference)."""
import contextlib
import io
import json
import os

import numpy as np
import pytest
import torch

import nerf_attention as na
from nerf_attention.analyze import analyze_kv_cache, analyze_tensor
from nerf_attention.experiments.svd import run_svd_experiment

GOLDEN = json.loads(open(os.path.join(os.path.dirname(__file__), 'golden', 'analysis.json')).read())
# the reference works in float32 numpy / LAPACK on the CPU, this build in float64 torch on the device
RTOL = 2e-4


def close(a, b):
    return abs(a - b) <= RTOL * max(1.0, abs(b))


def run_all(tmp_path, device):
    kv_dir = tmp_path / 'kv'
    with contextlib.redirect_stdout(io.StringIO()):
        na.extract_kv_cache_synthetic(output_dir=kv_dir, **GOLDEN['shape'])      # bit-identical to the reference generator
        result = analyze_kv_cache(kv_dir, tmp_path / 'analysis', device=device)
        svd = run_svd_experiment(kv_dir, tmp_path / 'svd', device=device)
    written = json.loads((tmp_path / 'analysis' / 'analysis_results.json').read_text())
    ref = GOLDEN['analysis_results']
    assert written['metadata'] == {**ref['metadata'], **{k: v for k, v in written['metadata'].items() if k not in ref['metadata']}}
    assert [s['layer'] for s in written['layer_summaries']] == [s['layer'] for s in ref['layer_summaries']]
    for mine, theirs in zip(written['layer_summaries'], ref['layer_summaries']):
        assert all(close(mine[k], theirs[k]) for k in theirs), (mine, theirs)
    assert all(close(written['assessment'][k], v) for k, v in ref['assessment'].items())
    assert close(result.avg_autocorr_keys, ref['assessment']['avg_autocorr_keys'])
    assert close(result.avg_spectral_values, ref['assessment']['avg_spectral_values'])

    blob = torch.load(kv_dir / 'layer_02.pt', weights_only=True)
    for tag, tensor in (('K', blob['keys'][1]), ('V', blob['values'][1])):
        mine, theirs = analyze_tensor(tensor, f'L2_H1_{tag}', device=device), GOLDEN[f'tensor_L2_H1_{tag}']
        assert mine['name'] == theirs['name'] and mine['shape'] == theirs['shape']
        assert close(mine['lag1_autocorrelation'], theirs['lag1_autocorrelation'])
        assert np.allclose(mine['mean_autocorrelation'], theirs['mean_autocorrelation'], atol=RTOL)
        assert all(close(mine['spectral_energy'][k], v) for k, v in theirs['spectral_energy'].items())
        assert mine['rank']['full_rank'] == theirs['rank']['full_rank']
        assert abs(mine['rank']['effective_rank_99'] - theirs['rank']['effective_rank_99']) <= 1
        assert close(mine['rank']['top_sv_fraction'], theirs['rank']['top_sv_fraction'])
        assert close(mine['rank']['top_10_sv_fraction'], theirs['rank']['top_10_sv_fraction'])

    ref_svd = GOLDEN['svd_results']
    assert json.loads((tmp_path / 'svd' / 'svd_results.json').read_text()) == svd
    assert [r['name'] for r in svd] == [r['name'] for r in ref_svd]
    for mine, theirs in zip(svd, ref_svd):
        assert list(mine) == list(theirs)                                        # same keys, same order
        for k, v in theirs.items():
            assert (close(mine[k], v) if isinstance(v, float) else mine[k] == v), (k, mine[k], v)


def test_analysis_and_svd_match_reference_cpu(tmp_path):
    run_all(tmp_path, 'cpu')


@pytest.mark.gpu
def test_analysis_and_svd_match_reference_cuda(cuda_device, tmp_path):
    run_all(tmp_path, 'cuda')


def test_constant_and_short_signals():
    flat = analyze_tensor(torch.ones(8, 4), 'flat', device='cpu')
    assert flat['lag1_autocorrelation'] == 0.0 and flat['spectral_energy']['top_10pct'] == 1.0
    assert len(flat['mean_autocorrelation']) == 51                               # lags beyond the sequence stay zero
