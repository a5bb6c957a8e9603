"""Two-rank run of the sweep CLI on real GPUs (skipped on a one-GPU box): every rank reads only its layers' files,
fits its shard, one all-gather, rank 0 writes fit_results.json."""
import json
import os
import subprocess
import sys

import pytest
import torch

import nerf_attention as na

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_sweep_cli(cuda_device, tmp_path):
    kv_dir, out_dir = tmp_path / 'kv', tmp_path / 'fits'
    na.extract_kv_cache_synthetic(seq_len=256, num_layers=8, num_kv_heads=4, head_dim=128, output_dir=kv_dir,
                                  device='cuda')
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, 'nerf-attention_b200'))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29541', '-m', 'nerf_attention.fit',
           '--kv_dir', str(kv_dir), '--output_dir', str(out_dir), '--epochs', '60', '--precision', 'bf16']
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    records = json.loads((out_dir / 'fit_results.json').read_text())
    assert len(records) == 5 * 4 * 2 * 7                          # layers {0,2,4,6,7} x 4 heads x K/V x 7 configs
    assert [r['name'] for r in records][:2] == ['L0_H0_key_tiny', 'L0_H0_key_small']
    assert all(0.0 < r['final_cosine_mean'] <= 1.0 and r['seq_len'] == 256 for r in records)
    assert len(list(out_dir.glob('*_medium_model.pt'))) == 5 * 4 * 2
