"""Small single-architecture fit for ncu captures: `nfits` fits of one config, few epochs,
direct launches (NERFATTN_NO_GRAPH=1) so the kernel order per epoch is fixed:
layer0, fwd_sine x L, fwd_out, (dw, dx) x (L+1), layer0_grad, adam, tick."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200'))
os.environ.setdefault('NERFATTN_NO_GRAPH', '1')
import torch
import nerf_attention as na
from nerf_attention.extract import synthetic_head

name = sys.argv[1] if len(sys.argv) > 1 else 'medium'
nfits = int(sys.argv[2]) if len(sys.argv) > 2 else 40
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prec = sys.argv[4] if len(sys.argv) > 4 else 'bf16'
cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
tensors = [synthetic_head(16, h % 8, 2048, 32, 8, 128)[h // 8 % 2] for h in range(min(nfits, 16))]
torch.manual_seed(0)
jobs = [na.FitJob(tensors[i % len(tensors)], cfg) for i in range(nfits)]
res = na.fit_many(jobs, epochs=epochs, device='cuda', verbose=False, precision=prec)
torch.cuda.synchronize()
print('ok', name, nfits, epochs, prec, res[0].losses[-1])
