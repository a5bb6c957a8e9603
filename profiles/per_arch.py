"""Per-architecture throughput of the batched fit (40 fits of one architecture per call, like one
group of the 280-fit sweep): fit-epochs/s and algorithmic TFLOP/s, CUDA-event timed.
usage: python profiles/per_arch.py [epochs] [precision] [nfits] [arch,arch,..]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200'))
import torch
import nerf_attention as na
from nerf_attention.extract import synthetic_head

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
prec = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
nfits = int(sys.argv[3]) if len(sys.argv) > 3 else 40
only = sys.argv[4].split(',') if len(sys.argv) > 4 else None
N, D = 2048, 128
tensors = [synthetic_head(16, h % 8, N, 32, 8, D)[h // 8 % 2] for h in range(16)]
out = []
for cfg in na.CONFIGS_FULL:
    if only and cfg.name not in only:
        continue
    H, L = cfg.hidden_features, cfg.hidden_layers
    flop = 6 * N * (L * H * H + H * D) + 4 * N * H
    torch.manual_seed(0)
    def run(ep):
        b = None
        for rep in range(2):
            jobs = [na.FitJob(tensors[i % len(tensors)], cfg) for i in range(nfits)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            na.fit_many(jobs, epochs=ep, device='cuda', verbose=False, precision=prec)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            b = dt if b is None else min(b, dt)
        return b
    # slope between two epoch counts: the call's fixed host set-up (model construction, H2D, D2H) cancels
    best = (run(4 * epochs) - run(epochs)) / 3.0
    rec = {'arch': cfg.name, 'H': H, 'L': L, 'fits': nfits, 'epochs': epochs, 'precision': prec,
           'us_per_epoch': best / epochs * 1e6, 'fit_epochs_per_sec': nfits * epochs / best,
           'tflops': nfits * epochs * flop / best / 1e12}
    out.append(rec)
    print(json.dumps(rec), flush=True)
