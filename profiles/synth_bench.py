"""GPU synthetic KV generator vs the CPU (reference-identical) generator: agreement and time.
usage: python profiles/synth_bench.py r01 > profiles/synth_r01.md"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'nerf-attention_b200'))
import numpy as np
import torch
from nerf_attention.extract import synthetic_head, synthetic_heads_cuda

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
rows = []
synthetic_heads_cuda([(0, 0)], 256, 32, 8, 128)
torch.cuda.synchronize()
for n, pairs in [(2048, [(l, h) for l in (0, 8, 16, 24, 31) for h in range(4)]), (4096, [(16, h) for h in range(8)]),
                 (32768, [(31, h) for h in range(8)])]:
    t0 = time.perf_counter()
    k, v = synthetic_heads_cuda(pairs, n, 32, 8, 128)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    ncheck = 2 if n > 4096 else 4
    t0 = time.perf_counter()
    ref = [synthetic_head(l, h, n, 32, 8, 128) for l, h in pairs[:ncheck]]
    cpu_s = (time.perf_counter() - t0) / ncheck
    kd = np.concatenate([(k[i].cpu().numpy().astype(np.float64) - ref[i][0].numpy()).ravel() for i in range(ncheck)])
    vd = np.concatenate([(v[i].cpu().numpy().astype(np.float64) - ref[i][1].numpy()).ravel() for i in range(ncheck)])
    rows.append({'seq_len': n, 'heads': len(pairs), 'gpu_ms_total': gpu_s * 1e3, 'gpu_ms_per_head': gpu_s * 1e3 / len(pairs),
                 'cpu_s_per_head': cpu_s, 'keys_max_abs_diff': float(np.abs(kd).max()), 'values_max_abs_diff': float(np.abs(vd).max()),
                 'keys_bit_identical': float((kd == 0).mean()), 'values_bit_identical': float((vd == 0).mean())})
with open(os.path.join(HERE, f'synth_{tag}.json'), 'w') as f:
    json.dump(rows, f, indent=1)
print('| tokens | heads | GPU ms (all heads) | GPU ms / head | CPU s / head | max abs diff keys / values | bit-identical keys / values |')
print('|---|---|---|---|---|---|---|')
for r in rows:
    print(f"| {r['seq_len']} | {r['heads']} | {r['gpu_ms_total']:.1f} | {r['gpu_ms_per_head']:.2f} | {r['cpu_s_per_head']:.2f} | "
          f"{r['keys_max_abs_diff']:.1e} / {r['values_max_abs_diff']:.1e} | {100 * r['keys_bit_identical']:.2f} % / {100 * r['values_bit_identical']:.2f} % |")
