"""Experiment 3: truncated-SVD baseline at matched compression ratios (reference experiments/svd.py:19-85).

Same records and ``svd_results.json`` as the reference.  The factorisation runs once per tensor on ``device``
(the reference repeats the same CPU SVD for every target ratio); library code (torch.linalg), SURVEY.md 8f-4.
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from nerf_attention.types import KVMetadata


def run_svd_experiment(kv_dir: Path, base_dir: Path, target_compressions: list[float] | None = None,
                       device=None) -> list[dict]:
    kv_dir, base_dir = Path(kv_dir), Path(base_dir)
    base_dir.mkdir(parents=True, exist_ok=True)
    if target_compressions is None:
        target_compressions = [2.0, 4.0, 8.0, 16.0]
    dev = torch.device(device if device is not None else ('cuda' if torch.cuda.is_available() else 'cpu'))
    with open(kv_dir / 'metadata.json') as f:
        metadata = KVMetadata.from_dict(json.load(f))

    all_results: list[dict] = []
    for layer_idx in sorted({0, metadata.num_layers // 2, metadata.num_layers - 1}):
        filepath = kv_dir / f'layer_{layer_idx:02d}.pt'
        if not filepath.exists():
            continue
        data = torch.load(filepath, map_location='cpu', weights_only=True)
        for head_idx in range(min(metadata.num_kv_heads, 4)):
            for kv_type, tensor in (('key', data['keys'][head_idx]), ('value', data['values'][head_idx])):
                seq_len, d_head = tensor.shape
                raw_bytes = seq_len * d_head * 2                      # KV cache is float16
                t = tensor.to(dev, torch.float64)
                U, S, Vt = torch.linalg.svd(t, full_matrices=False)
                first = len(all_results)
                for target_cr in target_compressions:
                    # svd_bytes = (seq_len * rank + rank + rank * d_head) * 4
                    rank = max(1, int(raw_bytes / (target_cr * 4 * (seq_len + 1 + d_head))))
                    rank = min(rank, min(seq_len, d_head))
                    reconstructed = (U[:, :rank] * S[:rank]) @ Vt[:rank, :]
                    svd_bytes = (seq_len * rank + rank + rank * d_head) * 4
                    cos_sim = F.cosine_similarity(reconstructed, t, dim=1)
                    all_results.append({
                        'name': f'L{layer_idx}_H{head_idx}_{kv_type}_svd_r{rank}', 'method': 'svd',
                        'layer': layer_idx, 'head': head_idx, 'kv_type': kv_type, 'rank': rank,
                        'target_compression': target_cr, 'actual_compression': float(raw_bytes / svd_bytes),
                        'final_cosine_mean': float(cos_sim.mean().item()),
                        'final_cosine_min': float(cos_sim.min().item()),
                        'final_cosine_std': float(cos_sim.std().item()),
                        'raw_size_bytes': raw_bytes, 'svd_size_bytes': svd_bytes, 'seq_len': seq_len, 'd_head': d_head,
                    })
                print(f"  L{layer_idx}_H{head_idx}_{kv_type}: "
                      + " | ".join(f"r{r['rank']}={r['final_cosine_mean']:.4f}@{r['actual_compression']:.1f}x"
                                   for r in all_results[first:]))

    with open(base_dir / 'svd_results.json', 'w') as f:
        json.dump(all_results, f, indent=2)
    print("\nSVD Summary:")
    for tc in target_compressions:
        kr = [r for r in all_results if r['kv_type'] == 'key' and r['target_compression'] == tc]
        vr = [r for r in all_results if r['kv_type'] == 'value' and r['target_compression'] == tc]
        if kr and vr:
            print(f"  {tc:.0f}x: keys CosSim={np.mean([r['final_cosine_mean'] for r in kr]):.4f}, "
                  f"values CosSim={np.mean([r['final_cosine_mean'] for r in vr]):.4f}")
    return all_results
