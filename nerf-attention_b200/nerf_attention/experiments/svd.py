"""Experiment 3: truncated-SVD baseline at matched compression ratios (reference experiments/svd.py:19-85).

Same records and ``svd_results.json`` as the reference.  The factorisation runs once per tensor on ``device``
(the reference repeats the same CPU SVD for every target ratio); library code (torch.linalg), SURVEY.md 8f-4.
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from nerf_attention.types import KVMetadata

# key order of one record, reference svd.py:61-77
RECORD_KEYS = ('name', 'method', 'layer', 'head', 'kv_type', 'rank', 'target_compression', 'actual_compression',
               'final_cosine_mean', 'final_cosine_min', 'final_cosine_std', 'raw_size_bytes', 'svd_size_bytes',
               'seq_len', 'd_head')


def matched_rank(seq_len: int, d_head: int, target_cr: float) -> int:
    """Largest rank whose fp32 factors (U_r, S_r, V_r) fit ``raw fp16 bytes / target_cr`` (svd.py:49-52)."""
    raw_bytes = seq_len * d_head * 2
    rank = max(1, int(raw_bytes / (target_cr * 4 * (seq_len + 1 + d_head))))
    return min(rank, seq_len, d_head)


def _row_cosines(approx: torch.Tensor, exact: torch.Tensor) -> torch.Tensor:
    """F.cosine_similarity(dim=1) with its eps = 1e-8 clamp on each norm."""
    num = (approx * exact).sum(dim=1)
    return num / (approx.norm(dim=1).clamp_min(1e-8) * exact.norm(dim=1).clamp_min(1e-8))


def run_svd_experiment(kv_dir: Path, base_dir: Path, target_compressions: list[float] | None = None,
                       device=None) -> list[dict]:
    kv_dir, base_dir = Path(kv_dir), Path(base_dir)
    base_dir.mkdir(parents=True, exist_ok=True)
    targets = [2.0, 4.0, 8.0, 16.0] if target_compressions is None else list(target_compressions)
    dev = torch.device(device if device is not None else ('cuda' if torch.cuda.is_available() else 'cpu'))
    metadata = KVMetadata.from_dict(json.loads((kv_dir / 'metadata.json').read_text()))

    records: list[dict] = []
    for layer_idx in sorted({0, metadata.num_layers // 2, metadata.num_layers - 1}):
        path = kv_dir / f'layer_{layer_idx:02d}.pt'
        if not path.exists():
            continue
        blob = torch.load(path, map_location='cpu', weights_only=True)
        for head_idx in range(min(metadata.num_kv_heads, 4)):
            for kv_type, key in (('key', 'keys'), ('value', 'values')):
                exact = blob[key][head_idx].to(dev, torch.float64)
                seq_len, d_head = exact.shape
                raw_bytes = seq_len * d_head * 2                      # the KV cache is float16
                U, S, Vt = torch.linalg.svd(exact, full_matrices=False)
                shown = []
                for target_cr in targets:
                    rank = matched_rank(seq_len, d_head, target_cr)
                    cos = _row_cosines((U[:, :rank] * S[:rank]) @ Vt[:rank], exact)
                    svd_bytes = (seq_len * rank + rank + rank * d_head) * 4
                    values = (f'L{layer_idx}_H{head_idx}_{kv_type}_svd_r{rank}', 'svd', layer_idx, head_idx, kv_type, rank,
                              target_cr, float(raw_bytes / svd_bytes), float(cos.mean()), float(cos.min()),
                              float(cos.std()), raw_bytes, svd_bytes, seq_len, d_head)
                    records.append(dict(zip(RECORD_KEYS, values)))
                    shown.append(f"r{rank}={records[-1]['final_cosine_mean']:.4f}@{records[-1]['actual_compression']:.1f}x")
                print(f"  L{layer_idx}_H{head_idx}_{kv_type}: " + " | ".join(shown))

    (base_dir / 'svd_results.json').write_text(json.dumps(records, indent=2))
    print("\nSVD Summary:")
    for tc in targets:
        means = {kind: [r['final_cosine_mean'] for r in records if r['kv_type'] == kind and r['target_compression'] == tc]
                 for kind in ('key', 'value')}
        if means['key'] and means['value']:
            print(f"  {tc:.0f}x: keys CosSim={np.mean(means['key']):.4f}, values CosSim={np.mean(means['value']):.4f}")
    return records
