"""BASELINE config 4: SIREN position->KV decode fused with q.k vs a measured HBM KV-read, latency table.
Writes profiles/decode_table_<tag>.json and prints a markdown table.  Both sides are batched over
`heads` (layer, head) pairs per launch and timed with CUDA events (one head is far below launch latency)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200'))
import torch
import nerf_attention as na
from nerf_attention.evaluate import profile_decode

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 256
out = {}
for name in ('medium', 'tiny', 'large'):
    cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
    torch.manual_seed(0)
    models = [na.SIREN(cfg, 128) for _ in range(4)]
    seqs = [512, 1024, 2048, 4096, 8192, 16384, 32768]
    out[name] = profile_decode(models, seqs, heads_per_launch=heads if name != 'large' else heads // 2, warmup=5, runs=30)
    print(f'\n### {name} (H={cfg.hidden_features}, L={cfg.hidden_layers}), {heads} heads per launch')
    print('| tokens | KV-read us | KV-read GB/s | SIREN fp32 us | SIREN bf16 us | bf16 TFLOP/s | bf16 / KV-read | ns/token/head KV | ns/token/head SIREN bf16 |')
    print('|---|---|---|---|---|---|---|---|---|')
    for r in out[name]:
        print(f"| {r['seq_len']} | {r['kvread_us']:.1f} | {r['kvread_gbs']:.0f} | {r['siren_fp32_us']:.1f} | {r['siren_bf16_us']:.1f} | "
              f"{r['siren_bf16_tflops']:.0f} | {r['siren_bf16_over_kvread']:.1f}x | {1e3 * r['kvread_us_per_token_head']:.3f} | "
              f"{1e3 * r['siren_bf16_us_per_token_head']:.3f} |")
(Path(__file__).parent / f'decode_table_{tag}.json').write_text(json.dumps(out, indent=1))
