"""single-query attention from key/value SIRENs vs from an fp16 KV cache in HBM, per launch of
`heads` heads.  Writes profiles/attention_table_<tag>.json and prints a markdown table."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200'))
import torch
import nerf_attention as na
from nerf_attention.evaluate import profile_attention

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 128
out = {}
for name in ('medium', 'tiny'):
    cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
    torch.manual_seed(0)
    models = [na.SIREN(cfg, 128) for _ in range(4)]
    out[name] = profile_attention(models, [512, 2048, 8192, 32768], heads_per_launch=heads)
    print(f'\n### {name} (H={cfg.hidden_features}, L={cfg.hidden_layers}), {heads} heads per launch')
    print('| tokens | KV attention us | KV GB/s | SIREN attention bf16 us | SIREN / KV |')
    print('|---|---|---|---|---|')
    for r in out[name]:
        print(f"| {r['seq_len']} | {r['kvread_attention_us']:.1f} | {r['kvread_gbs']:.0f} | "
              f"{r['siren_attention_bf16_us']:.1f} | {r['siren_over_kvread']:.1f}x |")
(Path(__file__).parent / f'attention_table_{tag}.json').write_text(json.dumps(out, indent=1))
