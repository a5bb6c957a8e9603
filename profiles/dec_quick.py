import sys
sys.path.insert(0, 'nerf-attention_b200')
import torch, nerf_attention as na
from nerf_attention.evaluate import profile_decode
for name in ('medium', 'tiny', 'deep'):
    cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
    torch.manual_seed(0)
    models = [na.SIREN(cfg, 128) for _ in range(4)]
    for r in profile_decode(models, [2048, 8192], heads_per_launch=256, warmup=3, runs=10):
        print(name, r['seq_len'], 'bf16 us', round(r['siren_bf16_us'], 1), 'TF', round(r['siren_bf16_tflops']))
