import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, 'nerf-attention_b200'); sys.path.insert(0, '.')
import torch
import nerf_attention as na
from gpu_util import *
from oracle import siren_oracle as orc
from nerf_attention.evaluate import PackedModels
for name, n in [('medium', 512), ('tiny', 2048), ('large', 256), ('deep', 1024), ('medium', 333)]:
    cfg = next(c for c in na.CONFIGS_FULL if c.name == name)
    heads = 3
    states = [seeded_state(cfg, 128, 70 + i) for i in range(heads)]
    g = torch.Generator().manual_seed(1)
    means = [torch.randn(1, 128, generator=g) * 0.1 for _ in range(heads)]
    stds = [torch.rand(1, 128, generator=g) + 0.5 for _ in range(heads)]
    q = torch.randn(heads, 128, generator=g).half()
    packed = PackedModels([model_from_state(cfg, 128, s) for s in states], n, means, stds)
    out = packed.decode_qk(q.cuda(), 'fp32').clone()
    for i in range(heads):
        r32 = orc.decode_scores(states[i], cfg.omega_0, means[i], stds[i], q[i], n)
        r64 = orc.decode_scores(states[i], cfg.omega_0, means[i], stds[i], q[i], n, dtype=torch.float64)
        print(name, n, i, 'gpu-vs-cpu32 %.2e  gpu-vs-f64 %.2e  cpu32-vs-f64 %.2e' % (rel_err(out[i].cpu(), r32), rel_err(out[i].cpu(), r64), rel_err(r32, r64)))
