"""Helpers shared by the GPU parity tests."""

import numpy as np
import torch

import nerf_attention as na
from oracle import siren_oracle as orc


def model_from_state(cfg, d, state):
    model = na.SIREN(cfg, out_features=d)
    model.load_state_dict(state)
    return model


def seeded_state(cfg, d, seed):
    torch.manual_seed(seed)
    return orc.init_state(cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, d)


def flat(state):
    return torch.cat([v.reshape(-1) for v in state.values()])


def smooth_tensor(seed, n, d):
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(0, 1, n).unsqueeze(1)
    f = torch.rand(1, d, generator=g) * 6 + 1
    ph = torch.rand(1, d, generator=g) * 6.28
    return (0.7 * torch.sin(6.2831853 * f * t + ph) + 0.2 * torch.randn(n, d, generator=g)
            + 0.3 * torch.rand(1, d, generator=g))


def gpu_fit(kv, cfg, epochs, precision, state, **kw):
    job = na.FitJob(kv, cfg, model_from_state(cfg, kv.shape[1], state))
    return na.fit_many([job], epochs=epochs, device='cuda', verbose=False, precision=precision, **kw)[0]


def oracle_fit(kv, cfg, epochs, state):
    return orc.fit(kv, cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, epochs=epochs, lr=1e-4,
                   device='cpu', log_every=10 ** 9, init={k: v.clone() for k, v in state.items()})


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
