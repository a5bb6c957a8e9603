"""SIREN model and the fit entry point, API-compatible with the reference
(nerf_attention/siren.py): ``SineLayer``, ``SIREN``, ``fit_siren``.

The modules are real ``nn.Module``s built from ``nn.Linear`` on the CPU so that a
seeded construction draws exactly the numbers the reference draws and the
``state_dict`` keys are identical (``network.{i}.linear.weight`` ...).  Training
does not go through autograd: ``fit_siren`` hands the packed weights to the
sm_100a kernels (nerf_attention.batched.fit_many) and loads the result back.
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn

from nerf_attention.types import FitResult, SIRENConfig


class SineLayer(nn.Module):
    """y = sin(omega_0 * (x W^T + b)); reference nerf_attention/siren.py:17-34."""

    def __init__(self, in_features: int, out_features: int,
                 omega_0: float = 30.0, is_first: bool = False):
        super().__init__()
        self.omega_0 = omega_0
        self.linear = nn.Linear(in_features, out_features)
        # first layer U(+-1/in); deeper layers U(+-sqrt(6/in)/omega_0); the bias gets the same
        # bound as the weight (siren.py:25-31)
        limit = 1.0 / in_features if is_first else math.sqrt(6.0 / in_features) / omega_0
        _siren_uniform_(self.linear, limit)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.sin(self.omega_0 * self.linear(x))


def _siren_uniform_(linear: nn.Linear, limit: float) -> None:
    with torch.no_grad():
        linear.weight.uniform_(-limit, limit)
        linear.bias.uniform_(-limit, limit)


class SIREN(nn.Module):
    """1 -> H sine, L x (H -> H) sine, H -> out_features linear (siren.py:37-67)."""

    def __init__(self, config: SIRENConfig, out_features: int):
        super().__init__()
        self.siren_config = config
        h, w0 = config.hidden_features, config.omega_0
        stack: list[nn.Module] = [SineLayer(1, h, omega_0=w0, is_first=True)]
        stack += [SineLayer(h, h, omega_0=w0) for _ in range(config.hidden_layers)]
        head = nn.Linear(h, out_features)
        _siren_uniform_(head, math.sqrt(6.0 / h) / w0)
        stack.append(head)
        self.network = nn.Sequential(*stack)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.network(x)

    def count_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def size_bytes(self) -> int:
        return 4 * self.count_parameters()     # fp32 weights

    # ---- packing for the native path: state_dict order == include/nerfattn.h params layout
    def packed_parameters(self) -> list[torch.nn.Parameter]:
        out = []
        for layer in self.network:
            lin = layer.linear if isinstance(layer, SineLayer) else layer
            out += [lin.weight, lin.bias]
        return out


def fit_siren(
    kv_tensor: torch.Tensor,
    config: SIRENConfig,
    epochs: int = 5000,
    lr: float = 1e-4,
    device: str = 'cuda',
    log_every: int = 500,
    verbose: bool = True,
    precision: str | None = None,
) -> FitResult:
    """Fit one SIREN to one (seq_len, d_head) KV tensor (reference siren.py:70-149).

    Same signature and return record as the reference; ``precision`` ('fp32' |
    'bf16', default $NERFATTN_PRECISION or 'fp32') is the only addition.  The
    caller's tensor is never modified.
    """
    from nerf_attention.batched import FitJob, fit_many
    return fit_many([FitJob(kv_tensor, config)], epochs=epochs, lr=lr, device=device,
                    log_every=log_every, verbose=verbose, precision=precision)[0]
