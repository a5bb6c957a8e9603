"""Record types of the SIREN fitting path.

Field names, order and defaults follow the reference so that code written
against ``nerf_attention.types`` keeps working unchanged
(reference: nerf_attention/types.py:10-37 for SIRENConfig/FitResult,
:40-63 for KVMetadata, :87-100 for the two architecture tables).
Only the types the hot path touches are provided here; the analysis-only
records (LayerSummary, AnalysisResult) belong to out-of-scope components.
"""

from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch
import torch.nn as nn


@dataclass(frozen=True)
class SIRENConfig:
    """Architecture of one SIREN: 1 -> H sine, L x (H -> H) sine, H -> D linear."""

    hidden_features: int = 256   # H
    hidden_layers: int = 2       # L, the H->H sine layers (first layer not counted)
    omega_0: float = 30.0
    name: str = 'medium'

    # -- helpers used by the native path (not part of the reference surface) --
    def param_count(self, out_features: int) -> int:
        h, l = self.hidden_features, self.hidden_layers
        return 2 * h + l * (h * h + h) + (h * out_features + out_features)

    def flops_per_epoch(self, seq_len: int, out_features: int) -> int:
        """Algorithmic GEMM work of one full-batch training step."""
        h, l = self.hidden_features, self.hidden_layers
        return 6 * seq_len * (l * h * h + h * out_features) + 4 * seq_len * h


@dataclass
class FitResult:
    model: nn.Module
    config: SIRENConfig
    target_mean: torch.Tensor
    target_std: torch.Tensor
    losses: list[float]
    final_mse: float
    final_cosine_mean: float
    final_cosine_min: float
    final_cosine_std: float
    per_pos_mse: np.ndarray
    cosine_sims: np.ndarray
    compression_ratio: float
    raw_size_bytes: int
    siren_size_bytes: int
    train_time_seconds: float
    seq_len: int
    d_head: int
    num_parameters: int


@dataclass
class KVMetadata:
    model_name: str
    num_layers: int
    num_kv_heads: int
    seq_len: int
    head_dim: int
    actual_tokens: int
    dtype: str = 'float32'   # dtype of the layer_XX.pt files, not of the live KV cache

    def to_dict(self) -> dict[str, Any]:
        return dataclasses.asdict(self)

    @classmethod
    def from_dict(cls, d: dict[str, Any]) -> 'KVMetadata':
        known = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in known})


def __getattr__(name: str):
    # LayerSummary / AnalysisResult (reference types.py:66-84) live next to the analysis that fills them
    if name in ('LayerSummary', 'AnalysisResult'):
        from nerf_attention import analyze
        return getattr(analyze, name)
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')


def _table(rows: list[tuple[int, int, float, str]]) -> list[SIRENConfig]:
    return [SIRENConfig(h, l, w, n) for h, l, w, n in rows]


# quickstart / --quick sweep
CONFIGS_QUICK: list[SIRENConfig] = _table([
    (128, 1, 30.0, 'small'),
    (256, 2, 30.0, 'medium'),
])

# the 7-architecture sweep
CONFIGS_FULL: list[SIRENConfig] = _table([
    (64, 1, 30.0, 'tiny'),
    (128, 1, 30.0, 'small'),
    (256, 2, 30.0, 'medium'),
    (512, 2, 30.0, 'large'),
    (256, 3, 30.0, 'deep'),
    (256, 2, 60.0, 'hifreq'),
    (256, 2, 15.0, 'lofreq'),
])
