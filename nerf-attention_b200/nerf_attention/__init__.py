"""B200-native drop-in for the SIREN fit / reconstruction path of ruskaruma/nerf-attention.

Same import surface as the reference package (reference nerf_attention/__init__.py:1-21).  Limits: there is no CPU
path (``device='cpu'`` raises; the reference falls back to the CPU); ``extract_kv_cache`` (real-LLM extraction) and the
three figure functions ``plot_pareto_frontier`` / ``plot_keys_vs_values`` / ``generate_summary_figure`` import but raise
NotImplementedError -- they are outside the hot path ("out of scope").
"""

from nerf_attention.types import (
    CONFIGS_FULL,
    CONFIGS_QUICK,
    AnalysisResult,
    FitResult,
    KVMetadata,
    LayerSummary,
    SIRENConfig,
)
from nerf_attention.analyze import analyze_kv_cache
from nerf_attention.siren import SIREN, SineLayer, fit_siren
from nerf_attention.batched import FitJob, fit_many
from nerf_attention.extract import extract_kv_cache, extract_kv_cache_synthetic
from nerf_attention.fit import fit_kv_cache
from nerf_attention.evaluate import (
    generate_summary_figure,
    load_results,
    per_position_cosine,
    plot_keys_vs_values,
    plot_pareto_frontier,
    plot_per_position_error,
    profile_decode,
    profile_latency,
)

__all__ = [
    'CONFIGS_FULL', 'CONFIGS_QUICK', 'AnalysisResult', 'FitResult', 'KVMetadata', 'LayerSummary', 'SIRENConfig',
    'analyze_kv_cache',
    'SIREN', 'SineLayer', 'fit_siren', 'FitJob', 'fit_many',
    'extract_kv_cache', 'extract_kv_cache_synthetic', 'fit_kv_cache',
    'load_results', 'per_position_cosine', 'plot_per_position_error', 'profile_decode',
    'profile_latency', 'plot_pareto_frontier', 'plot_keys_vs_values', 'generate_summary_figure',
]
