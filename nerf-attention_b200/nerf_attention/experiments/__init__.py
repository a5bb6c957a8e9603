"""Follow-up experiment drivers of the reference that only ever call ``fit_siren``
(reference nerf_attention/experiments/scaling.py), re-pointed at the batched B200 path."""

from nerf_attention.experiments.scaling import (
    crossover_data,
    run_full_layer_profile,
    run_scaling_experiment,
)
from nerf_attention.experiments.svd import run_svd_experiment

__all__ = ['run_scaling_experiment', 'run_full_layer_profile', 'crossover_data', 'run_svd_experiment']
