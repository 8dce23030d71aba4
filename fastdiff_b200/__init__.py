"""fastdiff_b200 -- B200-native (sm_100a) implementation of FastDiff's reverse-diffusion sampling hot path.

Public surface = the reference's own names for this path:
  FastDiff                                   modules/FastDiff/module/FastDiff_model.py:10
  sampling_given_noise_schedule, compute_hyperparams_given_schedule, map_noise_scale_to_time_step,
  calc_diffusion_step_embedding, std_normal  modules/FastDiff/module/util.py
"""
from .model import FastDiff  # noqa: F401
from .sampler import (  # noqa: F401
    calc_diffusion_step_embedding,
    compute_hyperparams_given_schedule,
    map_noise_scale_to_time_step,
    sampling_given_noise_schedule,
    std_normal,
)

__all__ = [
    "FastDiff", "sampling_given_noise_schedule", "compute_hyperparams_given_schedule",
    "map_noise_scale_to_time_step", "calc_diffusion_step_embedding", "std_normal",
]
