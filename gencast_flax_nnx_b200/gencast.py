"""`gencast.GenCast` of the reference: the top-level module that owns the denoiser
and the sampler (gencast/gencast.py:130-294)."""
from __future__ import annotations

from typing import Optional

from .configs import (DenoiserArchitectureConfig, NoiseConfig, NoiseEncoderConfig, SamplerConfig, TaskConfig,
                      num_outputs)
from .denoiser import Denoiser
from .dpm_solver_plus_plus_2s import Sampler
from .rngs import Rngs
from .xarray_lite import Dataset


class GenCast:
    """Reference: gencast/gencast.py:145-185 (constructor), :289-294 (full_sampling)."""

    def __init__(self, task_config: TaskConfig, denoiser_architecture_config: DenoiserArchitectureConfig,
                 sampler_config: Optional[SamplerConfig] = None, noise_config: Optional[NoiseConfig] = None,
                 noise_encoder_config: Optional[NoiseEncoderConfig] = None, gpu_mesh=None,
                 rngs: Optional[Rngs] = None, **denoiser_kwargs):
        self.rngs = rngs if rngs is not None else Rngs(0)
        self._task_config = task_config
        # Output size is set from the task, as in the reference (:158-169).
        denoiser_architecture_config.node_output_size = num_outputs(task_config)
        self.denoiser = Denoiser(noise_encoder_config, denoiser_architecture_config, rngs=self.rngs,
                                 gpu_mesh=gpu_mesh, **denoiser_kwargs)
        self._sampler_config = sampler_config
        self._noise_config = noise_config
        self._sampler = None
        if sampler_config is not None:
            sc = sampler_config
            self._sampler = Sampler(self.denoiser, sc.max_noise_level, sc.min_noise_level, sc.num_noise_levels,
                                    sc.rho, sc.stochastic_churn_rate, sc.churn_min_noise_level,
                                    sc.churn_max_noise_level, sc.noise_level_inflation_factor)

    def full_sampling(self, inputs: Dataset, targets_template: Dataset, forcings: Optional[Dataset] = None,
                      **kwargs) -> Dataset:
        if self._sampler is None:
            raise ValueError("Sampler config must be specified to run inference.")   # gencast.py:290-291
        return self._sampler(inputs, targets_template, forcings, rngs=self.rngs, **kwargs)

    def loss(self, *args, **kwargs):
        raise NotImplementedError("training (GenCast.loss, gencast/gencast.py:229-280) is outside the "
                                  "accelerated path; see DESIGN.md")
