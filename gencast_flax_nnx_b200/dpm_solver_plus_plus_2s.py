"""`dpm_solver_plus_plus_2s.Sampler` of the reference, on B200 kernels.

Mirrors gencast/dpm_solver_plus_plus_2s.py:21-177: same constructor arguments,
`__call__(inputs, targets_template, forcings=None, rngs=None)` and errors.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import stacking
from .engine import SamplerEngine, noise_schedule, stochastic_churn_rate_schedule
from .xarray_lite import Dataset


class Sampler:
    def __init__(self, denoiser, max_noise_level: float, min_noise_level: float, num_noise_levels: int,
                 rho: float, stochastic_churn_rate: float, churn_min_noise_level: float,
                 churn_max_noise_level: float, noise_level_inflation_factor: float, *,
                 evaluate_discarded_call: bool = True, use_cuda_graph: bool = True,
                 initial_noise: str = "spherical"):
        self._noise_levels = noise_schedule(max_noise_level, min_noise_level, num_noise_levels, rho)
        self._stochastic_churn = stochastic_churn_rate > 0
        # The reference's loop calls utils.apply_stochastic_churn_arr, which it does not define
        # (gencast/dpm_solver_plus_plus_2s.py:131 vs gencast/samplers_utils.py:434), and both of its drivers pass
        # 0.0 (training/train.py:167, training/evaluation.py:69).  Here churn runs with the semantics of the
        # documented apply_stochastic_churn (samplers_utils.py:434-452) on the device state.
        self._per_step_churn_rates = stochastic_churn_rate_schedule(self._noise_levels, stochastic_churn_rate,
                                                                    churn_min_noise_level, churn_max_noise_level)
        self._noise_level_inflation_factor = noise_level_inflation_factor
        self._denoiser = denoiser
        self.sigma_data = 1.0
        self._evaluate_discarded_call = evaluate_discarded_call
        self._use_graph = use_cuda_graph
        if initial_noise not in ("spherical", "white"):
            raise ValueError("initial_noise must be 'spherical' (the reference's, samplers_utils.py:333-346) or 'white'")
        self._initial_noise = initial_noise
        self._noise_gen = None
        self._engine: Optional[SamplerEngine] = None

    @property
    def noise_levels(self) -> np.ndarray:
        return self._noise_levels

    def sampler_engine(self) -> SamplerEngine:
        if self._engine is None:
            self._engine = SamplerEngine(self._denoiser.engine, self._noise_levels, self._evaluate_discarded_call,
                                         churn_rates=self._per_step_churn_rates if self._stochastic_churn else None,
                                         noise_level_inflation_factor=self._noise_level_inflation_factor)
        return self._engine

    def __call__(self, inputs: Dataset, targets_template: Dataset, forcings: Optional[Dataset] = None,
                 rngs=None, *, init_noise: Optional[np.ndarray] = None) -> Dataset:
        """One 12 h step.  `init_noise` ([G, batch, n_out], unit variance) overrides the generator.

        Like the reference (samplers_utils.py:333-346) the default initial state is isotropic
        spherical-harmonic white noise (spherical_noise.py), drawn on the device from `rngs.noise()`;
        `initial_noise='white'` draws independent grid-point noise instead.
        Surface variables are carried once (the reference broadcasts them over 13 levels and
        selects level 0, :59, :93-95 -- an equivalent state).
        """
        if rngs is None:
            raise ValueError("Must pass rngs: nnx.Rngs(...) to Sampler")          # :54-55
        key = rngs.noise()
        den = self._denoiser
        engine = den._maybe_init(inputs, targets_template, forcings)
        sizes = dict(targets_template.sizes)
        sizes.setdefault("batch", 1)
        with torch.cuda.device(engine.device):
            inp, frc = den.stack_constants(inputs, forcings, sizes)
            engine.set_constant_features(den.member_major(inp), den.member_major(frc))
            res = self.sample_on_device(targets_template, key, init_noise=init_noise)
            out = res.reshape(sizes["batch"], engine.G, engine.n_out).permute(1, 0, 2)
            return den.stacker.from_nodes(out, targets_template)

    def sample_on_device(self, targets_template: Dataset, key: int, *, init_noise: Optional[np.ndarray] = None) -> torch.Tensor:
        """Runs the solver on the engine's resident constant features (set_constant_features) and returns
        the prediction as a device tensor [batch * G, n_out] fp32 (member-major blocks of grid rows), valid
        until the next call.  `key` seeds the initial noise (one value of `rngs.noise()`)."""
        engine = self._denoiser.engine
        batch = dict(targets_template.sizes).get("batch", 1)
        if engine.B != batch:
            raise ValueError(f"Sampler was initialised for batch size {engine.B}, got {batch}")
        se = self.sampler_engine()
        with torch.cuda.device(engine.device):
            if init_noise is not None:
                noise = np.ascontiguousarray(np.transpose(np.asarray(init_noise, np.float32), (1, 0, 2))).reshape(
                    batch * engine.G, engine.n_out)
            else:
                gen = torch.Generator(device=engine.device)
                gen.manual_seed(int(key) & 0x7FFFFFFFFFFFFFFF)
                if self._initial_noise == "spherical":
                    if self._noise_gen is None:
                        from .spherical_noise import SphericalNoise
                        self._noise_gen = SphericalNoise(targets_template.coords["lat"], targets_template.coords["lon"],
                                                         engine.device)
                    noise = self._noise_gen.sample_nodes(engine.n_out, members=batch, generator=gen)
                else:
                    noise = torch.randn(batch * engine.G, engine.n_out, generator=gen, device=engine.device)
            churn_noise = None
            if se.num_churn_steps > 0:
                # one fresh unit-variance draw per churned step, like the reference's spherical_white_noise_like(x, rngs)
                g2 = self._churn_generator(key, engine.device)
                if self._initial_noise == "spherical":
                    if self._noise_gen is None:
                        from .spherical_noise import SphericalNoise
                        self._noise_gen = SphericalNoise(targets_template.coords["lat"], targets_template.coords["lon"],
                                                         engine.device)
                    draws = [self._noise_gen.sample_nodes(engine.n_out, members=batch, generator=g2)
                             for _ in range(se.num_churn_steps)]
                else:
                    draws = [torch.randn(batch * engine.G, engine.n_out, generator=g2, device=engine.device)
                             for _ in range(se.num_churn_steps)]
                churn_noise = torch.stack(draws)
            return se.sample(noise, use_graph=self._use_graph, churn_noise=churn_noise)

    @staticmethod
    def _churn_generator(key: int, device) -> torch.Generator:
        g = torch.Generator(device=device)
        g.manual_seed((int(key) * 0x9E3779B1 + 1) & 0x7FFFFFFFFFFFFFFF)
        return g
