"""`denoiser.Denoiser` of the reference, on B200 kernels.

Mirrors gencast/denoiser.py:142-202 (Denoiser) and :205-341 (DenoiserArchitecture):
same constructor arguments, same call signature and error behaviour, same
Dataset in / Dataset out contract.  The network itself runs in DenoiserEngine.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import stacking
from .device_stacking import DeviceStacker
from .configs import DenoiserArchitectureConfig, NoiseEncoderConfig
from .engine import ChannelLayout, DenoiserEngine
from .graph import build_denoiser_graphs
from .params import init_perturbed, init_reference_like, param_shapes
from .xarray_lite import DataArray, Dataset


class Denoiser:
    """Wraps the GNN-transformer-GNN network with noise-level encoding.

    Reference: gencast/denoiser.py:153-159 for the constructor (`rngs` seeds the
    lazily created parameters, `gpu_mesh` is accepted and unused exactly as in the
    reference drivers, training/train_helpers.py:143-150).  Extra keyword-only
    arguments select what the reference leaves to JAX: `params` (a flat dict keyed
    by the NNX attribute paths), `compute_dtype`
    ('bf16' tensor-core path or 'f32' parity path) and `device`.
    """

    def __init__(self, noise_encoder_config: Optional[NoiseEncoderConfig],
                 denoiser_architecture_config: DenoiserArchitectureConfig, rngs=None, gpu_mesh=None, *,
                 params: Optional[Dict[str, np.ndarray]] = None, compute_dtype: str = "bf16",
                 device=None, param_init: str = "reference"):
        self._noise_cfg = noise_encoder_config or NoiseEncoderConfig()
        self._arch = denoiser_architecture_config
        self._rngs = rngs
        self._params = params
        self._compute_dtype = compute_dtype
        self._device = device
        self._param_init = param_init
        self._engine: Optional[DenoiserEngine] = None
        self._grid_key = None
        self._stacker: Optional[DeviceStacker] = None

    # -- lazy construction, as in DenoiserArchitecture._maybe_init (denoiser.py:343-416)
    def _maybe_init(self, inputs: Dataset, noisy_targets: Dataset, forcings: Optional[Dataset]) -> DenoiserEngine:
        lat, lon = np.asarray(inputs.coords["lat"]), np.asarray(inputs.coords["lon"])
        key = (lat.tobytes(), lon.tobytes())
        if self._engine is not None:
            if key != self._grid_key:
                raise ValueError("Denoiser was initialised for a different lat/lon grid")
            return self._engine
        forc = forcings if forcings is not None else Dataset({}, inputs.coords)
        # a name shared by forcings and noisy targets: the noisy target wins, as `forcings.assign(noisy_targets)` does in
        # the reference (gencast/denoiser.py:184)
        forc = forc.drop_vars([n for n in forc.keys() if n in noisy_targets])
        sizes = dict(inputs.sizes)
        sizes.setdefault("batch", 1)
        layout = ChannelLayout(
            num_input_channels=sum(c for _, c in stacking.channel_layout(inputs)),
            forcing_vars=tuple(stacking.channel_layout(forc)),
            target_vars=tuple(stacking.channel_layout(noisy_targets)))
        st = self._arch.sparse_transformer_config
        graphs = build_denoiser_graphs(lat, lon, self._arch.mesh_size, st.attention_k_hop,
                                       self._arch.radius_query_fraction_edge_length)
        if self._params is None:
            shapes = param_shapes(self._arch, layout.num_data_channels, layout.num_targets, self._noise_cfg)
            seed = 0
            if self._rngs is not None:
                seed = self._rngs.params()
            init = init_reference_like if self._param_init == "reference" else init_perturbed
            self._params = init(shapes, seed=seed)
        # all batch elements of a call are evaluated together when they share the noise level (the
        # sampler's case): the engine is built for that many members
        self._engine = DenoiserEngine(graphs, self._arch, self._params, layout, self._noise_cfg,
                                      compute_dtype=self._compute_dtype, device=self._device,
                                      members=int(sizes.get("batch", 1)))
        self._grid_key = key
        return self._engine

    @property
    def engine(self) -> DenoiserEngine:
        if self._engine is None:
            raise RuntimeError("Denoiser is initialised lazily on its first call")
        return self._engine

    @property
    def params(self) -> Dict[str, np.ndarray]:
        return self._params

    @property
    def stacker(self) -> DeviceStacker:
        if self._stacker is None:
            self._stacker = DeviceStacker(self.engine.device)
        return self._stacker

    def stack_constants(self, inputs: Dataset, forcings: Optional[Dataset], sizes):
        """(inputs, forcings) as device tensors [G, B, C] (layout transposes run on the GPU)."""
        forc = forcings if forcings is not None else Dataset({}, inputs.coords)
        names = {n for n, _ in self.engine.layout.forcing_vars}
        forc = forc.drop_vars([n for n in forc.keys() if n not in names])      # forcings overridden by noisy targets
        return self.stacker.to_nodes("inputs", inputs, sizes), self.stacker.to_nodes("forcings", forc, sizes)

    def __call__(self, inputs: Dataset, noisy_targets: Dataset, noise_levels: DataArray,
                 forcings: Optional[Dataset] = None, **kwargs) -> Dataset:
        if tuple(noise_levels.dims) != ("batch",):
            raise ValueError("noise_levels expected to be shape (batch,).")     # denoiser.py:188-189
        engine = self._maybe_init(inputs, noisy_targets, forcings)
        sizes = dict(noisy_targets.sizes)
        sizes.setdefault("batch", 1)
        batch = sizes["batch"]
        if noise_levels.shape[0] != batch:
            raise ValueError("noise_levels must have one entry per batch element")
        levels = np.asarray(noise_levels.data, np.float64)
        if engine.B != batch:
            raise ValueError(f"Denoiser was initialised for batch size {engine.B}, got {batch}")
        with torch.cuda.device(engine.device):
            inp, frc = self.stack_constants(inputs, forcings, sizes)
            noisy = self.stacker.to_nodes("noisy", noisy_targets, sizes)
            engine.set_constant_features(self.member_major(inp), self.member_major(frc))
            engine.set_network_input(self.member_major(noisy))
            uniq = np.unique(levels)
            if len(uniq) == 1:
                f = engine.forward(engine.sigma_context(float(levels[0])))[:, :engine.n_out]
            else:
                # batch elements at different noise levels (the reference conditions every element on its own level,
                # gencast/denoiser.py:190-198): the members are evaluated together once per distinct level and each
                # keeps the rows computed with its level
                f = torch.empty(batch * engine.G, engine.n_out, dtype=torch.float32, device=engine.device)
                for lv in uniq:
                    g = engine.forward(engine.sigma_context(float(lv)))
                    for b in np.nonzero(levels == lv)[0]:
                        f[b * engine.G:(b + 1) * engine.G] = g[b * engine.G:(b + 1) * engine.G, :engine.n_out]
            out = f.reshape(batch, engine.G, engine.n_out).permute(1, 0, 2)
            return self.stacker.from_nodes(out, noisy_targets)

    @staticmethod
    def member_major(nodes: torch.Tensor) -> torch.Tensor:
        """[G, B, C] (the reference's node-major layout, gencast/denoiser.py:833-837) -> [B * G, C]."""
        return nodes.permute(1, 0, 2).reshape(nodes.shape[0] * nodes.shape[1], nodes.shape[2]).contiguous()
