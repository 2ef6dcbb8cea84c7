"""Configuration dataclasses, same names and fields as the reference.

Reference: gencast/gencast.py:57-115 (TASK, SamplerConfig, NoiseConfig),
gencast/denoiser.py:47-139 (NoiseEncoderConfig, SparseTransformerConfig,
DenoiserArchitectureConfig), graphcast/graphcast.py:61-143 (variable lists,
pressure levels, TaskConfig).  The reference uses chex dataclasses; plain
dataclasses keep the same constructor keywords and attribute names.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

PRESSURE_LEVELS_WEATHERBENCH_13 = (50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000)

ALL_ATMOSPHERIC_VARS = (
    "potential_vorticity", "specific_rain_water_content", "specific_snow_water_content",
    "geopotential", "temperature", "u_component_of_wind", "v_component_of_wind",
    "specific_humidity", "vertical_velocity", "vorticity", "divergence", "relative_humidity",
    "ozone_mass_mixing_ratio", "specific_cloud_liquid_water_content",
    "specific_cloud_ice_water_content", "fraction_of_cloud_cover",
)
TARGET_SURFACE_NO_PRECIP_VARS = (
    "2m_temperature", "mean_sea_level_pressure", "10m_v_component_of_wind", "10m_u_component_of_wind",
)
TARGET_ATMOSPHERIC_VARS = (
    "temperature", "geopotential", "u_component_of_wind", "v_component_of_wind",
    "vertical_velocity", "specific_humidity",
)
GENERATED_FORCING_VARS = ("year_progress_sin", "year_progress_cos", "day_progress_sin", "day_progress_cos")
STATIC_VARS = ("geopotential_at_surface", "land_sea_mask")


@dataclasses.dataclass(frozen=True)
class TaskConfig:
    """Reference: graphcast/graphcast.py:135-143."""
    input_variables: Tuple[str, ...]
    target_variables: Tuple[str, ...]
    forcing_variables: Tuple[str, ...]
    pressure_levels: Tuple[int, ...]
    input_duration: str


# Reference: gencast/gencast.py:57-71.
TASK = TaskConfig(
    input_variables=(TARGET_SURFACE_NO_PRECIP_VARS + TARGET_ATMOSPHERIC_VARS
                     + GENERATED_FORCING_VARS + STATIC_VARS),
    target_variables=TARGET_SURFACE_NO_PRECIP_VARS + TARGET_ATMOSPHERIC_VARS,
    forcing_variables=GENERATED_FORCING_VARS,
    pressure_levels=PRESSURE_LEVELS_WEATHERBENCH_13,
    input_duration="24h",
)


@dataclasses.dataclass(frozen=True)
class SamplerConfig:
    """Reference: gencast/gencast.py:74-108 (same defaults)."""
    max_noise_level: float = 80.0
    min_noise_level: float = 0.03
    num_noise_levels: int = 20
    rho: float = 7.0
    stochastic_churn_rate: float = 2.5
    churn_min_noise_level: float = 0.75
    churn_max_noise_level: float = float("inf")
    noise_level_inflation_factor: float = 1.05


@dataclasses.dataclass(frozen=True)
class NoiseConfig:
    """Reference: gencast/gencast.py:111-115."""
    training_noise_level_rho: float = 7.0
    training_max_noise_level: float = 88.0
    training_min_noise_level: float = 0.02


@dataclasses.dataclass(frozen=True)
class NoiseEncoderConfig:
    """Reference: gencast/denoiser.py:47-68."""
    apply_log_first: bool = True
    base_period: float = 16.0
    num_frequencies: int = 32
    output_sizes: Tuple[int, int] = (32, 16)


@dataclasses.dataclass
class SparseTransformerConfig:
    """Reference: gencast/denoiser.py:71-97.

    Only attention_type == 'triblockdiag_mha' semantics are implemented (the
    reference default; 'splash_mha' is a TPU kernel).  The block_* and mask_type
    fields are accepted and ignored: our attention kernel works on the exact
    k-hop pattern, which is what the tri-block mask evaluates to.
    """
    attention_k_hop: int
    d_model: int
    num_layers: int = 16
    num_heads: int = 4
    attention_type: str = "triblockdiag_mha"
    mask_type: str = "lazy"
    block_q: int = 1024
    block_kv: int = 512
    block_kv_compute: int = 256
    block_q_dkv: int = 512
    block_kv_dkv: int = 1024
    block_kv_dkv_compute: int = 1024
    ffw_winit_final_mult: float = 0.0
    attn_winit_final_mult: float = 0.0
    ffw_hidden: int = 2048


@dataclasses.dataclass
class DenoiserArchitectureConfig:
    """Reference: gencast/denoiser.py:100-139."""
    sparse_transformer_config: SparseTransformerConfig
    mesh_size: int
    latent_size: int = 512
    hidden_layers: int = 1
    radius_query_fraction_edge_length: float = 0.6
    norm_conditioning_features: Tuple[str, ...] = ("noise_level_encodings",)
    grid2mesh_aggregate_normalization: Optional[float] = None
    node_output_size: Optional[int] = None


def num_outputs(task: TaskConfig) -> int:
    """Output channels per grid node (reference: gencast/gencast.py:158-169)."""
    n_surface = len(set(task.target_variables) - set(ALL_ATMOSPHERIC_VARS))
    n_atmos = len(set(task.target_variables) & set(ALL_ATMOSPHERIC_VARS))
    return n_surface + len(task.pressure_levels) * n_atmos


# Named model sizes of BASELINE.json / SURVEY.md §8.
def named_config(name: str):
    """Returns (resolution_deg, DenoiserArchitectureConfig) for 'tiny' | 'nano' | '1deg' | '0p25deg'.

    'nano' and '1deg' are BASELINE.json configs (nano = 2.5 deg,
    mesh 4, L=256; 1deg = mesh 5, L=512; both 16 layers, 4 heads, ffw 2048, k=8).
    'tiny' is a test-only size the CPU oracle finishes in about a second.
    """
    if name == "tiny":
        st = SparseTransformerConfig(attention_k_hop=2, d_model=128, num_layers=2, num_heads=4, ffw_hidden=256)
        return 10.0, DenoiserArchitectureConfig(sparse_transformer_config=st, mesh_size=2, latent_size=128)
    if name == "nano":
        st = SparseTransformerConfig(attention_k_hop=8, d_model=256, num_layers=16, num_heads=4, ffw_hidden=2048)
        return 2.5, DenoiserArchitectureConfig(sparse_transformer_config=st, mesh_size=4, latent_size=256)
    if name == "1deg":
        st = SparseTransformerConfig(attention_k_hop=8, d_model=512, num_layers=16, num_heads=4, ffw_hidden=2048)
        return 1.0, DenoiserArchitectureConfig(sparse_transformer_config=st, mesh_size=5, latent_size=512)
    if name == "0p25deg":
        st = SparseTransformerConfig(attention_k_hop=8, d_model=512, num_layers=16, num_heads=4, ffw_hidden=2048)
        return 0.25, DenoiserArchitectureConfig(sparse_transformer_config=st, mesh_size=6, latent_size=512)
    raise ValueError(f"unknown config {name!r}")
