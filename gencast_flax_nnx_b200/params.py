"""Parameter tree of the GenCast denoiser, keyed by the reference's NNX paths.

The flat keys are the attribute paths of the reference's Flax NNX module tree
joined with '/', so a trained checkpoint of the reference maps one to one
(reference: common/mlp.py:40-265,
common/deep_typed_graph_net.py:355-490, gencast/sparse_transformer.py:252-634,
gencast/denoiser.py:365-414).  Linear kernels are [in, out], biases [out].
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .configs import DenoiserArchitectureConfig, NoiseEncoderConfig

COND_DIM = 16  # norm_conditioning_dim (reference: common/deep_typed_graph_net.py:159,207)

_G2M = "denoiser/predictor/grid2mesh_gnn"
_M2G = "denoiser/predictor/mesh2grid_gnn"
_TFM = "denoiser/predictor/mesh_gnn/batch_first_transformer"
_ENC = "denoiser/noise_level_encoder"


def _mlp_entries(prefix: str, n_in: int, hidden: int, n_out: int, cond: bool) -> List[Tuple[str, Tuple[int, ...]]]:
    """Entries of one MLPWithNormConditioning (reference: common/mlp.py:68-203)."""
    out = [
        (f"{prefix}/network/network/layers/0/kernel", (n_in, hidden)),
        (f"{prefix}/network/network/layers/0/bias", (hidden,)),
        (f"{prefix}/network/network/layers/2/kernel", (hidden, n_out)),
        (f"{prefix}/network/network/layers/2/bias", (n_out,)),
    ]
    if cond:
        out += [
            (f"{prefix}/norm_conditioning_layer/conditional_linear_layer/kernel", (COND_DIM, 2 * n_out)),
            (f"{prefix}/norm_conditioning_layer/conditional_linear_layer/bias", (2 * n_out,)),
        ]
    return out


def param_shapes(arch: DenoiserArchitectureConfig, data_channels: int, n_out: int,
                 noise_cfg: NoiseEncoderConfig = NoiseEncoderConfig(),
                 node_struct: int = 3, edge_struct: int = 4) -> Dict[str, Tuple[int, ...]]:
    """All parameters with their shapes, in the reference's creation order."""
    L = arch.latent_size
    st = arch.sparse_transformer_config
    D, F = st.d_model, st.ffw_hidden
    assert arch.hidden_layers == 1, "GenCast uses one hidden layer per MLP (reference: train_helpers.py:137)"
    entries: List[Tuple[str, Tuple[int, ...]]] = []
    # sigma encoder (reference: common/mlp.py:230-253)
    n_in = 2 * noise_cfg.num_frequencies
    for i, n_o in enumerate(noise_cfg.output_sizes):
        entries += [(f"{_ENC}/linear_{i}/kernel", (n_in, n_o)), (f"{_ENC}/linear_{i}/bias", (n_o,))]
        n_in = n_o
    assert n_in == COND_DIM
    node_in = node_struct + data_channels
    # grid2mesh encoder (reference: denoiser.py:365-386)
    entries += _mlp_entries(f"{_G2M}/embedder_network/embed_edge_fns/grid2mesh", edge_struct, L, L, True)
    entries += _mlp_entries(f"{_G2M}/embedder_network/embed_node_fns/grid_nodes", node_in, L, L, True)
    entries += _mlp_entries(f"{_G2M}/embedder_network/embed_node_fns/mesh_nodes", node_in, L, L, True)
    gn = f"{_G2M}/processor_networks/0/graph_network"
    entries += _mlp_entries(f"{gn}/update_edge_fns/grid2mesh/edge_fn", 3 * L, L, L, True)
    entries += _mlp_entries(f"{gn}/update_node_fns/grid_nodes/node_fn", L, L, L, True)
    entries += _mlp_entries(f"{gn}/update_node_fns/mesh_nodes/node_fn", 2 * L, L, L, True)
    # mesh transformer (reference: sparse_transformer.py:296-305,461-483,608-622)
    for i in range(st.num_layers):
        b = f"{_TFM}/blocks/{i}"
        for p in ("q_proj", "k_proj", "v_proj"):
            entries.append((f"{b}/attn_module/{p}/linear/kernel", (D, D)))
        entries += [(f"{b}/attn_module/final_linear/kernel", (D, D)),
                    (f"{b}/attn_module/final_linear/bias", (D,)),
                    (f"{b}/ffw_module/mlp/layers/0/kernel", (D, F)),
                    (f"{b}/ffw_module/mlp/layers/0/bias", (F,)),
                    (f"{b}/ffw_module/mlp/layers/2/kernel", (F, D)),
                    (f"{b}/ffw_module/mlp/layers/2/bias", (D,)),
                    (f"{b}/norm_cond_attn/conditional_linear_layer/kernel", (COND_DIM, 2 * D)),
                    (f"{b}/norm_cond_attn/conditional_linear_layer/bias", (2 * D,)),
                    (f"{b}/norm_cond_ffw/conditional_linear_layer/kernel", (COND_DIM, 2 * D)),
                    (f"{b}/norm_cond_ffw/conditional_linear_layer/bias", (2 * D,))]
    entries += [(f"{_TFM}/final_norm_cond/conditional_linear_layer/kernel", (COND_DIM, 2 * D)),
                (f"{_TFM}/final_norm_cond/conditional_linear_layer/bias", (2 * D,))]
    # mesh2grid decoder (reference: denoiser.py:395-414)
    entries += _mlp_entries(f"{_M2G}/embedder_network/embed_edge_fns/mesh2grid", edge_struct, L, L, True)
    gn = f"{_M2G}/processor_networks/0/graph_network"
    entries += _mlp_entries(f"{gn}/update_edge_fns/mesh2grid/edge_fn", 3 * L, L, L, True)
    entries += _mlp_entries(f"{gn}/update_node_fns/grid_nodes/node_fn", 2 * L, L, L, True)
    # Present in the reference's tree, evaluated, result unused.
    entries += _mlp_entries(f"{gn}/update_node_fns/mesh_nodes/node_fn", L, L, L, True)
    entries += _mlp_entries(f"{_M2G}/decoder_network/embed_node_fns/grid_nodes", L, L, n_out, False)
    return dict(entries)


def init_perturbed(shapes: Dict[str, Tuple[int, ...]], seed: int = 1) -> Dict[str, np.ndarray]:
    """Random O(1/sqrt(fan_in)) weights for parity fixtures and benchmarks.

    NOT the reference initialisation: at the reference's init every transformer
    block is the identity and every conditional norm is a no-op (SURVEY.md fact
    4: gencast/denoiser.py:93-95, common/mlp.py:43-46), which would make parity
    vacuous.  Kernels ~ N(0, 1/fan_in), biases ~ N(0, 0.01^2), conditional
    linears ~ N(0, 0.1^2/16).
    """
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in shapes.items():
        if name.endswith("/kernel"):
            if "conditional_linear_layer" in name:
                std = 0.1 / np.sqrt(shape[0])
            else:
                std = 1.0 / np.sqrt(shape[0])
            out[name] = (rng.standard_normal(shape) * std).astype(np.float32)
        else:
            std = 0.1 if "conditional_linear_layer" in name else 0.01
            out[name] = (rng.standard_normal(shape) * std).astype(np.float32)
    return out


def mlp_prefixes():
    """Handy names of the MLPWithNormConditioning blocks."""
    g2m = f"{_G2M}/processor_networks/0/graph_network"
    m2g = f"{_M2G}/processor_networks/0/graph_network"
    return dict(
        g2m_edge_embed=f"{_G2M}/embedder_network/embed_edge_fns/grid2mesh",
        g2m_grid_embed=f"{_G2M}/embedder_network/embed_node_fns/grid_nodes",
        g2m_mesh_embed=f"{_G2M}/embedder_network/embed_node_fns/mesh_nodes",
        g2m_edge_update=f"{g2m}/update_edge_fns/grid2mesh/edge_fn",
        g2m_grid_update=f"{g2m}/update_node_fns/grid_nodes/node_fn",
        g2m_mesh_update=f"{g2m}/update_node_fns/mesh_nodes/node_fn",
        m2g_edge_embed=f"{_M2G}/embedder_network/embed_edge_fns/mesh2grid",
        m2g_edge_update=f"{m2g}/update_edge_fns/mesh2grid/edge_fn",
        m2g_grid_update=f"{m2g}/update_node_fns/grid_nodes/node_fn",
        m2g_mesh_update=f"{m2g}/update_node_fns/mesh_nodes/node_fn",
        m2g_output=f"{_M2G}/decoder_network/embed_node_fns/grid_nodes",
        transformer=_TFM,
        noise_encoder=_ENC,
    )


def init_reference_like(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, num_layers: int = 16,
                        attn_final_mult: float = 0.0, ffw_final_mult: float = 0.0) -> Dict[str, np.ndarray]:
    """Same initial *distributions* as the reference (not the same draws: NNX's RNG is JAX threefry).

    MLP kernels xavier-uniform, biases 0 (common/mlp.py:167-199); conditional linears
    truncated-normal(1e-8) (common/mlp.py:43-46); q/k/v and the first FFW layer
    variance_scaling(2/num_layers, fan_in, truncated normal) (gencast/sparse_transformer.py:256,275);
    attention output projection and second FFW layer variance_scaling(mult/num_layers), mult = 0 by
    default, i.e. zeros (gencast/denoiser.py:93-95, sparse_transformer.py:257,302); noise-level encoder
    variance_scaling(2, fan_in, uniform) (common/mlp.py:225-228).
    """
    rng = np.random.default_rng(seed)

    def trunc_normal(shape, std):
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2
        return x * std / 0.87962566103423978   # jax's truncated-normal variance correction

    out = {}
    for name, shape in shapes.items():
        if name.endswith("/bias"):
            out[name] = np.zeros(shape, np.float32)
            continue
        fan_in, fan_out = shape
        if "conditional_linear_layer" in name:
            w = trunc_normal(shape, 1e-8)
        elif "/noise_level_encoder/" in name:
            lim = np.sqrt(3.0 * 2.0 / fan_in)
            w = rng.uniform(-lim, lim, shape)
        elif "/attn_module/final_linear/" in name:
            w = trunc_normal(shape, np.sqrt(attn_final_mult / num_layers / fan_in)) if attn_final_mult > 0 else np.zeros(shape)
        elif "/ffw_module/mlp/layers/2/" in name:
            w = trunc_normal(shape, np.sqrt(ffw_final_mult / num_layers / fan_in)) if ffw_final_mult > 0 else np.zeros(shape)
        elif "/attn_module/" in name or "/ffw_module/" in name:
            w = trunc_normal(shape, np.sqrt(2.0 / num_layers / fan_in))
        else:
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            w = rng.uniform(-lim, lim, shape)
        out[name] = w.astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# Weight interchange: the flat dict <-> files / NNX state
# ----------------------------------------------------------------------------------------------

def save_npz(params: Dict[str, np.ndarray], path: str) -> None:
    """One array per NNX attribute path ('/' separated), float32."""
    np.savez(path, **{k.replace("/", "|"): np.asarray(v, np.float32) for k, v in params.items()})


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as f:
        return {k.replace("|", "/"): f[k] for k in f.files}


def from_nnx_state(flat_state) -> Dict[str, np.ndarray]:
    """Flattens `nnx.state(model, nnx.Param).flat_state()` (an iterable of (path tuple, variable)) into the
    flat dict used here.  The Orbax checkpoints of the reference additionally wrap the update functions
    in `graph_network` / `edge_fn` / `node_fn` levels, which are already part of these paths
    (training/evaluation.py:143-148); integer path components (list indices) are kept as digits."""
    out = {}
    for path, var in flat_state:
        value = getattr(var, "value", var)
        out["/".join(str(p) for p in path)] = np.asarray(value, np.float32)
    return out


def check_complete(params: Dict[str, np.ndarray], shapes: Dict[str, Tuple[int, ...]]) -> None:
    """Raises with the list of missing / mis-shaped entries."""
    missing = [k for k in shapes if k not in params]
    bad = [f"{k}: {tuple(params[k].shape)} != {shapes[k]}" for k in shapes if k in params and tuple(params[k].shape) != tuple(shapes[k])]
    if missing or bad:
        raise ValueError(f"parameter tree mismatch; missing={missing[:5]}{'...' if len(missing) > 5 else ''} bad={bad[:5]}")
