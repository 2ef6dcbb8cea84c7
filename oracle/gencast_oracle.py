"""CPU oracle for the GenCast denoiser + DPM-Solver++ 2S hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under gencast_flax_nnx_b200/ imports this
module; it is the checker for tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.

It restates, op by op and *as written* (dense tri-block-diagonal attention,
[e|s|r] concatenation before the edge MLP, scatter-add segment sum), the
arithmetic of the reference fgiral000/gencast-flax-nnx on torch CPU tensors in
float64 (ground truth) or float32 (the timed "reference CPU path").  Every
function cites the reference file:line it follows.

PARITY PIN STATUS.  The reference ships no test, golden vector or fixture for
any function on this path, and JAX / Flax / jraph are not
installable here, so the reference cannot be executed natively.  Pins in place:
  * tests/golden/refshim_*.npz — outputs of the reference's OWN module code
    (common/mlp.py, common/typed_graph_net.py, common/deep_typed_graph_net.py,
    gencast/sparse_transformer.py, gencast/transformer.py, gencast/denoiser.py
    graph wiring) executed in this container under a numpy stand-in for the
    jax / flax.nnx / jraph API surface (tools/refshim/, generator
    tools/make_refshim_golden.py).  That pins wiring, concatenation order, masks,
    residuals and padding against the reference source itself.
  * the third-party primitives the stand-in has to supply (nnx.Linear,
    nnx.LayerNorm eps=1e-6 fast variance, jax.nn.swish / gelu(tanh) / softmax,
    jraph.segment_sum) are restated from their published definitions:
    **parity unpinned** for those.
  * static graph tables are pinned bit-exactly against the importable parts of
    the reference (tools/make_graph_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

Params = Mapping[str, np.ndarray]

_G2M = "denoiser/predictor/grid2mesh_gnn"
_M2G = "denoiser/predictor/mesh2grid_gnn"
_TFM = "denoiser/predictor/mesh_gnn/batch_first_transformer"
_ENC = "denoiser/noise_level_encoder"


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a)).to(dtype)


# ----------------------------------------------------------------------------
# Primitives
# ----------------------------------------------------------------------------

def swish(x):
    """jax.nn.swish = x * sigmoid(x) (reference use: common/deep_typed_graph_net.py:61-62)."""
    return x * torch.sigmoid(x)


def gelu_tanh(x):
    """jax.nn.gelu default approximate=True (reference use: common/mlp.py:215, sparse_transformer.py:264)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def linear(p: Params, prefix: str, x, dtype, bias: bool = True):
    """flax.nnx.Linear: x @ kernel[in,out] + bias."""
    y = x @ _t(p[f"{prefix}/kernel"], dtype)
    if bias:
        y = y + _t(p[f"{prefix}/bias"], dtype)
    return y


def layer_norm(x, eps: float = 1e-6):
    """flax.nnx.LayerNorm without scale/bias, use_fast_variance=True.

    var = E[x^2] - E[x]^2 clipped at 0 (reference construction:
    common/mlp.py:95-103, sparse_transformer.py:482-483,620).
    """
    mean = x.mean(-1, keepdim=True)
    var = ((x * x).mean(-1, keepdim=True) - mean * mean).clamp_min(0.0)
    return (x - mean) * torch.rsqrt(var + eps)


def linear_norm_conditioning(p: Params, prefix: str, x, cond, dtype):
    """x * (1 + s) + o with [s | o] = Linear(cond) (reference: common/mlp.py:59-65)."""
    so = linear(p, f"{prefix}/conditional_linear_layer", cond, dtype)
    s, o = torch.chunk(so, 2, dim=-1)
    return x * (s + 1.0) + o


def mlp_with_norm_conditioning(p: Params, prefix: str, x, cond, dtype,
                               use_layer_norm: bool = True, use_cond: bool = True):
    """Linear -> swish -> Linear -> LayerNorm -> conditional affine.

    Reference: common/mlp.py:115-147 (MLP body :152-203).  x is [n, B, C];
    cond is [B, 16] and is broadcast over the leading node axis (:127-131).
    """
    h = swish(linear(p, f"{prefix}/network/network/layers/0", x, dtype))
    y = linear(p, f"{prefix}/network/network/layers/2", h, dtype)
    if use_layer_norm:
        y = layer_norm(y)
    if use_cond:
        y = linear_norm_conditioning(p, f"{prefix}/norm_conditioning_layer", y, cond[None, :, :], dtype)
    return y


def segment_sum(data, segment_ids: np.ndarray, num_segments: int):
    """jraph.segment_sum / jax.ops.segment_sum: scatter-add, zeros for empty segments.

    Reference call sites: common/typed_graph_net.py:173,182.
    """
    out = torch.zeros((num_segments,) + tuple(data.shape[1:]), dtype=data.dtype)
    out.index_add_(0, torch.as_tensor(segment_ids, dtype=torch.long), data)
    return out


# ----------------------------------------------------------------------------
# Noise level encoder
# ----------------------------------------------------------------------------

def fourier_features(values, base_period: float, num_frequencies: int):
    """[cos(2 pi k v / P), sin(...)] for k = 1..K (reference: common/model_utils.py:728-757)."""
    freqs = np.arange(1, num_frequencies + 1) / base_period
    ang = torch.as_tensor(2 * np.pi * freqs).to(values.dtype)
    v = values[..., None] * ang
    return torch.cat([torch.cos(v), torch.sin(v)], dim=-1)


def noise_level_encoder(p: Params, sigma, dtype, base_period=16.0, num_frequencies=32):
    """log sigma -> Fourier features -> Linear -> gelu -> Linear  (reference: common/mlp.py:255-265)."""
    z = torch.log(sigma.to(dtype))
    f = fourier_features(z, base_period, num_frequencies)
    h = gelu_tanh(linear(p, f"{_ENC}/linear_0", f, dtype))
    return linear(p, f"{_ENC}/linear_1", h, dtype)


# ----------------------------------------------------------------------------
# Encoder: grid2mesh GNN
# ----------------------------------------------------------------------------

def _batch_second(x2d, batch: int):
    """[n, C] -> [n, B, C] (reference: gencast/denoiser.py:833-837)."""
    return x2d[:, None, :].expand(-1, batch, -1)


def grid2mesh_gnn(p: Params, g: Mapping[str, np.ndarray], grid_node_features, cond, dtype):
    """One embed + one interaction step on the grid->mesh bipartite graph.

    Reference: gencast/denoiser.py:602-688 (feature assembly: structural
    features then data channels; mesh nodes get zeros for the data block),
    common/deep_typed_graph_net.py:493-581, common/typed_graph_net.py:88-195,
    :295-326.  Aggregation is float32 in the reference (denoiser.py:371); in this
    oracle it runs in the oracle dtype (>= f32).
    Returns (mesh [V,B,L], grid [G,B,L]) with residuals applied.
    """
    B = grid_node_features.shape[1]
    V = g["g2m_mesh_feat"].shape[0]
    c_data = grid_node_features.shape[-1]
    grid_in = torch.cat([_batch_second(_t(g["g2m_grid_feat"], dtype), B), grid_node_features], dim=-1)
    mesh_in = torch.cat([_batch_second(_t(g["g2m_mesh_feat"], dtype), B),
                         torch.zeros((V, B, c_data), dtype=dtype)], dim=-1)
    edge_in = _batch_second(_t(g["g2m_edge_feat"], dtype), B)
    emb = f"{_G2M}/embedder_network"
    e0 = mlp_with_norm_conditioning(p, f"{emb}/embed_edge_fns/grid2mesh", edge_in, cond, dtype)
    g0 = mlp_with_norm_conditioning(p, f"{emb}/embed_node_fns/grid_nodes", grid_in, cond, dtype)
    m0 = mlp_with_norm_conditioning(p, f"{emb}/embed_node_fns/mesh_nodes", mesh_in, cond, dtype)
    gn = f"{_G2M}/processor_networks/0/graph_network"
    s = torch.as_tensor(g["g2m_senders"], dtype=torch.long)
    r = torch.as_tensor(g["g2m_receivers"], dtype=torch.long)
    # Edge update on [e | sender | receiver] (typed_graph_net.py:134-159, :301-305).
    e1 = mlp_with_norm_conditioning(p, f"{gn}/update_edge_fns/grid2mesh/edge_fn",
                                    torch.cat([e0, g0[s], m0[r]], dim=-1), cond, dtype)
    # Node updates (typed_graph_net.py:161-195, :315-326): grid nodes receive nothing.
    agg = segment_sum(e1, g["g2m_receivers"], V)
    g1 = mlp_with_norm_conditioning(p, f"{gn}/update_node_fns/grid_nodes/node_fn", g0, cond, dtype)
    m1 = mlp_with_norm_conditioning(p, f"{gn}/update_node_fns/mesh_nodes/node_fn",
                                    torch.cat([m0, agg], dim=-1), cond, dtype)
    # Residuals (deep_typed_graph_net.py:569-581).
    return m0 + m1, g0 + g1


# ----------------------------------------------------------------------------
# Processor: mesh transformer with tri-block-diagonal attention, as written
# ----------------------------------------------------------------------------

def mask_block_size(mask_csr) -> int:
    """Reference: gencast/sparse_transformer.py:86-96."""
    m = (mask_csr != 0).tocsc()
    n = m.shape[0]
    first_row = np.full(n, n); last_row = np.full(n, -1)
    coo = m.tocoo()
    np.minimum.at(first_row, coo.col, coo.row)
    np.maximum.at(last_row, coo.col, coo.row)
    cols = np.arange(n)
    return int(max((cols - first_row + 1).max(), (last_row - cols + 1).max()))


def mask_block_diags(mask_csr, num_padding: int, bs: int) -> torch.Tensor:
    """[3, nb, bs, bs] boolean diag / upper / lower blocks (reference: sparse_transformer.py:163-201)."""
    from scipy import sparse
    n = mask_csr.shape[0] + num_padding
    coo = (mask_csr != 0).tocoo()
    m = sparse.csr_matrix((np.ones(coo.nnz, dtype=np.int8), (coo.row, coo.col)), shape=(n, n))
    nb = n // bs
    zero = np.zeros((bs, bs), dtype=bool)
    diag = [m[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs].toarray().astype(bool) for i in range(nb)]
    upper = [m[i * bs:(i + 1) * bs, (i + 1) * bs:(i + 2) * bs].toarray().astype(bool) for i in range(nb - 1)] + [zero]
    lower = [zero] + [m[(i + 1) * bs:(i + 2) * bs, i * bs:(i + 1) * bs].toarray().astype(bool) for i in range(nb - 1)]
    return torch.from_numpy(np.stack([np.stack(diag), np.stack(upper), np.stack(lower)]))


def triblockdiag_mha(p: Params, prefix: str, x, mask, num_heads: int, dtype):
    """Reference: gencast/sparse_transformer.py:309-354 and :100-125 (joint softmax).

    x: [B, nb, bs, D]; mask: [3, nb, bs, bs] bool.
    """
    B, nb, bs, D = x.shape
    d = D // num_heads
    def proj(name):
        return (x @ _t(p[f"{prefix}/{name}/linear/kernel"], dtype)).reshape(B, nb, bs, num_heads, d)
    q, k, v = proj("q_proj"), proj("k_proj"), proj("v_proj")
    zk = torch.zeros_like(k[:, :1])
    k = torch.cat([zk, k, zk], dim=1)
    v = torch.cat([zk, v, zk], dim=1)
    scale = d ** -0.5
    def qk(keys):
        return torch.einsum("bnqhd,bnkhd->bnhqk", q, keys) * scale
    logits = [qk(k[:, 1:-1]), qk(k[:, 2:]), qk(k[:, :-2])]
    neg = torch.tensor(-1e30, dtype=dtype)
    logits = [torch.where(mask[i][None, :, None], l, neg) for i, l in enumerate(logits)]
    m = torch.stack([l.max(-1, keepdim=True).values for l in logits]).max(0).values
    un = [torch.exp(l - m) for l in logits]
    denom = sum(u.sum(-1, keepdim=True) for u in un)
    w = [u / denom for u in un]
    def av(wts, vals):
        return torch.einsum("bnhqk,bnkhd->bnqhd", wts, vals)
    out = av(w[0], v[:, 1:-1]) + av(w[1], v[:, 2:]) + av(w[2], v[:, :-2])
    out = out.reshape(B, nb, bs, D)
    return linear(p, f"{prefix}/final_linear", out, dtype)


def mesh_transformer(p: Params, khop_csr, x, cond, num_layers: int, num_heads: int, dtype):
    """Pre-LN blocks with conditional norms, then final conditional norm.

    Reference: gencast/transformer.py:94-121 ([V,B,D] <-> [B,V,D]),
    gencast/sparse_transformer.py:486-525 (Block), :554-567 (padding and mask),
    :624-634 (final norm).  x: [V, B, D] -> [V, B, D].
    """
    V = x.shape[0]
    bs = mask_block_size(khop_csr)
    pad = int(np.ceil(V / bs) * bs - V)
    mask = mask_block_diags(khop_csr, pad, bs)
    y = x.transpose(0, 1)                                   # [B, V, D]
    cond_b = cond[:, None, :]
    for i in range(num_layers):
        b = f"{_TFM}/blocks/{i}"
        h = linear_norm_conditioning(p, f"{b}/norm_cond_attn", layer_norm(y), cond_b, dtype)
        hp = torch.nn.functional.pad(h, (0, 0, 0, pad))
        hp = hp.reshape(hp.shape[0], hp.shape[1] // bs, bs, hp.shape[-1])
        a = triblockdiag_mha(p, f"{b}/attn_module", hp, mask, num_heads, dtype)
        a = a.reshape(a.shape[0], V + pad, a.shape[-1])[:, :V]
        y = y + a
        h = linear_norm_conditioning(p, f"{b}/norm_cond_ffw", layer_norm(y), cond_b, dtype)
        f = gelu_tanh(linear(p, f"{b}/ffw_module/mlp/layers/0", h, dtype))
        y = y + linear(p, f"{b}/ffw_module/mlp/layers/2", f, dtype)
    y = linear_norm_conditioning(p, f"{_TFM}/final_norm_cond", layer_norm(y), cond_b, dtype)
    return y.transpose(0, 1)


# ----------------------------------------------------------------------------
# Decoder: mesh2grid GNN
# ----------------------------------------------------------------------------

def mesh2grid_gnn(p: Params, g: Mapping[str, np.ndarray], mesh_nodes, grid_nodes, cond, dtype):
    """Edge embed, one interaction step, output MLP on grid nodes.

    Reference: gencast/denoiser.py:730-768 with the decoder configuration
    :395-414 (embed_nodes=False; output MLP without LayerNorm/conditioning,
    deep_typed_graph_net.py:469-485).  The mesh-node update of the reference is
    evaluated there but its result is never read; it is
    omitted here because it cannot influence the output.
    """
    B = grid_nodes.shape[1]
    G = grid_nodes.shape[0]
    e0 = mlp_with_norm_conditioning(p, f"{_M2G}/embedder_network/embed_edge_fns/mesh2grid",
                                    _batch_second(_t(g["m2g_edge_feat"], dtype), B), cond, dtype)
    gn = f"{_M2G}/processor_networks/0/graph_network"
    s = torch.as_tensor(g["m2g_senders"], dtype=torch.long)
    r = torch.as_tensor(g["m2g_receivers"], dtype=torch.long)
    e1 = mlp_with_norm_conditioning(p, f"{gn}/update_edge_fns/mesh2grid/edge_fn",
                                    torch.cat([e0, mesh_nodes[s], grid_nodes[r]], dim=-1), cond, dtype)
    agg = segment_sum(e1, g["m2g_receivers"], G)
    g1 = mlp_with_norm_conditioning(p, f"{gn}/update_node_fns/grid_nodes/node_fn",
                                    torch.cat([grid_nodes, agg], dim=-1), cond, dtype)
    g2 = grid_nodes + g1
    return mlp_with_norm_conditioning(p, f"{_M2G}/decoder_network/embed_node_fns/grid_nodes", g2, cond, dtype,
                                      use_layer_norm=False, use_cond=False)


# ----------------------------------------------------------------------------
# Denoiser, preconditioning, sampler
# ----------------------------------------------------------------------------

def stack_by_sorted_name(variables: Mapping[str, torch.Tensor]) -> torch.Tensor:
    """Concatenate [G,B,c] blocks in sorted-name order (reference: model_utils.py:649-652)."""
    return torch.cat([variables[k] for k in sorted(variables.keys())], dim=-1)


def denoiser_forward(p: Params, g: Mapping[str, np.ndarray], arch: Mapping[str, int],
                     grid_node_features, sigma, dtype):
    """One network evaluation F(features, sigma) -> [G, B, n_out].

    Reference: gencast/denoiser.py:172-202 (sigma encoding) and :303-341
    (encoder, processor, decoder).  `grid_node_features` is the already stacked
    [G, B, C_data] tensor of denoiser.py:794-806.
    """
    x = grid_node_features.to(dtype)
    cond = noise_level_encoder(p, sigma, dtype)
    mesh, grid = grid2mesh_gnn(p, g, x, cond, dtype)
    mesh = mesh_transformer(p, g["khop"], mesh, cond, arch["num_layers"], arch["num_heads"], dtype)
    return mesh2grid_gnn(p, g, mesh, grid, cond, dtype)


def c_in(sigma):
    """Reference: gencast/dpm_solver_plus_plus_2s.py:181-182 (sigma_data = 1)."""
    return (sigma ** 2 + 1.0) ** -0.5


def c_out(sigma):
    """Reference: gencast/dpm_solver_plus_plus_2s.py:184-185."""
    return sigma / ((sigma ** 2 + 1.0) ** 0.5)


def c_skip(sigma):
    """Reference: gencast/dpm_solver_plus_plus_2s.py:187-188."""
    return 1.0 / (sigma ** 2 + 1.0)


def assemble_features(inputs_stacked, forcings: Mapping[str, torch.Tensor],
                      noisy_targets: Mapping[str, torch.Tensor]):
    """[inputs | sorted(forcings U noisy targets)] (reference: denoiser.py:184, :794-797)."""
    merged = dict(forcings)
    merged.update(noisy_targets)
    return torch.cat([inputs_stacked, stack_by_sorted_name(merged)], dim=-1)


def preconditioned_denoiser(p, g, arch, inputs_stacked, forcings, noisy_targets: Mapping[str, torch.Tensor],
                            sigma, dtype, network_fn=None):
    """D(x, sigma) = c_out F(c_in x, sigma) + c_skip x, per target variable.

    Reference: gencast/dpm_solver_plus_plus_2s.py:190-205.  noisy_targets maps
    target name -> [G,B,c]; the network output channels are split back in
    sorted-name order (model_utils.py:687-725).
    """
    sig = sigma.to(dtype)
    scaled = {k: v.to(dtype) * c_in(sig)[None, :, None] for k, v in noisy_targets.items()}
    feats = assemble_features(inputs_stacked.to(dtype), {k: v.to(dtype) for k, v in forcings.items()}, scaled)
    # network_fn(features [G,B,C], sigma [B]) -> [G,B,n_out] replaces the GenCast network: used to pin this function
    # and the solver loop against the reference's own sampler code run around a toy network
    # (tools/make_sampler_golden.py)
    raw = denoiser_forward(p, g, arch, feats, sig, dtype) if network_fn is None else network_fn(feats, sig)
    out, i = {}, 0
    for k in sorted(noisy_targets.keys()):
        c = noisy_targets[k].shape[-1]
        out[k] = raw[..., i:i + c] * c_out(sig)[None, :, None] + noisy_targets[k].to(dtype) * c_skip(sig)[None, :, None]
        i += c
    return out


def rho_inverse_cdf(min_value, max_value, rho, cdf):
    """Reference: gencast/samplers_utils.py:350-383."""
    return (min_value ** (1 / rho) + cdf * (max_value ** (1 / rho) - min_value ** (1 / rho))) ** rho


def noise_schedule(max_noise_level=80.0, min_noise_level=0.03, num_noise_levels=20, rho=7.0) -> np.ndarray:
    """Descending noise levels with a trailing zero (reference: gencast/samplers_utils.py:395-412)."""
    levels = rho_inverse_cdf(min_noise_level, max_noise_level, rho, np.linspace(1, 0, num_noise_levels))
    return np.append(levels, 0.0)


def stochastic_churn_rate_schedule(noise_levels, stochastic_churn_rate=0.0, churn_min_noise_level=0.05,
                                   churn_max_noise_level=50.0) -> np.ndarray:
    """Per-step churn rates (reference: gencast/samplers_utils.py:414-431)."""
    noise_levels = np.asarray(noise_levels, np.float64)
    n = len(noise_levels) - 1
    per_step = min(stochastic_churn_rate / n, np.sqrt(2) - 1)
    return ((churn_min_noise_level <= noise_levels[:-1]) & (noise_levels[:-1] <= churn_max_noise_level)) * per_step


def apply_stochastic_churn(x: Mapping[str, torch.Tensor], noise_level: float, churn_rate: float, inflation: float,
                           unit_noise: Mapping[str, torch.Tensor]):
    """x at a higher noise level, and that level (reference: gencast/samplers_utils.py:434-452; `unit_noise` stands for
    its spherical_white_noise_like draw)."""
    new_level = noise_level * (1.0 + churn_rate)
    extra = math.sqrt(max(new_level ** 2 - noise_level ** 2, 0.0)) * inflation
    return {k: v + unit_noise[k].to(v.dtype) * extra for k, v in x.items()}, new_level


def dpm_solver_2s(p, g, arch, inputs_stacked, forcings, init_x: Mapping[str, torch.Tensor],
                  sigmas: Sequence[float], dtype, num_steps: Optional[int] = None,
                  trace: Optional[list] = None, network_fn=None, churn_rates: Optional[Sequence[float]] = None,
                  inflation: float = 1.0, churn_noise: Optional[Sequence[Mapping[str, torch.Tensor]]] = None):
    """Deterministic DPM-Solver++ 2S loop (stochastic churn = 0).

    Reference: gencast/dpm_solver_plus_plus_2s.py:120-158.  With `churn_rates` (per step, :38-43) the state is first
    moved to a higher noise level with fresh noise (:127-137; the reference calls an `_arr` variant of
    apply_stochastic_churn that it does not define -- the documented function, samplers_utils.py:434-452, is what is
    restated); `churn_noise[k]` is the unit-variance draw of the k-th churned step.  init_x already holds
    noise * sigmas[0] (:78).  sigma is clamped to >= 1e-6 before each denoiser
    call (:85).  On the last iteration (sigma_next == 0) the reference still
    evaluates the second denoiser call and discards it (:148-153); it is skipped
    here as it cannot affect the result.
    """
    x = {k: v.to(dtype) for k, v in init_x.items()}
    B = next(iter(x.values())).shape[1]
    n = len(sigmas) - 1 if num_steps is None else num_steps
    def D(state, s):
        s_safe = max(float(s), 1e-6)
        return preconditioned_denoiser(p, g, arch, inputs_stacked, forcings, state,
                                       torch.full((B,), s_safe, dtype=dtype), dtype, network_fn)
    k_churn = 0
    for i in range(n):
        sigma, sigma_next = float(sigmas[i]), float(sigmas[i + 1])
        if churn_rates is not None and churn_rates[i] > 0:
            x, sigma = apply_stochastic_churn(x, sigma, float(churn_rates[i]), inflation, churn_noise[k_churn])
            k_churn += 1
        sigma_mid = math.sqrt(sigma * sigma_next)
        den = D(x, sigma)
        if sigma_next == 0:
            x = den
        else:
            a = sigma_mid / sigma
            x_mid = {k: a * x[k] + (1 - a) * den[k] for k in x}
            den_mid = D(x_mid, sigma_mid)
            b = sigma_next / sigma
            x = {k: b * x[k] + (1 - b) * den_mid[k] for k in x}
        if trace is not None:
            trace.append({k: v.clone() for k, v in x.items()})
    return x
