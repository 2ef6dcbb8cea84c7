"""Shared builders for the parity tests: one seeded case = graphs + perturbed weights +
synthetic inputs, in both the oracle's and the engine's formats."""
from __future__ import annotations

import dataclasses
import functools

import numpy as np
import torch

from gencast_flax_nnx_b200 import configs, graph, params, stacking, synthetic
from gencast_flax_nnx_b200.engine import ChannelLayout


@dataclasses.dataclass
class Case:
    name: str
    arch: configs.DenoiserArchitectureConfig
    graphs: graph.DenoiserGraphs
    layout: ChannelLayout
    params: dict
    inputs: object
    targets: object
    forcings: object
    inp_nodes: np.ndarray      # [G, 1, C_in]
    frc_nodes: np.ndarray      # [G, 1, C_f]
    frc_vars: dict             # name -> [G, 1, c]
    target_vars: list          # [(name, channels)] sorted

    @property
    def oracle_graph(self):
        g = self.graphs
        return dict(g2m_grid_feat=g.g2m_grid_feat, g2m_mesh_feat=g.g2m_mesh_feat, g2m_edge_feat=g.g2m_edge_feat,
                    g2m_senders=g.g2m_senders, g2m_receivers=g.g2m_receivers, m2g_senders=g.m2g_senders,
                    m2g_receivers=g.m2g_receivers, m2g_edge_feat=g.m2g_edge_feat, khop=g.khop)

    @property
    def oracle_arch(self):
        st = self.arch.sparse_transformer_config
        return dict(num_layers=st.num_layers, num_heads=st.num_heads)

    def split_targets(self, nodes):
        """[G, B, n_out] -> {name: [G, B, c]} in sorted-name order."""
        out, i = {}, 0
        for n, c in self.target_vars:
            out[n] = nodes[:, :, i:i + c]
            i += c
        return out


@functools.lru_cache(maxsize=4)
def make_case(name: str, seed: int = 0) -> Case:
    res, arch = configs.named_config(name)
    lat, lon = graph.regular_grid(res)
    st = arch.sparse_transformer_config
    graphs = graph.build_denoiser_graphs(lat, lon, arch.mesh_size, st.attention_k_hop)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=1, seed=seed)
    sizes = dict(targets.sizes)
    inp_nodes, _ = stacking.dataset_to_nodes(inputs, sizes)
    frc_nodes, frc_layout = stacking.dataset_to_nodes(forcings, sizes)
    frc_vars, i = {}, 0
    for n, c in frc_layout:
        frc_vars[n] = frc_nodes[:, :, i:i + c]
        i += c
    layout = ChannelLayout(num_input_channels=inp_nodes.shape[-1], forcing_vars=tuple(frc_layout),
                           target_vars=tuple(stacking.channel_layout(targets)))
    shapes = params.param_shapes(arch, layout.num_data_channels, layout.num_targets)
    p = params.init_perturbed(shapes, seed=1)
    return Case(name, arch, graphs, layout, p, inputs, targets, forcings, inp_nodes, frc_nodes, frc_vars,
                sorted(stacking.channel_layout(targets)))


def oracle_forward(case: Case, scaled_noisy: np.ndarray, sigma: float, dtype=torch.float64) -> np.ndarray:
    """Raw network output F for scaled noisy targets [G, n_out] -> [G, n_out]."""
    from oracle import gencast_oracle as o
    noisy = case.split_targets(torch.as_tensor(scaled_noisy[:, None, :]).to(dtype))
    frc = {k: torch.as_tensor(v).to(dtype) for k, v in case.frc_vars.items()}
    feats = o.assemble_features(torch.as_tensor(case.inp_nodes).to(dtype), frc, noisy)
    out = o.denoiser_forward(case.params, case.oracle_graph, case.oracle_arch, feats,
                             torch.tensor([sigma], dtype=dtype), dtype)
    return out[:, 0, :].numpy()


def per_variable_error(case: Case, got: np.ndarray, ref: np.ndarray) -> dict:
    """max|got - ref| / max|ref| per output variable (SURVEY.md §8c metric)."""
    errs, i = {}, 0
    for n, c in case.target_vars:
        r, g = ref[:, i:i + c], got[:, i:i + c]
        errs[n] = float(np.abs(g - r).max() / max(np.abs(r).max(), 1e-30))
        i += c
    return errs
