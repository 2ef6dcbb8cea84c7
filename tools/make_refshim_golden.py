#!/usr/bin/env python
"""Generates tests/golden/refshim_tiny.npz by executing the REFERENCE'S OWN module code
(/root/reference: common/mlp.py, common/typed_graph_net.py, common/deep_typed_graph_net.py,
gencast/sparse_transformer.py, gencast/transformer.py, common/typed_graph.py) under the numpy
stand-ins of tools/refshim/ for jax / flax.nnx / jraph / chex.

The three networks are constructed exactly as gencast/denoiser.py:365-414 constructs them and fed
exactly as gencast/denoiser.py:602-768 feeds them (that file itself needs xarray, so its few lines
of graph assembly are restated below with line citations).  Parameters are overwritten with the
seeded perturbed weights of gencast_flax_nnx_b200.params (paths = NNX attribute paths).  All
arithmetic is float64.

Only runnable where /root/reference exists (the build container).  The fixture it writes travels
with the repository; tests/test_oracle_golden.py checks the oracle against it everywhere and, when
the reference is present, regenerates it in memory and checks the committed file is current.
"""
from __future__ import annotations

import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("GENCAST_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

SIGMAS = (80.0, 1.0, 0.03)


def build_case():
    from gencast_flax_nnx_b200 import configs, graph, params
    res, arch = configs.named_config("tiny")
    lat, lon = graph.regular_grid(res)
    st = arch.sparse_transformer_config
    g = graph.build_denoiser_graphs(lat, lon, arch.mesh_size, st.attention_k_hop)
    c_data, n_out = 20, 7          # small data widths keep the fixture small; the wiring does not depend on them
    shapes = params.param_shapes(arch, c_data, n_out)
    p = params.init_perturbed(shapes, seed=1)
    rng = np.random.default_rng(11)
    feats = rng.standard_normal((g.num_grid_nodes, 1, c_data))
    cond = {s: rng.standard_normal((1, 16)) * 0.5 for s in SIGMAS}      # stands for the noise-level encoding
    return arch, g, p, feats, cond, c_data, n_out


def run_reference(arch, g, p, feats, cond, c_data, n_out):
    """Returns {name: array} of stage outputs computed by the reference's modules."""
    sys.path.insert(0, os.path.join(ROOT, "tools", "refshim"))
    sys.path.insert(0, REFERENCE)
    import flax.nnx as nnx                                   # the shim
    from common import deep_typed_graph_net, typed_graph
    from gencast import transformer
    from scipy import sparse

    rngs = nnx.Rngs(0)
    L = arch.latent_size
    G, V = g.num_grid_nodes, g.num_mesh_nodes
    f64 = np.float64

    def batch_second(x):                                      # gencast/denoiser.py:833-837
        return np.repeat(np.asarray(x, f64)[:, None, :], 1, axis=1)

    # --- graph templates (gencast/denoiser.py:443-600); node features = [struct | dummy data zeros]
    g2m_key = typed_graph.EdgeSetKey("grid2mesh", ("grid_nodes", "mesh_nodes"))
    g2m_tpl = typed_graph.TypedGraph(
        context=typed_graph.Context(n_graph=np.array([1]), features=()),
        nodes={"grid_nodes": typed_graph.NodeSet(n_node=np.array([G]),
                                                 features=np.concatenate([g.g2m_grid_feat, np.zeros((G, c_data), np.float32)], -1)),
               "mesh_nodes": typed_graph.NodeSet(n_node=np.array([V]),
                                                 features=np.concatenate([g.g2m_mesh_feat, np.zeros((V, c_data), np.float32)], -1))},
        edges={g2m_key: typed_graph.EdgeSet(n_edge=np.array([len(g.g2m_senders)]),
                                            indices=typed_graph.EdgesIndices(senders=g.g2m_senders, receivers=g.g2m_receivers),
                                            features=g.g2m_edge_feat)})
    ms, mr = __import__("gencast_flax_nnx_b200.graph", fromlist=["faces_to_edges"]).faces_to_edges(g.mesh.faces)
    mesh_key = typed_graph.EdgeSetKey("mesh", ("mesh_nodes", "mesh_nodes"))
    mesh_tpl = typed_graph.TypedGraph(
        context=typed_graph.Context(n_graph=np.array([1]), features=()),
        nodes={"mesh_nodes": typed_graph.NodeSet(n_node=np.array([V]), features=None)},
        edges={mesh_key: typed_graph.EdgeSet(n_edge=np.array([len(ms)]),
                                             indices=typed_graph.EdgesIndices(senders=ms, receivers=mr),
                                             features=np.zeros((len(ms), 4), np.float32))})
    m2g_key = typed_graph.EdgeSetKey("mesh2grid", ("mesh_nodes", "grid_nodes"))
    m2g_tpl = typed_graph.TypedGraph(
        context=typed_graph.Context(n_graph=np.array([1]), features=()),
        nodes={"grid_nodes": typed_graph.NodeSet(n_node=np.array([G]), features=None),
               "mesh_nodes": typed_graph.NodeSet(n_node=np.array([V]), features=None)},
        edges={m2g_key: typed_graph.EdgeSet(n_edge=np.array([len(g.m2g_senders)]),
                                            indices=typed_graph.EdgesIndices(senders=g.m2g_senders, receivers=g.m2g_receivers),
                                            features=g.m2g_edge_feat)})

    # --- networks, as gencast/denoiser.py:365-414
    g2m_gnn = deep_typed_graph_net.DeepTypedGraphNet(
        activation="swish", aggregate_normalization=None, edge_latent_size=dict(grid2mesh=L), embed_edges=True,
        embed_nodes=True, f32_aggregation=True, include_sent_messages_in_node_update=False, mlp_hidden_size=L,
        mlp_num_hidden_layers=1, node_latent_size=dict(grid_nodes=L, mesh_nodes=L), node_output_size=None,
        num_message_passing_steps=1, use_layer_norm=True, use_norm_conditioning=True, rngs=rngs, gpu_mesh=None,
        graph_template=g2m_tpl)
    st_kwargs = dataclasses.asdict(arch.sparse_transformer_config)
    mesh_gnn = transformer.MeshTransformer(transformer_kwargs=st_kwargs, rngs=rngs, gpu_mesh=None, graph_template=mesh_tpl)
    m2g_gnn = deep_typed_graph_net.DeepTypedGraphNet(
        activation="swish", edge_latent_size=dict(mesh2grid=L), embed_nodes=False, f32_aggregation=False,
        include_sent_messages_in_node_update=False, mlp_hidden_size=L, mlp_num_hidden_layers=1,
        node_latent_size=dict(grid_nodes=L, mesh_nodes=L), node_output_size=dict(grid_nodes=n_out),
        num_message_passing_steps=1, use_layer_norm=True, use_norm_conditioning=True, rngs=rngs, gpu_mesh=None,
        graph_template=m2g_tpl)

    out = {}
    first = True
    for sigma, c in cond.items():
        # --- _run_grid2mesh_gnn (gencast/denoiser.py:602-688)
        grid_in = np.concatenate([batch_second(g.g2m_grid_feat), feats], -1)
        mesh_in = np.concatenate([batch_second(g.g2m_mesh_feat), np.zeros((V, 1, c_data))], -1)
        graph_in = g2m_tpl._replace(
            edges={g2m_key: g2m_tpl.edges[g2m_key]._replace(features=batch_second(g.g2m_edge_feat))},
            nodes={"grid_nodes": g2m_tpl.nodes["grid_nodes"]._replace(features=grid_in),
                   "mesh_nodes": g2m_tpl.nodes["mesh_nodes"]._replace(features=mesh_in)})
        if first:
            g2m_gnn(graph_in, c)             # lazy parameter creation (deep_typed_graph_net.py:499-501)
        # --- _run_mesh2grid_gnn needs latents to create its params: run once with zeros
            m2g_in0 = m2g_tpl._replace(
                edges={m2g_key: m2g_tpl.edges[m2g_key]._replace(features=batch_second(g.m2g_edge_feat))},
                nodes={"mesh_nodes": m2g_tpl.nodes["mesh_nodes"]._replace(features=np.zeros((V, 1, L))),
                       "grid_nodes": m2g_tpl.nodes["grid_nodes"]._replace(features=np.zeros((G, 1, L)))})
            m2g_gnn(m2g_in0, c)
            for mod, pre in ((g2m_gnn, "denoiser/predictor/grid2mesh_gnn"), (mesh_gnn, "denoiser/predictor/mesh_gnn"),
                             (m2g_gnn, "denoiser/predictor/mesh2grid_gnn")):
                seen, expected = load_params(mod, pre, p)
                # every parameter of the reference tree is fed from the flat dict and vice versa
                assert seen == expected, (pre, seen, expected)
            first = False
        o = g2m_gnn(graph_in, c)
        mesh_lat, grid_lat = o.nodes["mesh_nodes"].features, o.nodes["grid_nodes"].features
        # --- _run_mesh_gnn (gencast/denoiser.py:691-728)
        mg = mesh_tpl._replace(
            edges={mesh_key: mesh_tpl.edges[mesh_key]._replace(features=batch_second(mesh_tpl.edges[mesh_key].features))},
            nodes={"mesh_nodes": mesh_tpl.nodes["mesh_nodes"]._replace(features=mesh_lat)})
        mesh_out = mesh_gnn(mg, global_norm_conditioning=c).nodes["mesh_nodes"].features
        # --- _run_mesh2grid_gnn (gencast/denoiser.py:730-768)
        dg = m2g_tpl._replace(
            edges={m2g_key: m2g_tpl.edges[m2g_key]._replace(features=batch_second(g.m2g_edge_feat))},
            nodes={"mesh_nodes": m2g_tpl.nodes["mesh_nodes"]._replace(features=mesh_out),
                   "grid_nodes": m2g_tpl.nodes["grid_nodes"]._replace(features=grid_lat)})
        y = m2g_gnn(dg, c).nodes["grid_nodes"].features
        tag = f"s{sigma:g}"
        out[f"{tag}/mesh_latent"], out[f"{tag}/grid_latent"] = np.asarray(mesh_lat), np.asarray(grid_lat)
        out[f"{tag}/mesh_out"], out[f"{tag}/output"] = np.asarray(mesh_out), np.asarray(y)
    out["mask_block_size"] = np.asarray(mesh_gnn.batch_first_transformer._cfg.mask_block_size)
    return out


def load_params(module, prefix, p):
    """Overwrite every parameter of a reference module tree from the flat path -> array dict."""
    import re
    seen = 0
    for path, param in module.named_params(prefix):
        # NNX wraps update functions (EdgeWrapper / NodeWrapper hold `edge_fn` / `node_fn`), Sequential
        # holds `.layers`; the flat keys follow the same attribute names.
        key = path
        if key not in p:
            raise KeyError(f"reference parameter {path!r} (shape {param.value.shape}) has no counterpart; "
                           f"close keys: {[k for k in p if k.startswith(prefix)][:3]}")
        if p[key].shape != param.value.shape:
            raise ValueError(f"{path}: shape {param.value.shape} vs {p[key].shape}")
        param.value = np.asarray(p[key], np.float64)
        seen += 1
    expected = [k for k in p if k.startswith(prefix + "/")]
    return seen, len(expected)


def main():
    arch, g, p, feats, cond, c_data, n_out = build_case()
    out = run_reference(arch, g, p, feats, cond, c_data, n_out)
    out["features"] = feats
    for s, c in cond.items():
        out[f"s{s:g}/cond"] = c
    out["meta"] = np.asarray([c_data, n_out])
    dst = os.path.join(ROOT, "tests", "golden", "refshim_tiny.npz")
    np.savez_compressed(dst, **{k: np.asarray(v, np.float32 if np.asarray(v).dtype == np.float64 and k.endswith(("latent", "mesh_out")) else None) for k, v in out.items()})
    print(f"wrote {dst}: {os.path.getsize(dst) / 1e6:.2f} MB; keys: {sorted(out)[:6]} ...")


if __name__ == "__main__":
    main()
