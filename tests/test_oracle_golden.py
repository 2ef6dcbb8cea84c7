"""Pins the CPU oracle (oracle/gencast_oracle.py) against golden vectors produced by the
reference's own module code (tools/make_refshim_golden.py: /root/reference's mlp / typed_graph_net /
deep_typed_graph_net / sparse_transformer / transformer executed under numpy stand-ins for the JAX
API).  Runs on CPU everywhere; when the reference tree is present the fixture is also regenerated
and compared with the committed file."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "refshim_tiny.npz")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _case():
    import make_refshim_golden as m
    return m.build_case()


def _graph_dict(g):
    return dict(g2m_grid_feat=g.g2m_grid_feat, g2m_mesh_feat=g.g2m_mesh_feat, g2m_edge_feat=g.g2m_edge_feat,
                g2m_senders=g.g2m_senders, g2m_receivers=g.g2m_receivers, m2g_senders=g.m2g_senders,
                m2g_receivers=g.m2g_receivers, m2g_edge_feat=g.m2g_edge_feat, khop=g.khop)


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-30))


def test_oracle_matches_reference_module_outputs():
    from oracle import gencast_oracle as o
    arch, g, p, feats, cond, c_data, n_out = _case()
    gold = np.load(GOLDEN)
    np.testing.assert_array_equal(gold["features"], feats)
    gd = _graph_dict(g)
    st = arch.sparse_transformer_config
    assert int(gold["mask_block_size"]) == o.mask_block_size(g.khop)
    dt = torch.float64
    for sigma, c in cond.items():
        tag = f"s{sigma:g}"
        np.testing.assert_array_equal(gold[f"{tag}/cond"], c)
        ct = torch.as_tensor(c).to(dt)
        mesh, grid = o.grid2mesh_gnn(p, gd, torch.as_tensor(feats).to(dt), ct, dt)
        # the reference aggregates grid2mesh messages in float32 (gencast/denoiser.py:371); fixture latents are f32
        assert _rel(mesh.numpy(), gold[f"{tag}/mesh_latent"]) < 2e-6
        assert _rel(grid.numpy(), gold[f"{tag}/grid_latent"]) < 2e-6
        mesh_out = o.mesh_transformer(p, g.khop, mesh, ct, st.num_layers, st.num_heads, dt)
        assert _rel(mesh_out.numpy(), gold[f"{tag}/mesh_out"]) < 5e-6
        y = o.mesh2grid_gnn(p, gd, mesh_out, grid, ct, dt)
        assert _rel(y.numpy(), gold[f"{tag}/output"]) < 5e-6


@pytest.mark.skipif(not os.path.isdir(os.environ.get("GENCAST_REFERENCE", "/root/reference")),
                    reason="reference tree not present (GPU box)")
def test_committed_fixture_is_what_the_reference_produces():
    import make_refshim_golden as m
    arch, g, p, feats, cond, c_data, n_out = m.build_case()
    fresh = m.run_reference(arch, g, p, feats, cond, c_data, n_out)
    gold = np.load(GOLDEN)
    for k, v in fresh.items():
        assert _rel(v, gold[k]) < 1e-6, k


# ----------------------------------------------------------------------------------------------------------------
# Noise-level encoder, sampler body, preconditioning and stacking: tests/golden/refshim_sampler.npz is produced by the
# reference's own FourierFeaturesMLP / Sampler.__call__ / model_utils stacking code (tools/make_sampler_golden.py)
# ----------------------------------------------------------------------------------------------------------------
SAMPLER_GOLDEN = os.path.join(ROOT, "tests", "golden", "refshim_sampler.npz")


def _sampler_case():
    import make_sampler_golden as m
    from gencast_flax_nnx_b200 import stacking
    inputs, targets, forcings, noise = m.build_case()
    sizes = dict(targets.sizes)
    inp_nodes, _ = stacking.dataset_to_nodes(inputs, sizes)
    frc_nodes, frc_layout = stacking.dataset_to_nodes(forcings, sizes)
    frc, i = {}, 0
    for n, c in frc_layout:
        frc[n] = torch.as_tensor(frc_nodes[:, :, i:i + c]).double()
        i += c
    tnames = sorted(targets.keys())
    to_nodes = lambda arrs: {n: torch.as_tensor(stacking.variable_to_nodes(type(targets[n])(arrs[n], targets[n].dims), sizes)).double()
                             for n in tnames}
    return m, inputs, targets, forcings, noise, torch.as_tensor(inp_nodes).double(), frc, to_nodes, tnames, sizes


def test_oracle_noise_encoder_matches_reference_fourier_features_mlp():
    """oracle.noise_level_encoder == the reference's FourierFeaturesMLP.__call__ (common/mlp.py:255-265) on
    sigma in {80, 7.5, 1, 0.03, 1e-6}."""
    from oracle import gencast_oracle as o
    gold = np.load(SAMPLER_GOLDEN)
    p = {f"denoiser/noise_level_encoder/{k}": gold[f"encoder/{k}"] for k in
         ("linear_0/kernel", "linear_0/bias", "linear_1/kernel", "linear_1/bias")}
    cond = o.noise_level_encoder(p, torch.as_tensor(gold["encoder/sigmas"]), torch.float64)
    assert _rel(cond.numpy(), gold["encoder/cond"]) < 1e-12


def test_host_stacking_matches_reference_dataset_to_stacked():
    """stacking.dataset_to_nodes == lat_lon_to_leading_axes(dataset_to_stacked(.)) of the reference
    (common/model_utils.py:145-151, :626-659), bit for bit."""
    from gencast_flax_nnx_b200 import stacking
    gold = np.load(SAMPLER_GOLDEN)
    m, inputs, targets, forcings, noise, inp_nodes, frc, to_nodes, tnames, sizes = _sampler_case()
    np.testing.assert_array_equal(inp_nodes.numpy(), gold["stacking/inputs_nodes"])
    assert [n for n, _ in stacking.channel_layout(inputs)] == [str(s) for s in gold["stacking/input_names_sorted"]]


def test_oracle_sampler_matches_reference_sampler_body():
    """oracle.dpm_solver_2s / preconditioned_denoiser / assemble_features / noise_schedule against the output of the
    reference's own Sampler.__call__ (body_fn, denoise_arr, _preconditioned_denoiser;
    gencast/dpm_solver_plus_plus_2s.py:47-205) run around the same toy network, and engine.noise_schedule."""
    from oracle import gencast_oracle as o
    from gencast_flax_nnx_b200.engine import noise_schedule
    gold = np.load(SAMPLER_GOLDEN)
    m, inputs, targets, forcings, noise, inp_nodes, frc, to_nodes, tnames, sizes = _sampler_case()
    sig = gold["sampler/noise_levels"]
    np.testing.assert_allclose(o.noise_schedule(80.0, 0.03, m.NUM_LEVELS, 7.0), sig, rtol=1e-14)
    np.testing.assert_allclose(noise_schedule(80.0, 0.03, m.NUM_LEVELS, 7.0), sig, rtol=1e-14)
    assert not gold["sampler/churn_rates"].any()                    # churn rate 0 -> no stochastic churn
    from gencast_flax_nnx_b200.engine import stochastic_churn_rate_schedule
    for fn in (stochastic_churn_rate_schedule, o.stochastic_churn_rate_schedule):
        np.testing.assert_allclose(fn(sig, 2.5, 0.75, float("inf")), gold["sampler/churn_schedule_rate2p5"], rtol=1e-14)
    w = torch.as_tensor(gold["sampler/toy_w"])

    def net(feats, sigma):
        return torch.as_tensor(m.toy_network(feats.numpy(), sigma.numpy(), w.numpy()))

    for n in tnames:
        np.testing.assert_array_equal(noise[n], gold[f"sampler/noise/{n}"])
    init = {k: v * float(sig[0]) for k, v in to_nodes(noise).items()}
    # the first network call of the reference sees [inputs | sorted(forcings U c_in x)] (gencast/denoiser.py:184,:794-797)
    first = o.assemble_features(inp_nodes, frc, {k: v * float(o.c_in(torch.tensor(sig[0]))) for k, v in init.items()})
    assert _rel(first.numpy(), gold["sampler/first_call_features"]) < 1e-13
    res = o.dpm_solver_2s({}, {}, {}, inp_nodes, frc, init, sig, torch.float64, network_fn=net)
    want = to_nodes({n: gold[f"sampler/result/{n}"] for n in tnames})
    for n in tnames:
        assert _rel(res[n].numpy(), want[n].numpy()) < 1e-10, n
    # one preconditioned call with a different noise level per batch element (dpm...py:190-205)
    noisy = to_nodes({n: noise[n] * 2.0 for n in tnames})
    d = o.preconditioned_denoiser({}, {}, {}, inp_nodes, frc, noisy, torch.tensor([3.0, 0.2], dtype=torch.float64),
                                  torch.float64, network_fn=net)
    want = to_nodes({n: gold[f"precond/result/{n}"] for n in tnames})
    for n in tnames:
        assert _rel(d[n].numpy(), want[n].numpy()) < 1e-12, n


@pytest.mark.skipif(not os.path.isdir(os.environ.get("GENCAST_REFERENCE", "/root/reference")),
                    reason="reference tree not present (GPU box)")
def test_committed_sampler_fixture_is_what_the_reference_produces():
    import subprocess
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import make_sampler_golden as m; out = m.run_reference(); "
            "gold = np.load(%r, allow_pickle=False); "
            "bad = [k for k, v in out.items() if np.asarray(v).dtype.kind == 'f' and "
            "np.abs(np.asarray(v) - gold[k]).max() > 1e-12 * max(1.0, np.abs(gold[k]).max())]; "
            "assert not bad and set(out) == set(gold.files), bad") % (os.path.join(ROOT, "tools"), SAMPLER_GOLDEN)
    # a fresh interpreter: the stand-in packages shadow module names (xarray, jax) for the whole process
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
