"""Parity of the CUDA denoiser and sampler against the CPU oracle (fp64 restatement of the
reference, oracle/gencast_oracle.py) on seeded inputs and perturbed weights.

Tolerances (BASELINE.json north_star / BASELINE.md §4): fp32 path max relative error per
variable <= 1e-3 for one denoiser call; bf16 path <= 2e-2 against the same oracle.
"""
import numpy as np
import pytest
import torch

from helpers import make_case, oracle_forward, per_variable_error

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-3, "bf16": 2e-2}


def _engine(case, dtype):
    from gencast_flax_nnx_b200.engine import DenoiserEngine
    return DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype=dtype)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("name,sigmas", [("tiny", (80.0, 1.0, 0.03)), ("nano", (1.0,))])
def test_single_forward_matches_oracle(cuda_device, name, sigmas, dtype):
    case = make_case(name)
    eng = _engine(case, dtype)
    rng = np.random.default_rng(2)
    eng.set_constant_features(case.inp_nodes[:, 0], case.frc_nodes[:, 0])
    for sigma in sigmas:
        x = rng.standard_normal((eng.G, eng.n_out)).astype(np.float32)
        eng.set_network_input(x)
        got = eng.forward(eng.sigma_context(sigma))[:, :eng.n_out].cpu().numpy()
        ref = oracle_forward(case, x, sigma)
        errs = per_variable_error(case, got, ref)
        worst = max(errs.values())
        print(f"{name} {dtype} sigma={sigma}: worst per-variable error {worst:.3e}")
        assert worst <= TOL[dtype], errs


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_sampler_matches_oracle(cuda_device, dtype):
    """Full DPM-Solver++ 2S loop on the tiny case with a 4-level schedule (7 network evaluations)."""
    from gencast_flax_nnx_b200.engine import SamplerEngine, noise_schedule
    from oracle import gencast_oracle as o
    case = make_case("tiny")
    eng = _engine(case, dtype)
    sigmas = noise_schedule(80.0, 0.03, 4, 7.0)
    np.testing.assert_allclose(sigmas, o.noise_schedule(80.0, 0.03, 4, 7.0))
    eng.set_constant_features(case.inp_nodes[:, 0], case.frc_nodes[:, 0])
    noise = np.random.default_rng(3).standard_normal((eng.G, eng.n_out)).astype(np.float32)
    results = {}
    for discard in (True, False):
        se = SamplerEngine(eng, sigmas, evaluate_discarded_call=discard)
        eager = se.sample(noise, use_graph=False).cpu().numpy().copy()
        graphed = se.sample(noise, use_graph=True).cpu().numpy().copy()
        again = se.sample(noise, use_graph=True).cpu().numpy().copy()
        assert np.array_equal(eager, graphed) and np.array_equal(graphed, again)   # deterministic, graph == eager
        results[discard] = eager
    assert np.array_equal(results[True], results[False])       # the discarded call cannot change the result
    dt = torch.float64
    init = case.split_targets(torch.as_tensor(noise[:, None, :] * sigmas[0]).to(dt))
    frc = {k: torch.as_tensor(v).to(dt) for k, v in case.frc_vars.items()}
    ref = o.dpm_solver_2s(case.params, case.oracle_graph, case.oracle_arch, torch.as_tensor(case.inp_nodes).to(dt),
                          frc, init, sigmas, dt)
    ref = torch.cat([ref[n] for n, _ in case.target_vars], dim=-1)[:, 0].numpy()
    errs = per_variable_error(case, results[True], ref)
    print(f"sampler {dtype}: worst per-variable error {max(errs.values()):.3e}")
    assert max(errs.values()) <= (2e-3 if dtype == "f32" else 5e-2), errs


def test_public_api_round_trip(cuda_device):
    """GenCast.full_sampling / Denoiser.__call__ through Datasets equal the engine on arrays."""
    from gencast_flax_nnx_b200 import configs, gencast, stacking
    from gencast_flax_nnx_b200.rngs import Rngs
    from gencast_flax_nnx_b200.xarray_lite import DataArray
    case = make_case("tiny")
    res, arch = configs.named_config("tiny")
    sc = configs.SamplerConfig(num_noise_levels=3, stochastic_churn_rate=0.0)
    model = gencast.GenCast(configs.TASK, arch, sampler_config=sc, rngs=Rngs(0), params=case.params,
                            compute_dtype="f32")
    noisy = np.random.default_rng(5).standard_normal((case.graphs.num_grid_nodes, 1, 82)).astype(np.float32)
    noisy_ds = stacking.nodes_to_dataset(noisy, case.targets)
    out = model.denoiser(case.inputs, noisy_ds, DataArray(np.array([1.0], np.float32), ("batch",)), case.forcings)
    got, _ = stacking.dataset_to_nodes(out, dict(case.targets.sizes))
    ref = oracle_forward(case, noisy[:, 0], 1.0)
    assert max(per_variable_error(case, got[:, 0], ref).values()) <= 1e-3
    with pytest.raises(ValueError):
        model.denoiser(case.inputs, noisy_ds, DataArray(np.ones((1, 1), np.float32), ("batch", "x")), case.forcings)
    pred = model.full_sampling(case.inputs, case.targets, case.forcings)
    assert set(pred.keys()) == set(case.targets.keys())
    for k in pred.keys():
        assert pred[k].dims == case.targets[k].dims and pred[k].shape == case.targets[k].shape
        assert np.isfinite(pred[k].data).all()
    with pytest.raises(ValueError):
        model._sampler(case.inputs, case.targets, case.forcings, rngs=None)


def test_full_size_1deg_properties(cuda_device):
    """GenCast 1 deg (BASELINE.json configs[2]: 181x360 grid, mesh 5, L=512, 16 layers), where the CPU
    oracle is too slow to run in a test: size-independent properties instead.
      * the bf16 tensor-core path (tcgen05 GEMMs incl. the CTA-pair and TMA-store variants, tile
        attention) agrees with the fp32 path (FFMA GEMM, CSR attention) -- two disjoint kernel sets,
        the second of which is pinned to the oracle at nano size above;
      * relabelling the mesh (patch order vs the reference's band order) does not change the result
        beyond bf16 rounding (attention / segment sums are permutation equivariant);
      * repeated evaluation is bitwise reproducible.
    """
    from gencast_flax_nnx_b200.engine import DenoiserEngine
    case = make_case("1deg")
    rng = np.random.default_rng(4)
    x = rng.standard_normal((case.graphs.num_grid_nodes, 82)).astype(np.float32)

    def run(dtype, order):
        eng = DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype=dtype, mesh_order=order)
        eng.set_constant_features(case.inp_nodes[:, 0], case.frc_nodes[:, 0])
        eng.set_network_input(x)
        a = eng.read_output(eng.forward(eng.sigma_context(1.0)))
        b = eng.read_output(eng.forward(eng.sigma_context(1.0)))
        assert np.array_equal(a, b)
        del eng
        torch.cuda.empty_cache()
        return a

    ref32 = run("f32", "reference")
    assert np.isfinite(ref32).all()
    for order in ("patch", "reference"):
        got = run("bf16", order)
        errs = per_variable_error(case, got, ref32)
        print(f"1deg bf16 ({order} mesh order) vs f32: worst per-variable error {max(errs.values()):.3e}")
        assert max(errs.values()) <= 2e-2, errs
    got32 = run("f32", "patch")
    errs = per_variable_error(case, got32, ref32)
    print(f"1deg f32 patch vs reference mesh order: {max(errs.values()):.3e}")
    assert max(errs.values()) <= 1e-4, errs


def test_device_stacking_equals_host_stacking(cuda_device):
    """The GPU layout transposes reproduce stacking.py (common/model_utils.py:594-725) bit for bit."""
    from gencast_flax_nnx_b200 import graph, stacking, synthetic
    from gencast_flax_nnx_b200.device_stacking import DeviceStacker
    lat, lon = graph.regular_grid(30.0)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=2, seed=5)
    sizes = dict(targets.sizes)
    st = DeviceStacker(cuda_device)
    for key, ds in (("inputs", inputs), ("forcings", forcings)):
        host, _ = stacking.dataset_to_nodes(ds, sizes)
        dev = st.to_nodes(key, ds, sizes)
        assert np.array_equal(dev.cpu().numpy(), host)
    x = np.random.default_rng(0).standard_normal((len(lat) * len(lon), 2, 82)).astype(np.float32)
    ref = stacking.nodes_to_dataset(x, targets)
    got = st.from_nodes(torch.from_numpy(x).to(cuda_device), targets)
    for k in ref.keys():
        assert got[k].dims == ref[k].dims and np.array_equal(got[k].data, ref[k].data)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_batched_members_equal_single_member_runs(cuda_device, dtype):
    """Three members evaluated together (member-major row blocks, padded mesh blocks, offset index
    tables) give each member exactly what a one-member engine gives it; the full sampler too."""
    from gencast_flax_nnx_b200.engine import DenoiserEngine, SamplerEngine, noise_schedule
    case = make_case("tiny")
    B = 3
    rng = np.random.default_rng(8)
    G = case.graphs.num_grid_nodes
    x = rng.standard_normal((B, G, 82)).astype(np.float32)
    inp = np.stack([case.inp_nodes[:, 0] * (1 + 0.1 * b) for b in range(B)])
    frc = np.stack([case.frc_nodes[:, 0]] * B)
    sigmas = noise_schedule(80.0, 0.03, 3, 7.0)
    single, single_s = [], []
    for b in range(B):
        e1 = DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype=dtype)
        e1.set_constant_features(inp[b], frc[b])
        e1.set_network_input(x[b])
        single.append(e1.read_output(e1.forward(e1.sigma_context(1.0))))
        single_s.append(SamplerEngine(e1, sigmas).sample(x[b], use_graph=False).cpu().numpy().copy())
    eb = DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype=dtype, members=B)
    eb.set_constant_features(inp.reshape(B * G, -1), frc.reshape(B * G, -1))
    eb.set_network_input(x.reshape(B * G, 82))
    got = eb.read_output(eb.forward(eb.sigma_context(1.0))).reshape(B, G, 82)
    got_s = SamplerEngine(eb, sigmas).sample(x.reshape(B * G, 82), use_graph=True).cpu().numpy().reshape(B, G, 82)
    tol = 1e-6 if dtype == "f32" else 1e-6
    for b in range(B):
        assert np.abs(got[b] - single[b]).max() <= tol * np.abs(single[b]).max()
        assert np.abs(got_s[b] - single_s[b]).max() <= 10 * tol * np.abs(single_s[b]).max()


def test_device_stacker_pinned_inputs_and_leased_outputs(cuda_device):
    """Host edge of the public API: Datasets in page-locked memory are copied host -> device directly
    (same result as the packed path), and predictions are views of leased page-locked buffers that stay
    intact while the caller holds them and are reused once dropped."""
    import gc
    from gencast_flax_nnx_b200 import device_stacking, stacking
    case = make_case("tiny")
    st = device_stacking.DeviceStacker(cuda_device)
    sizes = dict(case.targets.sizes)
    a = st.to_nodes("inputs", case.inputs, sizes).clone()
    b = st.to_nodes("inputs", device_stacking.pin_dataset(case.inputs), sizes)
    assert torch.equal(a, b)
    ref, _ = stacking.dataset_to_nodes(case.inputs, sizes)
    np.testing.assert_array_equal(a.cpu().numpy(), ref)
    tgt_nodes, _ = stacking.dataset_to_nodes(case.targets, sizes)
    n1 = torch.from_numpy(tgt_nodes).to(cuda_device)
    d1 = st.from_nodes(n1, case.targets)
    held = {k: v.data for k, v in d1.items()}
    snap = {k: v.copy() for k, v in held.items()}
    outs = [st.from_nodes(n1 * float(i + 2), case.targets) for i in range(device_stacking.MAX_LEASED_OUTPUT_BUFFERS + 2)]
    for k in held:                                   # earlier results are untouched by later calls
        np.testing.assert_array_equal(held[k], snap[k])
        np.testing.assert_array_equal(outs[-1][k].data, snap[k] * float(len(outs) + 1))
        np.testing.assert_array_equal(snap[k], case.targets[k].data.astype(np.float32))
    leased_before = st._out_leased
    del outs, d1, held
    gc.collect()
    assert st._out_leased < leased_before            # dropped results hand their buffers back
    assert sum(len(v) for v in st._out_pool.values()) >= 1


def test_device_rollout_equals_host_rollout(cuda_device):
    """rollout.device_chunked_prediction (window kept on the GPU, gc_select_columns) yields, step for step,
    exactly what the reference-shaped host driver yields around GenCast.full_sampling."""
    from gencast_flax_nnx_b200 import configs, gencast, graph, rollout, synthetic
    from gencast_flax_nnx_b200.rngs import Rngs
    case = make_case("tiny")
    res, arch = configs.named_config("tiny")
    lat, lon = graph.regular_grid(res)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=2, seed=1, num_target_steps=3)
    sc = configs.SamplerConfig(num_noise_levels=3, stochastic_churn_rate=0.0)

    def model():
        return gencast.GenCast(configs.TASK, arch, sampler_config=sc, rngs=Rngs(7), params=case.params,
                               compute_dtype="bf16")
    host = rollout.chunked_prediction(
        lambda rng, inputs, targets_template, forcings: m1.full_sampling(inputs, targets_template, forcings),
        0, inputs, targets, forcings) if (m1 := model()) else None
    dev = rollout.device_chunked_prediction(model(), inputs, targets, forcings)
    assert dev.coords["time"].tolist() == host.coords["time"].tolist() == [12, 24, 36]
    for k in host.keys():
        assert dev[k].dims == host[k].dims
        np.testing.assert_array_equal(dev[k].data, host[k].data)
    assert not np.array_equal(host["2m_temperature"].data[:, 0], host["2m_temperature"].data[:, 1])


def test_benched_workload_1deg_four_members_equal_single_member_runs(cuda_device):
    """The default bench.py workload (GenCast 1 deg, 4 members evaluated together in bf16: CTA-pair GEMMs with
    staged gathers, TS-form attention over 4 x 81 query tiles, pipelined segment sums) gives every member what
    the one-member engine gives it, which test_full_size_1deg_properties ties to the fp32 path and, through
    it, to the oracle."""
    from gencast_flax_nnx_b200.engine import DenoiserEngine
    case = make_case("1deg")
    B, G = 4, case.graphs.num_grid_nodes
    rng = np.random.default_rng(11)
    x = rng.standard_normal((B, G, 82)).astype(np.float32)
    inp = np.stack([case.inp_nodes[:, 0] * (1 + 0.05 * b) for b in range(B)])
    frc = np.stack([case.frc_nodes[:, 0]] * B)
    e1 = DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype="bf16")
    single = []
    for b in (0, 3):
        e1.set_constant_features(inp[b], frc[b])
        e1.set_network_input(x[b])
        single.append(e1.read_output(e1.forward(e1.sigma_context(1.0))))
    del e1
    torch.cuda.empty_cache()
    eb = DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype="bf16", members=B)
    eb.set_constant_features(inp.reshape(B * G, -1), frc.reshape(B * G, -1))
    eb.set_network_input(x.reshape(B * G, 82))
    got = eb.read_output(eb.forward(eb.sigma_context(1.0))).reshape(B, G, 82)
    assert np.isfinite(got).all()
    for ref, b in zip(single, (0, 3)):
        err = np.abs(got[b] - ref).max() / np.abs(ref).max()
        print(f"1deg x 4 members, member {b}: max relative difference to the one-member engine {err:.2e}")
        assert err <= 1e-6
    assert not np.array_equal(got[0], got[3])


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("name,members", [("tiny", 1), ("tiny", 3), ("nano", 2)])
def test_c_forward_equals_python_sequencing(cuda_device, monkeypatch, name, members, dtype):
    """gc_denoiser_forward (the whole evaluation sequenced in C++, one call through the C ABI) gives bitwise the result
    of the same launches issued one by one from Python, eagerly and inside the sampler's CUDA graph (parallel branch)."""
    from gencast_flax_nnx_b200.engine import DenoiserEngine, SamplerEngine, noise_schedule
    case = make_case(name)
    B, G = members, case.graphs.num_grid_nodes
    rng = np.random.default_rng(12)
    x = rng.standard_normal((B * G, 82)).astype(np.float32)
    inp = np.concatenate([case.inp_nodes[:, 0] * (1 + 0.1 * b) for b in range(B)])
    frc = np.concatenate([case.frc_nodes[:, 0]] * B)
    sigmas = noise_schedule(80.0, 0.03, 3, 7.0)
    outs = {}
    for impl in ("py", "c"):
        monkeypatch.setenv("GENCAST_FORWARD", impl)
        eng = DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype=dtype, members=B)
        assert eng.forward_impl == impl
        eng.set_constant_features(inp, frc)
        eng.set_network_input(x)
        one = eng.read_output(eng.forward(eng.sigma_context(1.0)))
        se = SamplerEngine(eng, sigmas)
        outs[impl] = (one, se.sample(x, use_graph=True).cpu().numpy().copy(), se.sample(x, use_graph=False).cpu().numpy().copy())
    for a, b in zip(outs["py"], outs["c"]):
        assert np.isfinite(a).all() and np.array_equal(a, b)
    assert np.array_equal(outs["c"][1], outs["c"][2])


@pytest.mark.parametrize("name,members", [("tiny", 2), ("nano", 3)])
@pytest.mark.parametrize("impl", ["py", "c"])
def test_fused_edge_kernels_equal_unfused_paths(cuda_device, monkeypatch, name, members, impl):
    """bf16 evaluation with the fused edge kernels (gc_edge_mlp_rows for grid2mesh with the receivers' part folded into
    the per-level table, gc_edge_mlp_sum3 for mesh2grid) against GENCAST_EDGE_FUSED=0 (gc_edge_hidden -> gc_gemm ->
    gc_ln_cond_segment_sum): the same network up to bf16 rounding of intermediates, through both sequencers."""
    from gencast_flax_nnx_b200.engine import DenoiserEngine
    case = make_case(name)
    B, G = members, case.graphs.num_grid_nodes
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B * G, 82)).astype(np.float32)
    inp = np.concatenate([case.inp_nodes[:, 0] * (1 + 0.1 * b) for b in range(B)])
    frc = np.concatenate([case.frc_nodes[:, 0]] * B)
    monkeypatch.setenv("GENCAST_FORWARD", impl)
    outs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("GENCAST_EDGE_FUSED", fused)
        eng = DenoiserEngine(case.graphs, case.arch, case.params, case.layout, compute_dtype="bf16", members=B)
        assert eng.fuse_m2g == (fused == "1") and eng.fuse_g2m == (fused == "1")
        eng.set_constant_features(inp, frc)
        eng.set_network_input(x)
        outs[fused] = eng.read_output(eng.forward(eng.sigma_context(0.7, pin=True)))
    a, b = outs["1"], outs["0"]
    assert np.isfinite(a).all() and np.isfinite(b).all()
    scale = np.abs(b).max(axis=0) + 1e-6
    assert (np.abs(a - b).max(axis=0) / scale).max() < 1.5e-2       # bf16 noise of two roundings of the edge latents


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_sampler_with_stochastic_churn_matches_oracle(cuda_device, dtype):
    """DPM-Solver++ 2S with stochastic churn (gencast/samplers_utils.py:414-452, dpm...2s.py:127-137): churned steps
    move the state to sigma (1 + rate) with fresh noise before the 2S update; same noises on both sides."""
    from gencast_flax_nnx_b200.engine import SamplerEngine, noise_schedule, stochastic_churn_rate_schedule
    from oracle import gencast_oracle as o
    case = make_case("tiny")
    eng = _engine(case, dtype)
    sigmas = noise_schedule(80.0, 0.03, 4, 7.0)
    rates = stochastic_churn_rate_schedule(sigmas, 2.5, 0.75, float("inf"))
    np.testing.assert_allclose(rates, o.stochastic_churn_rate_schedule(sigmas, 2.5, 0.75, float("inf")))
    assert 0 < (rates > 0).sum() < len(rates)                       # some steps churn, the low-noise ones do not
    eng.set_constant_features(case.inp_nodes[:, 0], case.frc_nodes[:, 0])
    rng = np.random.default_rng(13)
    noise = rng.standard_normal((eng.G, eng.n_out)).astype(np.float32)
    se = SamplerEngine(eng, sigmas, churn_rates=rates, noise_level_inflation_factor=1.05)
    cn = rng.standard_normal((se.num_churn_steps, eng.G, eng.n_out)).astype(np.float32)
    got = se.sample(noise, use_graph=True, churn_noise=torch.from_numpy(cn)).cpu().numpy().copy()
    eager = se.sample(noise, use_graph=False, churn_noise=torch.from_numpy(cn)).cpu().numpy().copy()
    assert np.array_equal(got, eager)
    with pytest.raises(ValueError):
        se.sample(noise)
    dt = torch.float64
    init = case.split_targets(torch.as_tensor(noise[:, None, :] * sigmas[0]).to(dt))
    frc = {k: torch.as_tensor(v).to(dt) for k, v in case.frc_vars.items()}
    cnoise = [case.split_targets(torch.as_tensor(c[:, None, :]).to(dt)) for c in cn]
    ref = o.dpm_solver_2s(case.params, case.oracle_graph, case.oracle_arch, torch.as_tensor(case.inp_nodes).to(dt), frc, init,
                          sigmas, dt, churn_rates=rates, inflation=1.05, churn_noise=cnoise)
    ref = torch.cat([ref[n] for n, _ in case.target_vars], dim=-1)[:, 0].numpy()
    errs = per_variable_error(case, got, ref)
    print(f"sampler with churn {dtype}: worst per-variable error {max(errs.values()):.3e}")
    assert max(errs.values()) <= (2e-3 if dtype == "f32" else 5e-2), errs
    # and it differs from the deterministic sampler
    det = SamplerEngine(eng, sigmas).sample(noise, use_graph=False).cpu().numpy()
    assert np.abs(det - got).max() > 1e-3


@pytest.mark.parametrize("order", ["nan_outside", "norm_outside"])
def test_device_rollout_with_normalization_and_nan_cleaning_equals_host_wrappers(cuda_device, order):
    """Autoregressive rollout with the window on the GPU around InputsAndResiduals + NaNCleaner (gc_normalize_cast,
    gc_unnormalize_residual) yields, step for step and bit for bit, what the reference-shaped host wrappers
    (common/normalization.py:200-238, gencast/nan_cleaning.py:129-156) yield around GenCast.full_sampling."""
    from gencast_flax_nnx_b200 import configs, gencast, graph, nan_cleaning, normalization, rollout, synthetic
    from gencast_flax_nnx_b200.rngs import Rngs
    from gencast_flax_nnx_b200.xarray_lite import DataArray, Dataset
    case = make_case("tiny")
    res, arch = configs.named_config("tiny")
    lat, lon = graph.regular_grid(res)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=2, seed=2, num_target_steps=3)
    rng = np.random.default_rng(4)
    var = "2m_temperature"            # stands for the reference's sea-surface temperature (NaN over land)
    # physical-looking inputs: offset and scale per variable; NaNs over "land" in the cleaned variable (all frames)
    lev = len(configs.TASK.pressure_levels)
    stat = lambda lo, hi, v: (DataArray(np.linspace(lo, hi, lev).astype(np.float32), ("level",)) if "level" in inputs[v].dims
                              else DataArray(np.float32(0.5 * (lo + hi)), ()))
    std = Dataset({v: stat(2.0, 5.0, v) for v in inputs.keys()})
    mean = Dataset({v: stat(-1.0, 3.0, v) for v in inputs.keys()})
    dstd = Dataset({v: DataArray(np.float32(0.5 + 0.1 * i), ()) for i, v in enumerate(sorted(targets.keys()))})
    phys = {}
    for v, a in inputs.items():
        x = a.data * normalization._stat_like(a, std[v]) + normalization._stat_like(a, mean[v])
        if v == var:
            x = x.copy()
            x[..., 3:6, 5:11] = np.nan
        phys[v] = DataArray(x.astype(np.float32), a.dims)
    inputs = Dataset(phys, inputs.coords)
    sc = configs.SamplerConfig(num_noise_levels=3, stochastic_churn_rate=0.0)
    fill = Dataset({var: DataArray(np.float32(271.0), ())})

    def wrapped():
        m = gencast.GenCast(configs.TASK, arch, sampler_config=sc, rngs=Rngs(7), params=case.params, compute_dtype="bf16")
        if order == "nan_outside":
            w = normalization.InputsAndResiduals(m, std, mean, dstd)
            return nan_cleaning.NaNCleaner(w, var, fill, reintroduce_nans=True)
        w = nan_cleaning.NaNCleaner(m, var, Dataset({var: DataArray(np.float32(0.25), ())}), reintroduce_nans=True)
        return normalization.InputsAndResiduals(w, std, mean, dstd)

    m1 = wrapped()
    host = rollout.chunked_prediction(lambda rng, inputs, targets_template, forcings: m1.full_sampling(inputs, targets_template, forcings),
                                      0, inputs, targets, forcings)
    dev = rollout.device_chunked_prediction(wrapped(), inputs, targets, forcings)
    for k in host.keys():
        assert dev[k].dims == host[k].dims
        np.testing.assert_array_equal(dev[k].data, host[k].data)
    assert np.isnan(host[var].data[..., 3:6, 5:11]).all() and np.isfinite(host[var].data[..., 0, :]).all()
    assert np.isfinite(host["mean_sea_level_pressure"].data).all()


def test_denoiser_call_with_a_different_noise_level_per_batch_element(cuda_device):
    """Denoiser.__call__ conditions every batch element on its own noise level (gencast/denoiser.py:190-198): with
    levels (1.0, 0.1) element b equals the evaluation of the whole batch at level b's value."""
    from gencast_flax_nnx_b200 import configs, gencast, graph, stacking, synthetic
    from gencast_flax_nnx_b200.rngs import Rngs
    from gencast_flax_nnx_b200.xarray_lite import DataArray
    case = make_case("tiny")
    res, arch = configs.named_config("tiny")
    lat, lon = graph.regular_grid(res)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=2, seed=6)
    model = gencast.GenCast(configs.TASK, arch, rngs=Rngs(0), params=case.params, compute_dtype="f32")
    noisy = stacking.nodes_to_dataset(np.random.default_rng(1).standard_normal((len(lat) * len(lon), 2, 82)).astype(np.float32), targets)
    lv = lambda a, b: DataArray(np.array([a, b], np.float32), ("batch",))
    mixed = model.denoiser(inputs, noisy, lv(1.0, 0.1), forcings)
    hi, lo = model.denoiser(inputs, noisy, lv(1.0, 1.0), forcings), model.denoiser(inputs, noisy, lv(0.1, 0.1), forcings)
    for k in mixed.keys():
        ax = mixed[k].dims.index("batch")
        np.testing.assert_array_equal(np.take(mixed[k].data, 0, ax), np.take(hi[k].data, 0, ax))
        np.testing.assert_array_equal(np.take(mixed[k].data, 1, ax), np.take(lo[k].data, 1, ax))
        assert not np.array_equal(np.take(hi[k].data, 1, ax), np.take(lo[k].data, 1, ax))
