"""CPU tests of the host side: static graph tables against the reference's own geometry code
(tests/golden/graph_tiny.npz, tools/make_graph_golden.py) and its test invariants
(common/icosahedral_mesh_test.py, common/grid_mesh_connectivity_test.py), layout glue, schedules,
the rollout driver, and that the C-ABI library exports every symbol the header declares."""
import os
import re

import numpy as np
import pytest

from gencast_flax_nnx_b200 import configs, graph, stacking, synthetic
from gencast_flax_nnx_b200.xarray_lite import DataArray, Dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ static graphs
def test_graph_tables_match_reference_geometry_code_bit_for_bit():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "graph_tiny.npz"))
    res, arch = configs.named_config("tiny")
    lat, lon = graph.regular_grid(res)
    g = graph.build_denoiser_graphs(lat, lon, arch.mesh_size, arch.sparse_transformer_config.attention_k_hop)
    np.testing.assert_array_equal(g.mesh.vertices, gold["vertices"])
    np.testing.assert_array_equal(g.mesh.faces, gold["faces"])
    assert g.query_radius == pytest.approx(float(gold["radius"]), rel=0, abs=0)
    np.testing.assert_array_equal(g.g2m_senders, gold["g2m_senders"])
    np.testing.assert_array_equal(g.g2m_receivers, gold["g2m_receivers"])
    for k in ("g2m_grid_feat", "g2m_mesh_feat", "g2m_edge_feat"):
        np.testing.assert_array_equal(getattr(g, k), gold[k].astype(np.float32))


@pytest.mark.parametrize("splits", [0, 1, 2, 3, 4])
def test_icosphere_invariants(splits):
    """common/icosahedral_mesh_test.py:23-32, :95-127."""
    m = graph.icosphere(splits)
    assert m.vertices.shape == (10 * 4 ** splits + 2, 3) and m.faces.shape == (20 * 4 ** splits, 3)
    np.testing.assert_allclose(np.linalg.norm(m.vertices, axis=-1), 1.0, rtol=1e-6)
    v, f = m.vertices.astype(np.float64), m.faces
    normal = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 1]])
    normal /= np.linalg.norm(normal, axis=-1, keepdims=True)
    centre = v[f].mean(1)
    centre /= np.linalg.norm(centre, axis=-1, keepdims=True)
    np.testing.assert_allclose(np.einsum("ik,ik->i", normal, centre), 1.0, atol=6e-4)
    if splits:
        np.testing.assert_array_equal(graph.icosphere(splits - 1).vertices, m.vertices[:10 * 4 ** (splits - 1) + 2])


def test_faces_to_edges_order():
    """common/icosahedral_mesh_test.py:73-92."""
    s, r = graph.faces_to_edges(np.array([[0, 1, 2], [3, 4, 5]]))
    np.testing.assert_array_equal(s, [0, 3, 1, 4, 2, 5])
    np.testing.assert_array_equal(r, [1, 4, 2, 5, 0, 3])


def test_grid_positions_golden():
    """common/grid_mesh_connectivity_test.py:24-48."""
    q = 1 / np.sqrt(2)
    expected = np.array([[[q, 0, -q], [0, q, -q], [-q, 0, -q], [0, -q, -q]],
                         [[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0]],
                         [[q, 0, q], [0, q, q], [-q, 0, q], [0, -q, q]]])
    got = graph.grid_positions(np.array([-45.0, 0.0, 45.0]), np.array([0.0, 90.0, 180.0, 270.0]))
    np.testing.assert_allclose(got.reshape(3, 4, 3), expected, atol=1e-15)


def test_nano_structural_fingerprints():
    """SURVEY.md Appendix A (measured with the reference's mesh code): edge counts, k-hop size, band width."""
    res, arch = configs.named_config("nano")
    lat, lon = graph.regular_grid(res)
    g = graph.build_denoiser_graphs(lat, lon, arch.mesh_size, 8)
    assert (g.num_grid_nodes, g.num_mesh_nodes) == (10512, 2562)
    assert len(g.g2m_senders) == 16830 and len(g.m2g_senders) == 3 * 10512
    assert g.khop.nnz == 542922 and graph.mask_block_size(g.khop) == 649
    indeg = np.bincount(g.g2m_receivers, minlength=2562)
    assert indeg.min() == 3 and indeg.max() == 218
    assert np.bincount(g.g2m_senders, minlength=10512).min() >= 1


def test_mesh2grid_edges_are_the_containing_triangle():
    """common/grid_mesh_connectivity.py:89-133: three edges per grid point, grid-major, from the vertices
    of the mesh face that contains it (trimesh is unavailable, so this is checked geometrically)."""
    res, arch = configs.named_config("tiny")
    lat, lon = graph.regular_grid(res)
    g = graph.build_denoiser_graphs(lat, lon, arch.mesh_size, 2)
    G = g.num_grid_nodes
    np.testing.assert_array_equal(g.m2g_receivers, np.repeat(np.arange(G), 3))
    faces = {tuple(sorted(f)) for f in g.mesh.faces.tolist()}
    tri = g.m2g_senders.reshape(G, 3)
    assert all(tuple(sorted(t)) in faces for t in tri.tolist())
    p = graph.grid_positions(lat, lon).astype(np.float64)
    a, b, c = (g.mesh.vertices[tri[:, i]].astype(np.float64) for i in range(3))
    # the point's radial projection falls inside the triangle: barycentric coordinates of the
    # intersection of the ray with the triangle's plane are all >= 0 (up to rounding)
    n = np.cross(b - a, c - a)
    t = np.einsum("ij,ij->i", a, n) / np.einsum("ij,ij->i", p, n)
    x = p * t[:, None]
    area = np.einsum("ij,ij->i", n, n)
    w0 = np.einsum("ij,ij->i", np.cross(b - x, c - x), n) / area
    w1 = np.einsum("ij,ij->i", np.cross(c - x, a - x), n) / area
    w2 = 1 - w0 - w1
    assert min(w0.min(), w1.min(), w2.min()) > -1e-6


def test_csr_tiles_and_patch_order():
    res, arch = configs.named_config("nano")
    mesh = graph.permute_mesh_to_banded(graph.icosphere(arch.mesh_size))
    kh = graph.khop_neighbourhoods(mesh, 8)
    order = graph.patch_order(mesh.vertices, 128)
    assert sorted(order.tolist()) == list(range(mesh.vertices.shape[0]))
    kp = kh.tocsr()[order][:, order].tocsr()
    tp, tk, tm = graph.khop_tiles(kp, 128)
    assert tp[-1] == len(tk) == tm.shape[0] and tm.shape[1:] == (128, 4)
    dense = kp.toarray().astype(bool)
    total = 0
    for qt in range(len(tp) - 1):
        for t in range(tp[qt], tp[qt + 1]):
            bits = ((tm[t][:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(128, 128).astype(bool)
            blk = np.zeros((128, 128), bool)
            sub = dense[qt * 128:(qt + 1) * 128, tk[t] * 128:(tk[t] + 1) * 128]
            blk[:sub.shape[0], :sub.shape[1]] = sub
            assert (bits == blk).all()
            total += int(bits.sum())
    assert total == kh.nnz
    recv = np.array([2, 0, 2, 1, 0, 2])
    rp, perm = graph.csr_by_receiver(recv, 4)
    np.testing.assert_array_equal(rp, [0, 2, 3, 6, 6])
    np.testing.assert_array_equal(perm, [1, 4, 3, 0, 2, 5])


# ------------------------------------------------------------------ layout glue
def test_stacking_round_trip_and_channel_order():
    lat, lon = graph.regular_grid(30.0)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=2, seed=3)
    sizes = dict(targets.sizes)
    nodes, layout = stacking.dataset_to_nodes(inputs, sizes)
    assert nodes.shape == (len(lat) * len(lon), 2, sum(c for _, c in layout))
    assert [n for n, _ in layout] == sorted(inputs.keys())                    # common/model_utils.py:649-652
    # a (time, level) variable contributes time-major channels (common/model_utils.py:617-623)
    off = dict(zip([n for n, _ in layout], np.cumsum([0] + [c for _, c in layout])[:-1]))
    t = inputs["temperature"].data                                            # batch, time, level, lat, lon
    np.testing.assert_array_equal(nodes[5 * len(lon) + 7, 1, off["temperature"] + 1 * 13 + 4], t[1, 1, 4, 5, 7])
    rng = np.random.default_rng(0)
    x = rng.standard_normal((len(lat) * len(lon), 2, 82)).astype(np.float32)
    ds = stacking.nodes_to_dataset(x, targets)
    back, tl = stacking.dataset_to_nodes(ds, sizes)
    np.testing.assert_array_equal(back, x)
    assert [n for n, _ in tl] == sorted(targets.keys())
    with pytest.raises(ValueError):
        stacking.nodes_to_dataset(x[:, :, :81], targets)


def test_channel_layout_rows_cover_the_first_layer_kernel():
    from gencast_flax_nnx_b200.engine import ChannelLayout
    lat, lon = graph.regular_grid(30.0)
    inputs, targets, forcings = synthetic.make_example(lat, lon)
    sizes = dict(targets.sizes)
    inp, _ = stacking.dataset_to_nodes(inputs, sizes)
    lay = ChannelLayout(inp.shape[-1], tuple(stacking.channel_layout(forcings)), tuple(stacking.channel_layout(targets)))
    tgt, const = lay.reference_rows()
    assert sorted(np.concatenate([tgt, const]).tolist()) == list(range(3 + lay.num_data_channels))
    # reference order of the second block: sorted(forcings U targets) (gencast/denoiser.py:184, model_utils.py:649)
    merged = forcings.assign(targets)
    _, ml = stacking.dataset_to_nodes(merged, sizes)
    start = 3 + inp.shape[-1]
    pos = {}
    for n, c in ml:
        pos[n] = start
        start += c
    assert tgt[0] == pos[sorted(targets.keys())[0]]
    assert const[-1] == pos["year_progress_sin"]


# ------------------------------------------------------------------ schedules, sampler plan, rollout
def test_noise_schedule_and_sampler_plan():
    from gencast_flax_nnx_b200.engine import noise_schedule
    from oracle import gencast_oracle as o
    s = noise_schedule(80.0, 0.03, 20, 7.0)
    np.testing.assert_allclose(s, o.noise_schedule(80.0, 0.03, 20, 7.0), rtol=0, atol=0)
    assert len(s) == 21 and s[0] == pytest.approx(80.0) and s[-2] == pytest.approx(0.03) and s[-1] == 0.0
    assert np.all(np.diff(s) < 0)
    c = np.linspace(1, 0, 20)
    np.testing.assert_allclose(s[:-1], (0.03 ** (1 / 7) + c * (80 ** (1 / 7) - 0.03 ** (1 / 7))) ** 7)


def test_rollout_driver_matches_reference_semantics():
    """common/rollout.py:245-401 with a fake predictor: window roll, time relabelling, chunk validation."""
    from gencast_flax_nnx_b200 import rollout
    lat, lon = np.array([0.0, 10.0]), np.array([0.0, 10.0, 20.0])
    coords = dict(lat=lat, lon=lon, batch=np.arange(1))

    def field(values):
        v = np.asarray(values, np.float32)
        return DataArray(np.broadcast_to(v[None, :, None, None], (1, len(v), 2, 3)).copy(), ("batch", "time", "lat", "lon"))

    inputs = Dataset({"x": field([1, 2]), "f": field([10, 20]), "static": DataArray(np.ones((2, 3), np.float32), ("lat", "lon"))},
                     dict(coords, time=np.array([-12, 0])))
    targets = Dataset({"x": field([0, 0, 0, 0])}, dict(coords, time=np.array([12, 24, 36, 48])))
    forcings = Dataset({"f": field([30, 40, 50, 60])}, dict(coords, time=np.array([12, 24, 36, 48])))
    seen = []

    def predictor(rng, inputs, targets_template, forcings):
        seen.append((rng, inputs["x"].data[0, :, 0, 0].tolist(), inputs["f"].data[0, :, 0, 0].tolist(),
                     targets_template.coords["time"].tolist(), forcings["f"].data[0, :, 0, 0].tolist()))
        nxt = inputs["x"].data[:, -1:] + inputs["x"].data[:, -2:-1]            # Fibonacci step
        return Dataset({"x": DataArray(nxt, ("batch", "time", "lat", "lon"))}, targets_template.coords)

    out = rollout.chunked_prediction(predictor, 0, inputs, targets, forcings)
    assert out["x"].data[0, :, 0, 0].tolist() == [3, 5, 8, 13]
    assert out.coords["time"].tolist() == [12, 24, 36, 48]
    assert [s[1] for s in seen] == [[1, 2], [2, 3], [3, 5], [5, 8]]
    assert [s[2] for s in seen] == [[10, 20], [20, 30], [30, 40], [40, 50]]
    assert all(s[3] == [12] for s in seen)                       # chunk times relabelled to the first chunk's
    assert [s[4] for s in seen] == [[30], [40], [50], [60]]
    assert len({s[0] for s in seen}) == 4                        # a fresh key per chunk
    with pytest.raises(ValueError):
        rollout.chunked_prediction(predictor, 0, inputs, targets, forcings, num_steps_per_chunk=3)
    bad = Dataset({"x": field([0, 0, 0])}, dict(coords, time=np.array([12, 24, 48])))
    with pytest.raises(ValueError):
        rollout.chunked_prediction(predictor, 0, inputs, bad, forcings.isel(time=slice(0, 3)))


def test_api_errors_without_gpu():
    from gencast_flax_nnx_b200 import dpm_solver_plus_plus_2s as dpm
    churned = dpm.Sampler(None, 80.0, 0.03, 20, 7.0, 2.5, 0.75, float("inf"), 1.05)
    assert churned._stochastic_churn and (churned._per_step_churn_rates > 0).sum() == 14      # levels >= 0.75 churn
    assert np.isclose(churned._per_step_churn_rates.max(), 2.5 / 20)
    s = dpm.Sampler(None, 80.0, 0.03, 20, 7.0, 0.0, 0.75, float("inf"), 1.05)
    assert not s._stochastic_churn and not s._per_step_churn_rates.any()
    with pytest.raises(ValueError):
        s(None, None, None, rngs=None)
    assert configs.num_outputs(configs.TASK) == 82


# ------------------------------------------------------------------ the C ABI
def test_library_exports_every_declared_symbol():
    import ctypes
    from gencast_flax_nnx_b200 import _lib
    header = open(os.path.join(ROOT, "include", "gencast_b200.h")).read()
    declared = set(re.findall(r"GC_API\s+[\w\s\*]+?\b(gc_\w+)\s*\(", header))
    assert {"gc_gemm", "gc_khop_attention_tiles", "gc_ln_cond_segment_sum", "gc_dpm_update"} <= declared
    assert declared == set(_lib.SIGNATURES)                       # the ctypes table mirrors the header
    if not _lib.LIB_PATH.exists():
        pytest.skip("library not built yet (run __graft_entry__.build())")
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.gc_abi_version() >= 1
    assert loaded.gc_sizeof_gemm_args() == ctypes.sizeof(_lib.GemmArgs)      # struct layout agrees with the C side


def test_xla_ffi_shim_type_checks_against_the_c_abi():
    """The jax.ffi handlers (csrc/xla_ffi_shim.cc) cannot be compiled against jaxlib's headers here; a stand-in for the part
    of the XLA FFI C++ API they use (tools/ffi_stub) lets g++ type-check every handler body, i.e. every call into
    include/gencast_b200.h, so the shim cannot drift from the library's signatures unnoticed."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("needs g++ and the CUDA headers")
    shim = os.path.join(ROOT, "gencast_flax_nnx_b200", "csrc", "xla_ffi_shim.cc")
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tools", "ffi_stub"), "-I", "/usr/local/cuda/include", shim]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pre = subprocess.run(cmd[:2] + ["-E"] + cmd[4:], capture_output=True, text=True).stdout
    for name in ("gc_gemm(", "gc_edge_mlp_sum3(", "gc_edge_mlp_rows(", "gc_denoiser_forward(", "gc_khop_attention_gather("):
        assert name in pre, name                     # the handler bodies were really compiled (the stub header was found)


def test_param_interchange_round_trip(tmp_path):
    from gencast_flax_nnx_b200 import params
    res, arch = configs.named_config("tiny")
    shapes = params.param_shapes(arch, 260, 82)
    p = params.init_perturbed(shapes, seed=3)
    path = str(tmp_path / "w.npz")
    params.save_npz(p, path)
    q = params.load_npz(path)
    assert set(q) == set(p) and all(np.array_equal(p[k], q[k]) for k in p)
    params.check_complete(q, shapes)
    del q[next(iter(q))]
    with pytest.raises(ValueError):
        params.check_complete(q, shapes)

    class V:                                     # stands for an nnx.Param
        def __init__(self, v): self.value = v
    flat = params.from_nnx_state([(("denoiser", "blocks", 3, "kernel"), V(np.ones((2, 2))))])
    assert list(flat) == ["denoiser/blocks/3/kernel"]
    ref = params.init_reference_like(shapes)
    # at the reference initialisation every transformer block is the identity
    assert not ref["denoiser/predictor/mesh_gnn/batch_first_transformer/blocks/0/attn_module/final_linear/kernel"].any()


def test_window_update_table_matches_host_window_roll():
    """The column table of the on-device rollout (gc_select_columns) reproduces _get_next_inputs
    (common/rollout.py:379-401) on the stacked layout, for the full GenCast task variable set."""
    from gencast_flax_nnx_b200 import graph, rollout, stacking, synthetic
    from gencast_flax_nnx_b200.xarray_lite import merge
    lat, lon = graph.regular_grid(30.0)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=2, seed=3)
    sizes = dict(targets.sizes)
    rng = np.random.default_rng(0)
    pred = targets.map(lambda v: DataArray(rng.standard_normal(v.shape).astype(np.float32), v.dims))
    table = rollout.window_update_table(inputs, pred, forcings)
    srcs = [stacking.dataset_to_nodes(d, sizes)[0] for d in (inputs, pred, forcings)]
    new = np.stack([srcs[t >> 24][:, :, t & 0xffffff] for t in table], axis=-1)
    ref, _ = stacking.dataset_to_nodes(rollout._get_next_inputs(inputs, merge([pred, forcings])), sizes)
    np.testing.assert_array_equal(new, ref)
    assert (table >> 24).max() == 2 and ((table >> 24) == 1).sum() == 82
    # an input with a time axis that is neither predicted nor forced is rejected, as in the reference
    with pytest.raises(ValueError):
        rollout.window_update_table(inputs, pred.drop_vars(["2m_temperature"]), forcings)


def test_khop_compact_steps_reconstruct_the_pattern():
    """graph.khop_compact_steps: keys of a query tile are the sorted union of its rows' neighbours, padded columns
    carry no mask bit, and (keys, mask) reproduce the pattern exactly."""
    from scipy import sparse
    from gencast_flax_nnx_b200 import graph
    rng = np.random.default_rng(3)
    n = 300
    dense = rng.random((n, n)) < 0.07
    dense[np.arange(n), np.arange(n)] = True
    dense[:, 40:90] = False
    dense[np.arange(40, 90), np.arange(40, 90)] = True
    sp, keys, cm, work = graph.khop_compact_steps(sparse.csr_matrix(dense), 128, 64)
    nq = -(-n // 128)
    assert len(sp) == nq + 1 and len(keys) == sp[-1] * 64 and cm.shape == (sp[-1], 128, 2)
    assert sorted(work.tolist()) == list(range(nq)) and np.all(np.diff(np.diff(sp)[work]) <= 0)
    rebuilt = np.zeros_like(dense)
    bits = np.unpackbits(cm.view(np.uint8), axis=-1, bitorder="little").reshape(sp[-1], 128, 64).astype(bool)
    for t in range(nq):
        union = np.unique(np.nonzero(dense[t * 128:(t + 1) * 128])[1])
        k = keys[sp[t] * 64: sp[t + 1] * 64]
        assert sp[t + 1] - sp[t] == -(-len(union) // 64)
        np.testing.assert_array_equal(k[:len(union)], union)
        assert np.all(k[len(union):] == union[-1])
        b = bits[sp[t]:sp[t + 1]].transpose(1, 0, 2).reshape(128, -1)       # [row, compacted column]
        assert not b[:, len(union):].any()
        rows = min(128, n - t * 128)
        assert not b[rows:].any()
        for r in range(rows):
            rebuilt[t * 128 + r, k[b[r]]] = True
    np.testing.assert_array_equal(rebuilt, dense)


def test_multiple_runs_fan_out_and_packed_writer(tmp_path):
    """chunked_prediction_generator_multiple_runs (common/rollout.py:78-202): per-sample trajectories labelled with
    `sample`, per-sample initial conditions, device groups -> member assignment; PackedRolloutWriter / Reader round trip."""
    from gencast_flax_nnx_b200 import rollout
    from gencast_flax_nnx_b200.ensemble_writer import PackedRolloutReader, PackedRolloutWriter
    lat, lon = graph.regular_grid(30.0)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=1, seed=3, num_target_steps=2)
    S = 4
    # one initial condition per sample along a leading 'sample' dim
    inp_s = Dataset({k: DataArray(np.stack([v.data * (1 + s) for s in range(S)]), ("sample",) + tuple(v.dims))
                     for k, v in inputs.items()}, inputs.coords)

    def predictor(rng, inputs, targets_template, forcings):
        # prediction = last input frame + rng-dependent offset
        return Dataset({k: DataArray(inputs[k].isel(time=slice(-1, None)).transpose(*v.dims).data + np.float32(rng % 7), v.dims)
                        for k, v in targets_template.items()}, targets_template.coords)

    chunks = list(rollout.chunked_prediction_generator_multiple_runs(predictor, list(range(10, 10 + S)), inp_s, targets, forcings, S))
    assert len(chunks) == S * 2 and [int(c.coords["sample"]) for c in chunks] == [0, 0, 1, 1, 2, 2, 3, 3]
    assert [c.coords["time"].tolist() for c in chunks[:2]] == [[12], [24]]
    a, b = chunks[0]["2m_temperature"].data, chunks[2]["2m_temperature"].data
    assert not np.array_equal(a, b)
    with pytest.raises(AssertionError):
        list(rollout.chunked_prediction_generator_multiple_runs(predictor, list(range(3)), inputs, targets, forcings, 3, pmap_devices=[0, 1]))
    only = list(rollout.chunked_prediction_generator_multiple_runs(predictor, list(range(S)), inputs, targets, forcings, S, pmap_devices=[0, 1]))
    assert sorted({int(c.coords["sample"]) for c in only}) == [0, 1]            # rank 0 of 2 (no process group): first half
    # packed writer: [time, member, G, C]
    tmpl = targets.isel(time=slice(0, 1))
    path = str(tmp_path / "rollout.bin")
    w = PackedRolloutWriter(path, tmpl, [12, 24], num_members=S)
    sizes = dict(tmpl.sizes)
    for c in chunks:
        step = [12, 24].index(int(c.coords["time"][0]))
        nodes, _ = stacking.dataset_to_nodes(Dataset(c.data_vars, {k: v for k, v in c.coords.items() if k != "sample"}), sizes)
        w.write_step(step, nodes[:, 0], first_member=int(c.coords["sample"]))
    w.close()
    r = PackedRolloutReader(path)
    assert r.shape == (2, S, len(lat) * len(lon), 82)
    back = r.step(1, slice(2, 3))
    want = [c for c in chunks if int(c.coords["sample"]) == 2 and int(c.coords["time"][0]) == 24][0]
    for k in want.keys():
        np.testing.assert_array_equal(back[k].data, want[k].data)
