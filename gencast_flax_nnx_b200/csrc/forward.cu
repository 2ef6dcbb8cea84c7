// gc_denoiser_forward: one whole network evaluation F(c_in x, sigma) sequenced in C++ over the kernels of this
// library, so that a host binding (XLA FFI custom call, ctypes, C) makes ONE call per evaluation instead of ~135.
//
// Reference call tree this replaces: DenoiserArchitecture.__call__ (gencast/denoiser.py:303-341) ->
// _run_grid2mesh_gnn (:602-688), _run_mesh_gnn (:691-728), _run_mesh2grid_gnn (:730-768), i.e.
// DeepTypedGraphNet.__call__ (common/deep_typed_graph_net.py:493-581), MeshTransformer.__call__
// (gencast/transformer.py:94-121) -> Transformer / Block (gencast/sparse_transformer.py:486-525, :624-634).
// The launch sequence is the one gencast_flax_nnx_b200/engine.py documents (and can still issue itself from
// Python for per-kernel timing); both produce bitwise identical results (tests/test_denoiser_gpu.py).
//
// Enqueue-only: no allocation, no synchronisation.  With ws->branch_stream set, the grid-node update of the encoder
// and the decoder's receiver-side partial product (which nothing on the mesh side needs) are enqueued on that stream
// between ws->fork_event and ws->join_event, so that inside a captured CUDA graph they form a parallel branch.
#include "common.cuh"

namespace gc {
namespace {

struct Seg { const void* a; int64_t lda; const void* w; int k; };

struct Ctx {
  cudaStream_t st;
  int dtype;        // operand dtype of every GEMM
  int L;
  bool fuse_ln;     // second MLP layer + LayerNorm + affine + residual through gc_linear_ln_cond
};

int run_gemm(const Ctx& c, cudaStream_t st, const Seg* segs, int nseg, int64_t m, int n, void* out, int out_dtype,
             const float* bias, int act, const void* residual, int res_dtype, const void* g0, const int32_t* i0,
             const void* g1, const int32_t* i1, bool static_w) {
  gc_gemm_args a = {};
  for (int s = 0; s < nseg; ++s) {
    a.a[s] = segs[s].a; a.w[s] = segs[s].w; a.lda[s] = segs[s].lda; a.ldw[s] = segs[s].k; a.k[s] = segs[s].k;
  }
  a.num_segments = nseg; a.m = m; a.n = n; a.dtype = c.dtype;
  a.bias = bias; a.act = act;
  if (g0 != nullptr) { a.gather_src[0] = g0; a.gather_idx[0] = i0; a.ld_gather[0] = c.L; a.gather_dtype = c.dtype; }
  if (g1 != nullptr) { a.gather_src[1] = g1; a.gather_idx[1] = i1; a.ld_gather[1] = c.L; a.gather_dtype = c.dtype; }
  if (residual != nullptr) { a.residual = residual; a.ld_res = n; a.res_dtype = res_dtype; }
  a.out = out; a.ldo = n; a.out_dtype = out_dtype;
  a.flags = static_w ? GC_GEMM_STATIC_WEIGHTS : 0;
  return gc_gemm(st, &a);
}

#define GC_TRY(expr)            \
  do {                          \
    const int _rc = (expr);     \
    if (_rc != GC_OK) return _rc; \
  } while (0)

// MLP (+ LayerNorm + conditional affine [+ residual]): Linear -> swish -> Linear -> gc_ln_cond
int mlp_ln(const Ctx& c, cudaStream_t st, const gc_mlp2& w, const void* const* a, int64_t rows, void* h, void* y, void* out,
           int out_dtype, const float* so, const void* residual, int res_dtype) {
  Seg segs[GC_MAX_SEGMENTS];
  for (int s = 0; s < w.num_segments; ++s) segs[s] = Seg{a[s], w.k1[s], w.w1[s], w.k1[s]};
  GC_TRY(run_gemm(c, st, segs, w.num_segments, rows, c.L, h, c.dtype, w.b1, GC_ACT_SWISH, nullptr, 0, nullptr, nullptr, nullptr,
                  nullptr, true));
  if (c.fuse_ln)
    return gc_linear_ln_cond(st, h, c.L, rows, w.w2, c.L, w.b2, so, 1, residual, res_dtype, residual != nullptr ? c.L : 0, out,
                             out_dtype, c.L, c.L);
  const Seg s2{h, c.L, w.w2, c.L};
  GC_TRY(run_gemm(c, st, &s2, 1, rows, c.L, y, c.dtype, w.b2, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true));
  return gc_ln_cond(st, y, c.dtype, c.L, so, 1, residual, res_dtype, residual != nullptr ? c.L : 0, out, out_dtype, c.L, rows, c.L);
}

}  // namespace
}  // namespace gc

extern "C" int gc_sizeof_forward_structs(int32_t which) {
  switch (which) {
    case 0: return static_cast<int>(sizeof(gc_denoiser_model));
    case 1: return static_cast<int>(sizeof(gc_denoiser_graph));
    case 2: return static_cast<int>(sizeof(gc_sigma_context));
    case 3: return static_cast<int>(sizeof(gc_denoiser_workspace));
    case 4: return static_cast<int>(sizeof(gc_mlp2));
    case 5: return static_cast<int>(sizeof(gc_transformer_layer));
    default: return -1;
  }
}

extern "C" int gc_denoiser_forward(void* stream, const gc_denoiser_model* m, const gc_denoiser_graph* g,
                                   const gc_sigma_context* sc, const gc_denoiser_workspace* ws) {
  using namespace gc;
  GC_REQUIRE(m && g && sc && ws, "gc_denoiser_forward: null argument");
  GC_REQUIRE(m->dtype == GC_BF16 || m->dtype == GC_F32, "gc_denoiser_forward: dtype=%d", m->dtype);
  GC_REQUIRE(m->latent == 128 || m->latent == 256 || m->latent == 512, "gc_denoiser_forward: latent=%d", m->latent);
  GC_REQUIRE(m->num_layers >= 0 && (m->num_layers == 0 || m->layers != nullptr), "gc_denoiser_forward: transformer layers");
  GC_REQUIRE(sc->table != nullptr && sc->m0 != nullptr && sc->m_p != nullptr, "gc_denoiser_forward: incomplete sigma context");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaStream_t bs = reinterpret_cast<cudaStream_t>(ws->branch_stream);
  const bool branch = bs != nullptr && ws->fork_event != nullptr && ws->join_event != nullptr;
  const Ctx c{st, m->dtype, m->latent, (ws->flags & GC_FORWARD_FUSE_LN) != 0 && m->dtype == GC_BF16};
  const int L = m->latent, dt = m->dtype;
  const int64_t G = g->grid_rows, V = g->mesh_rows, E1 = g->g2m_edges, E2 = g->m2g_edges;
  auto T = [&](int row) { return sc->table + static_cast<int64_t>(row) * 2 * L; };
  const int c_tfinal = GC_COND_TRANSFORMER0 + 2 * m->num_layers;
  const int c_m2g_eu = c_tfinal + 2, c_m2g_gu = c_tfinal + 3;

  // ---- encoder (gencast/denoiser.py:602-688): grid-node embedding
  {
    const void* a[2] = {ws->xin, ws->a_const};
    GC_TRY(mlp_ln(c, st, m->grid_embed, a, G, ws->g_h, ws->g_y, ws->g0, dt, T(GC_COND_G2M_GRID_EMBED), nullptr, 0));
  }
  // grid branch: grid-node update of the encoder + the decoder's per-grid-node partial product
  auto grid_branch = [&](cudaStream_t s) -> int {
    const void* a[1] = {ws->g0};
    GC_TRY(mlp_ln(c, s, m->grid_update, a, G, ws->g_h2, ws->g_y2, ws->g_lat, dt, T(GC_COND_G2M_GRID_UPDATE), ws->g0, dt));
    const Seg sg{ws->g_lat, L, m->m2g_w1r, L};
    return run_gemm(c, s, &sg, 1, G, L, ws->g_p2, dt, nullptr, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true);
  };
  if (branch) {
    GC_CHECK_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->fork_event), st), "cudaEventRecord(fork)");
    GC_CHECK_CUDA(cudaStreamWaitEvent(bs, reinterpret_cast<cudaEvent_t>(ws->fork_event), 0), "cudaStreamWaitEvent(fork)");
    GC_TRY(grid_branch(bs));
    GC_CHECK_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ws->join_event), bs), "cudaEventRecord(join)");
  }
  // grid2mesh edge update: [e | n_s | n_r] W1 = e W1e + (n_s W1s)[senders] + (n_r W1r)[receivers]
  {
    const Seg sp{ws->g0, L, m->g2m_w1s, L};
    GC_TRY(run_gemm(c, st, &sp, 1, G, L, ws->g_p, dt, nullptr, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true));
    const bool fuse_g2m = (ws->flags & GC_FORWARD_FUSE_G2M) != 0 && sc->g2m_base != nullptr && dt == GC_BF16;
    if (fuse_g2m) {
      // the table already holds e' W1e' + b1 + (m0 W1r)[receivers]: hidden layer, second layer and bias in one kernel
      // the rows' LayerNorm statistics go to the hidden-layer buffer, which this path does not use (16 bytes per edge)
      GC_TRY(gc_edge_mlp_rows(st, sc->g2m_base, L, sc->g2m_base_rows, ws->g_p, g->g2m_senders, L, GC_ACT_SWISH, m->g2m_w2, L,
                              m->g2m_b2, ws->e_y, L, E1, L, reinterpret_cast<float*>(ws->e_h)));
    } else if (sc->g2m_base != nullptr) {
      GC_TRY(gc_edge_hidden(st, sc->g2m_base, L, sc->g2m_base_rows, ws->g_p, g->g2m_senders, L, sc->m_p, g->g2m_receivers, L,
                            GC_ACT_SWISH, ws->e_h, L, E1, L));
    } else {
      const Seg se{g->g2m_edge_ln, L, sc->g2m_w1e, L};
      GC_TRY(run_gemm(c, st, &se, 1, E1, L, ws->e_h, dt, sc->g2m_b1, GC_ACT_SWISH, nullptr, 0, ws->g_p, g->g2m_senders, sc->m_p,
                      g->g2m_receivers, false));
    }
    if (!fuse_g2m) {
      const Seg s2{ws->e_h, L, m->g2m_w2, L};
      GC_TRY(run_gemm(c, st, &s2, 1, E1, L, ws->e_y, dt, m->g2m_b2, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true));
    }
    GC_TRY(gc_ln_cond_segment_sum_stats(st, ws->e_y, dt, L, T(GC_COND_G2M_EDGE_UPDATE), 1 | GC_SEGSUM_IRREGULAR, g->g2m_row_ptr,
                                        g->g2m_perm, ws->m_agg, dt, L, V, L,
                                        fuse_g2m ? reinterpret_cast<const float*>(ws->e_h) : nullptr));
  }
  // mesh-node update (+ residual) -> fp32 transformer stream
  {
    const void* a[2] = {sc->m0, ws->m_agg};
    GC_TRY(mlp_ln(c, st, m->mesh_update, a, V, ws->m_h, ws->m_y, ws->x, GC_F32, T(GC_COND_G2M_MESH_UPDATE), sc->m0, dt));
  }
  if (!branch) GC_TRY(grid_branch(st));

  // ---- processor (gencast/sparse_transformer.py:486-525, :624-634)
  for (int i = 0; i < m->num_layers; ++i) {
    const gc_transformer_layer& l = m->layers[i];
    GC_TRY(gc_ln_cond(st, ws->x, GC_F32, L, T(GC_COND_TRANSFORMER0 + 2 * i), 1, nullptr, 0, 0, ws->t_h, dt, L, V, L));
    const Seg sq{ws->t_h, L, l.wqkv, L};
    GC_TRY(run_gemm(c, st, &sq, 1, V, 3 * L, ws->t_qkv, dt, nullptr, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true));
    if (g->attention_kind == GC_ATTENTION_GATHER) {
      GC_TRY(gc_khop_attention_gather(st, ws->t_qkv, 3 * L, g->step_ptr, g->keys, g->step_mask, g->work, g->num_q_tiles,
                                      g->mask_period, ws->t_o, L, V, m->heads, m->head_dim));
    } else if (g->attention_kind == GC_ATTENTION_TILES) {
      GC_TRY(gc_khop_attention_tiles(st, ws->t_qkv, 3 * L, g->tile_ptr, g->tile_kv, g->tile_mask, ws->t_o, L, V, m->heads,
                                     m->head_dim));
    } else {
      GC_TRY(gc_khop_attention(st, ws->t_qkv, dt, 3 * L, g->nbr_ptr, g->nbr_idx, g->max_degree, ws->t_o, L, V, m->heads,
                               m->head_dim));
    }
    const Seg so{ws->t_o, L, l.wo, L};
    GC_TRY(run_gemm(c, st, &so, 1, V, L, ws->x, GC_F32, l.bo, GC_ACT_NONE, ws->x, GC_F32, nullptr, nullptr, nullptr, nullptr, true));
    GC_TRY(gc_ln_cond(st, ws->x, GC_F32, L, T(GC_COND_TRANSFORMER0 + 2 * i + 1), 1, nullptr, 0, 0, ws->t_h, dt, L, V, L));
    const Seg s1{ws->t_h, L, l.w1, L};
    GC_TRY(run_gemm(c, st, &s1, 1, V, m->ffw_hidden, ws->t_f, dt, l.b1, GC_ACT_GELU_TANH, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                    true));
    const Seg s2{ws->t_f, m->ffw_hidden, l.w2, m->ffw_hidden};
    GC_TRY(run_gemm(c, st, &s2, 1, V, L, ws->x, GC_F32, l.b2, GC_ACT_NONE, ws->x, GC_F32, nullptr, nullptr, nullptr, nullptr, true));
  }
  GC_TRY(gc_ln_cond(st, ws->x, GC_F32, L, T(c_tfinal), 1, nullptr, 0, 0, ws->m_out, dt, L, V, L));

  // ---- decoder (gencast/denoiser.py:730-768)
  {
    const Seg sp{ws->m_out, L, m->m2g_w1s, L};
    GC_TRY(run_gemm(c, st, &sp, 1, V, L, ws->m_p, dt, nullptr, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true));
  }
  if (branch) GC_CHECK_CUDA(cudaStreamWaitEvent(st, reinterpret_cast<cudaEvent_t>(ws->join_event), 0), "cudaStreamWaitEvent(join)");
  const bool fuse = (ws->flags & GC_FORWARD_FUSE_M2G) != 0 && g->m2g_perm == nullptr && dt == GC_BF16;
  if (fuse) {
    const void* base = sc->m2g_base;
    int64_t base_rows = sc->m2g_base_rows;
    if (base == nullptr) {
      const Seg se{g->m2g_edge_ln, L, sc->m2g_w1e, L};
      GC_TRY(run_gemm(c, st, &se, 1, E2, L, ws->e_h, dt, sc->m2g_b1, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, false));
      base = ws->e_h;
      base_rows = E2;
    }
    // m2g_perm == NULL says the edges are stored grid-major, three per grid node: receiver of edge e = e / 3, which is
    // what a null receiver table means to gc_edge_mlp_sum3 (its receiver rows then travel by TMA)
    GC_TRY(gc_edge_mlp_sum3(st, base, L, base_rows, ws->m_p, g->m2g_senders, L, ws->g_p2, nullptr, L, GC_ACT_SWISH,
                            m->m2g_w2, L, m->m2g_b2, T(c_m2g_eu), 1, ws->g_agg, dt, L, G, L));
  } else {
    if (sc->m2g_base != nullptr) {
      GC_TRY(gc_edge_hidden(st, sc->m2g_base, L, sc->m2g_base_rows, ws->m_p, g->m2g_senders, L, ws->g_p2, g->m2g_receivers, L,
                            GC_ACT_SWISH, ws->e_h, L, E2, L));
    } else {
      const Seg se{g->m2g_edge_ln, L, sc->m2g_w1e, L};
      GC_TRY(run_gemm(c, st, &se, 1, E2, L, ws->e_h, dt, sc->m2g_b1, GC_ACT_SWISH, nullptr, 0, ws->m_p, g->m2g_senders, ws->g_p2,
                      g->m2g_receivers, false));
    }
    const Seg s2{ws->e_h, L, m->m2g_w2, L};
    GC_TRY(run_gemm(c, st, &s2, 1, E2, L, ws->e_y, dt, m->m2g_b2, GC_ACT_NONE, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true));
    GC_TRY(gc_ln_cond_segment_sum(st, ws->e_y, dt, L, T(c_m2g_eu), 1, g->m2g_row_ptr, g->m2g_perm, ws->g_agg, dt, L, G, L));
  }
  {
    const void* a[2] = {ws->g_lat, ws->g_agg};
    GC_TRY(mlp_ln(c, st, m->m2g_grid_update, a, G, ws->g_h, ws->g_y, ws->g2, dt, T(c_m2g_gu), ws->g_lat, dt));
  }
  // output MLP L -> L -> n_out (no LayerNorm, no conditioning: common/deep_typed_graph_net.py:469-485)
  {
    const Seg s1{ws->g2, L, m->output.w1[0], L};
    GC_TRY(run_gemm(c, st, &s1, 1, G, L, ws->g_h, dt, m->output.b1, GC_ACT_SWISH, nullptr, 0, nullptr, nullptr, nullptr, nullptr, true));
    const Seg s2{ws->g_h, L, m->output.w2, L};
    GC_TRY(run_gemm(c, st, &s2, 1, G, m->n_out_padded, ws->f_out, GC_F32, m->output.b2, GC_ACT_NONE, nullptr, 0, nullptr, nullptr,
                    nullptr, nullptr, true));
  }
  return GC_OK;
}
