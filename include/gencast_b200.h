/*
 * gencast_b200 — C ABI of the B200 (sm_100a) GenCast denoiser / sampler kernels.
 *
 * The reference (fgiral000/gencast-flax-nnx) is pure Python/JAX: it has no FFI
 * layer of its own.  Its hot path is made of the Python operators cited next to
 * each entry point below; these entry points are what an XLA FFI custom call
 * (jax.ffi.ffi_call) or any other host binding would bind in their place — see
 * INTEGRATION.md for the jax.ffi stub and gencast_flax_nnx_b200/_lib.py for the
 * ctypes binding used by this repo's own host code and tests.
 *
 * Conventions
 *   - every launcher is enqueue-only on `stream` (a cudaStream_t passed as
 *     void*): no allocation, no synchronisation, no default-stream work, no
 *     global mutable state, so calls are re-entrant and CUDA-graph capturable;
 *   - all buffers are caller-owned device pointers; row-major, leading
 *     dimension (`ld*`) counted in elements;
 *   - return value 0 = success; negative = error, message via gc_last_error()
 *     (thread-local).  Nothing throws across the boundary;
 *   - dtype codes: GC_F32 = 0 (float), GC_BF16 = 1 (__nv_bfloat16).
 */
#ifndef GENCAST_B200_H_
#define GENCAST_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define GC_API __attribute__((visibility("default")))
#else
#define GC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GC_F32 0
#define GC_BF16 1

#define GC_ACT_NONE 0
#define GC_ACT_SWISH 1      /* jax.nn.swish, common/deep_typed_graph_net.py:61-62 */
#define GC_ACT_GELU_TANH 2  /* jax.nn.gelu(approximate=True), gencast/sparse_transformer.py:264 */

#define GC_OK 0
#define GC_ERR_INVALID_ARGUMENT (-1)
#define GC_ERR_CUDA (-2)
#define GC_ERR_UNSUPPORTED (-3)

#define GC_GEMM_STATIC_WEIGHTS 1
#define GC_MAX_SEGMENTS 3

/* Library / device introspection. */
GC_API const char* gc_last_error(void);
GC_API int gc_abi_version(void);
/* 1 if the current device can run the tcgen05 path (compute capability 10.x). */
GC_API int gc_device_supports_tcgen05(void);

/*
 * Fused linear layer:  OUT = post( act( alpha * sum_s A_s . W_s^T + bias + addend
 *                                       + gather0[idx0[row]] + gather1[idx1[row]] ) ) + residual
 *
 * Replaces flax.nnx.Linear inside MLP / MLPWithNormConditioning
 * (common/mlp.py:152-203), the [e | s | r] concatenation + first edge-MLP
 * layer of the interaction network (common/typed_graph_net.py:134-159,
 * :301-305 — the concatenation becomes K-segments and row gathers of
 * per-node partial products), the [n | agg] node update (:315-326), and the
 * transformer projections with residual add (gencast/sparse_transformer.py:
 * 252-290, :353, :520-524).
 *
 * A_s: [m, k[s]] (ld lda[s]);  W_s: [n, k[s]] i.e. the Linear kernel stored
 * transposed, K contiguous (ld ldw[s]).  dtype selects the operand type and the
 * engine: GC_BF16 -> tcgen05/TMEM/TMA tensor-core kernel (fp32 accumulate),
 * GC_F32 -> fp32 FFMA kernel.  k[s] % 64 == 0, n % 128 == 0, pointers 16-byte
 * aligned, ld % 8 == 0.
 */
typedef struct gc_gemm_args {
  const void* a[GC_MAX_SEGMENTS];
  const void* w[GC_MAX_SEGMENTS];
  int64_t lda[GC_MAX_SEGMENTS];
  int64_t ldw[GC_MAX_SEGMENTS];
  int32_t k[GC_MAX_SEGMENTS];
  int32_t num_segments;
  int64_t m;
  int32_t n;
  int32_t dtype;            /* operand dtype of every A_s and W_s */
  const float* bias;        /* [n] or NULL */
  const float* alpha_dev;   /* device scalar or NULL (= 1) */
  const void* addend;       /* [m, n] added before the activation, or NULL */
  int64_t ld_addend;
  int32_t addend_dtype;
  int32_t gather_dtype;
  const void* gather_src[2];     /* tables gathered by row index, added before the activation, or NULL */
  const int32_t* gather_idx[2];  /* [m] row ids into gather_src[j] */
  int64_t ld_gather[2];
  int32_t act;              /* GC_ACT_* */
  int32_t res_dtype;
  const void* residual;     /* [m, n] added after the activation, or NULL */
  int64_t ld_res;
  void* out;                /* [m, n] */
  int64_t ldo;
  int32_t out_dtype;
  int32_t flags;                 /* GC_GEMM_STATIC_WEIGHTS: the W matrices are not written by earlier work on this
                                    stream, so their first tiles may be fetched before the preceding kernel has
                                    finished (under programmatic dependent launch); 0 = no assumption */
} gc_gemm_args;

GC_API int gc_gemm(void* stream, const gc_gemm_args* args);
/* sizeof(gc_gemm_args) as compiled into the library, for binding-side layout checks. */
GC_API int gc_sizeof_gemm_args(void);

/*
 * Row LayerNorm (no learned affine, eps 1e-6, var = E[x^2]-E[x]^2 >= 0) followed
 * by the conditional affine  y * scale + offset  (+ residual).
 * Replaces nnx.LayerNorm + LinearNormConditioning: common/mlp.py:59-65,
 * :121-145; gencast/sparse_transformer.py:518-523, :630-633, and the residual
 * adds of common/deep_typed_graph_net.py:569-581.
 * scale_offset: [2*cols] float = (1 + s | o) for this call's noise level
 * (produced by gc_cond_tables).  do_layer_norm = 0 applies the affine only.
 */
GC_API int gc_ln_cond(void* stream, const void* x, int32_t x_dtype, int64_t ldx,
               const float* scale_offset, int32_t do_layer_norm,
               const void* residual, int32_t res_dtype, int64_t ld_res,
               void* out, int32_t out_dtype, int64_t ldo,
               int64_t rows, int32_t cols);

/*
 * Deterministic receiver-sorted segment sum with LayerNorm + conditional affine
 * applied to every edge row on the fly:
 *   out[v] = sum_{j in [row_ptr[v], row_ptr[v+1])}  LN(y[edge_perm[j]]) * scale + offset
 * Replaces the tail of the edge MLP (common/mlp.py:121-145) together with
 * jraph.segment_sum (call sites common/typed_graph_net.py:173-182, f32
 * aggregation of common/deep_typed_graph_net.py:396-404).  The sum runs in
 * fp32 in a fixed order (no atomics) so results are bitwise reproducible.
 * edge_perm may be NULL (edges already receiver-sorted: mesh2grid, and grid2mesh as the engine stores it).
 * do_layer_norm: bit 0 = apply LayerNorm; bit 1 (GC_SEGSUM_IRREGULAR) = hint that segment lengths vary widely
 * (selects the higher-occupancy kernel variant; implied by a non-NULL edge_perm).
 */
#define GC_SEGSUM_IRREGULAR 2
GC_API int gc_ln_cond_segment_sum(void* stream, const void* y, int32_t y_dtype, int64_t ldy,
                           const float* scale_offset, int32_t do_layer_norm,
                           const int32_t* row_ptr, const int32_t* edge_perm,
                           void* out, int32_t out_dtype, int64_t ldo,
                           int64_t num_segments, int32_t cols);
/*
 * The same with the rows' LayerNorm statistics supplied by the kernel that produced y (gc_edge_mlp_rows): row_stats is
 * fp32 [rows of y, 4] = {sum, sum of squares} of the first and of the second half of the columns (NULL = the plain entry
 * above).  Mean and variance are then formed from the producer's fp32 accumulator instead of from the rounded bf16 row,
 * and the kernel no longer spends its instructions on reductions (bf16 rows with LayerNorm only).
 */
GC_API int gc_ln_cond_segment_sum_stats(void* stream, const void* y, int32_t y_dtype, int64_t ldy,
                           const float* scale_offset, int32_t do_layer_norm,
                           const int32_t* row_ptr, const int32_t* edge_perm,
                           void* out, int32_t out_dtype, int64_t ldo,
                           int64_t num_segments, int32_t cols, const float* row_stats);

/*
 * k-hop neighbourhood multi-head attention on the mesh (exact sparse pattern):
 *   out[i, h] = softmax_j( q[i,h] . k[j,h] / sqrt(d) ) v[j,h],  j in neighbours(i)
 * Replaces TriblockdiagMHA without its projections
 * (gencast/sparse_transformer.py:323-351, softmax :100-125, mask :163-201): the
 * tri-block mask evaluates to exactly this neighbour set, masked logits
 * contribute exp(-1e30 - max) = 0.
 * qkv: [nodes, 3*heads*head_dim] laid out (q | k | v); out: [nodes, heads*head_dim].
 * nbr_ptr/nbr_idx: CSR of the k-hop pattern (self included).
 */
GC_API int gc_khop_attention(void* stream, const void* qkv, int32_t dtype, int64_t ld_qkv,
                      const int32_t* nbr_ptr, const int32_t* nbr_idx, int32_t max_degree,
                      void* out, int64_t ldo, int64_t nodes, int32_t heads, int32_t head_dim);

/*
 * The same attention on tcgen05 tensor cores (bf16 only), driven by a block-sparse
 * tile list instead of a CSR: for query tile t (128 consecutive nodes) the pairs
 * tile_ptr[t] .. tile_ptr[t+1]-1 name the 128-key tiles tile_kv[] that contain at
 * least one neighbour, and tile_mask holds one 128 x 128 bit mask per pair
 * ([pair][row][4] uint32, bit (j % 32) of word (j / 32) = key tile_kv * 128 + j is a
 * neighbour of query t * 128 + row).  Same reference operator as gc_khop_attention.
 * tile_kv[p]: bits 0-23 the key tile index; optionally bits 24-25 / 26-27 the number of
 * leading / trailing 32-key sub-blocks of the pair that no query attends to (their mask bits
 * must be zero): S and P V are then formed over the live key range only.  0 = whole tile.
 * head_dim 64 or 128.
 */
GC_API int gc_khop_attention_tiles(void* stream, const void* qkv, int64_t ld_qkv, const int32_t* tile_ptr,
                                   const int32_t* tile_kv, const uint32_t* tile_mask, void* out, int64_t ldo,
                                   int64_t nodes, int32_t heads, int32_t head_dim);

/*
 * The same attention over per-query-tile COMPACTED key lists (bf16, head_dim 64 or 128): for query tile t (128
 * consecutive nodes) steps step_ptr[t] .. step_ptr[t+1]-1 each name 64 rows of qkv (keys[step * 64 + j]; the sorted
 * union of the tile's neighbours, the last step padded with any valid row) and mask[step][row] holds 64 bits: bit j
 * set = keys[step * 64 + j] is a neighbour of query t * 128 + row (padded columns: clear).  The kernel gathers the
 * K / V rows into dense tensor-core tiles itself, so a 128-query patch of the 1 deg mesh costs 11 steps of 64 keys
 * instead of 11.3 key tiles of 128.  mask_period > 0: mask index = step % mask_period (ensemble members evaluated
 * together share the masks; their key lists differ by a row offset).  work[num_q_tiles]: launch order of the query
 * tiles.  Same reference operator as gc_khop_attention (gencast/sparse_transformer.py:309-354).
 */
GC_API int gc_khop_attention_gather(void* stream, const void* qkv, int64_t ld_qkv, const int32_t* step_ptr,
                                    const int32_t* keys, const uint32_t* mask, const int32_t* work,
                                    int32_t num_q_tiles, int32_t mask_period, void* out, int64_t ldo, int64_t nodes,
                                    int32_t heads, int32_t head_dim);

/*
 * Noise-level conditioning for a batch of noise levels, all layers at once:
 *   cond      = Linear1(gelu_tanh(Linear0(fourier(log sigma))))     [16]
 *   table[i, l] = (1 + s_l | o_l),  [s_l | o_l] = cond . Wc_l + bc_l   [2*width]
 * Replaces FourierFeaturesMLP.__call__ (common/mlp.py:255-265,
 * common/model_utils.py:728-757) and every LinearNormConditioning linear
 * (common/mlp.py:59-64).  Weights fp32: w0 [2*num_freq, 32], b0 [32],
 * w1 [32, 16], b1 [16], wc [layers, 16, 2*width], bc [layers, 2*width].
 * table: [num_sigma, layers, 2*width] fp32.
 */
GC_API int gc_cond_tables(void* stream, const float* sigma, int32_t num_sigma,
                   const float* w0, const float* b0, const float* w1, const float* b1,
                   float base_period, int32_t num_frequencies,
                   const float* wc, const float* bc, int32_t layers, int32_t width,
                   float* table);

/*
 * Folds a conditional affine that sits in front of a Linear into that Linear:
 *   w_out[n, k] = w[n, k] * scale[k];   bias_out[n] = bias[n] + sum_k offset[k] * w[n, k]
 * so that Linear(x * scale + offset) == x . w_out^T + bias_out.  Used for the
 * statically embedded edge latents, whose LayerNorm output is a constant and only
 * the noise-level affine changes per call (common/mlp.py:62-65 feeding
 * common/typed_graph_net.py:301-305).  w: [n, k] (dtype), bias fp32 or NULL.
 */
GC_API int gc_fold_affine_into_linear(void* stream, const void* w, int32_t dtype, int64_t ldw,
                               const float* bias, const float* scale_offset,
                               void* w_out, int64_t ldw_out, float* bias_out,
                               int32_t n, int32_t k);

/*
 * Preconditioning + DPM-Solver++ 2S state update, elementwise:
 *   D     = c_out * f + c_skip * x_cur
 *   x_new = a * x_base + (1 - a) * D
 *   x_out = x_new (fp32);   xin_out = c_in_next * x_new (operand dtype)
 * Replaces Sampler._preconditioned_denoiser and the two update lines of body_fn
 * (gencast/dpm_solver_plus_plus_2s.py:145-153, :198-205).  sched: device
 * float[4] = (c_out, c_skip, a, c_in_next).  All arrays [rows, cols] with the
 * given leading dimensions; xin_out may be NULL.
 */
GC_API int gc_dpm_update(void* stream, const float* f, int64_t ldf, const float* x_cur, const float* x_base,
                  int64_t ldx, const float* sched, float* x_out, void* xin_out, int32_t xin_dtype,
                  int64_t ld_xin, int64_t rows, int32_t cols);

/*
 * 2-D convert / pad / scale: dst[r, c] = c < cols_src ? scale * src[r, c] : 0.
 * Layout glue of gencast/denoiser.py:794-806 (feature assembly) on device.
 * scale_dev may be NULL (= 1).
 */
GC_API int gc_cast_pad(void* stream, const void* src, int32_t src_dtype, int64_t ld_src, int32_t cols_src,
                void* dst, int32_t dst_dtype, int64_t ld_dst, int32_t cols_dst,
                const float* scale_dev, int64_t rows);

/*
 * Column selection from up to three fp32 matrices: out[r, j] = src_k[r, c] with
 * (k, c) = (table[j] >> 24, table[j] & 0xffffff).  This is the autoregressive window
 * update of the reference's rollout on device -- common/rollout.py:362-368 and
 * _get_next_inputs :379-401: the next step's stacked inputs are columns of the
 * previous inputs (older frames shift), of the prediction and of the step's forcings.
 * src1 / src2 may be NULL if the table does not reference them; out must not alias a source.
 */
GC_API int gc_select_columns(void* stream, const float* src0, int64_t ld0, const float* src1, int64_t ld1,
                      const float* src2, int64_t ld2, const int32_t* table, float* out, int64_t ldo,
                      int64_t rows, int32_t cols_out);

/*
 * Input side of the reference's predictor wrappers on the device-resident rollout window: per stacked channel c
 *   dst[r, c] = cast( fill_post_c( ( fill_pre_c(src[r, c]) - loc[c] ) / scale[c] ) ),   fill_x(v) = isnan(v) ? fill_x[c] : v
 * i.e. normalization.normalize (common/normalization.py:31-50, called at :154-155, :216-217) with NaNCleaner._clean
 * (gencast/nan_cleaning.py:47-53) on either side of it, written straight into the denoiser's (padded, possibly bf16)
 * constant-feature operand.  Any of loc / scale / fill_pre / fill_post may be NULL (identity); a NaN fill keeps NaNs.
 * Exactly rounded fp32 subtract and divide: bitwise what the reference's array arithmetic gives.
 */
GC_API int gc_normalize_cast(void* stream, const float* src, int64_t ld_src, int32_t cols, const float* loc,
                             const float* scale, const float* fill_pre, const float* fill_post, void* dst,
                             int32_t dst_dtype, int64_t ld_dst, int64_t rows);

/*
 * Output side: un-normalise the network's prediction, add the last input frame where the variable is also an input
 * (InputsAndResiduals._unnormalize_prediction_and_add_input, common/normalization.py:114-133) and optionally put NaNs
 * back where the inputs had them (NaNCleaner._maybe_reintroduce_nans, gencast/nan_cleaning.py:55-64):
 *   out[r, c] = pred[r, c] * scale[c] (+ loc[c]) (+ fill(window[r, res_col[c]], res_fill[c]) if res_col[c] >= 0);
 *   out[r, c] = NaN if window[r, nan_cols[c * nan_per_col + k]] is NaN for some k (entries < 0 are skipped).
 * `window` is the stacked physical input window of the step ([rows, ld_window] fp32).  Separate, exactly rounded fp32
 * multiply and adds (no contraction).  scale / loc / res_col / res_fill / nan_cols may be NULL.
 */
GC_API int gc_unnormalize_residual(void* stream, const float* pred, int64_t ld_pred, int32_t cols, const float* scale,
                                   const float* loc, const float* window, int64_t ld_window, const int32_t* res_col,
                                   const float* res_fill, const int32_t* nan_cols, int32_t nan_per_col, float* out,
                                   int64_t ldo, int64_t rows);

/*
 * Hidden layer of an edge MLP whose edge-feature part is known in advance:
 *   out[e, :] = act( base[e % period, :] + g0[idx0[e], :] + g1[idx1[e], :] )        (bf16 in / out, fp32 sum)
 * The reference's edge update is MLP([e | n_s | n_r]) (common/typed_graph_net.py:134-159, :301-305); its first
 * layer splits into e @ W1e + n_s @ W1s + n_r @ W1r, and in GenCast's encoder / decoder the edge features e are
 * static structural embeddings (gencast/denoiser.py:662-675, :753-755) whose only run-time dependence is the
 * noise-level conditioning, so base = e' @ W1e' + b1 is a table per noise level, shared by all ensemble members
 * (period = edges of one member; rows = members * period, member-major).  g1 / idx1 may be NULL.
 * cols in {128, 256, 512}; act as in gc_gemm.
 */
GC_API int gc_edge_hidden(void* stream, const void* base, int64_t ld_base, int64_t period, const void* g0,
                   const int32_t* idx0, int64_t ld0, const void* g1, const int32_t* idx1, int64_t ld1,
                   int32_t act, void* out, int64_t ldo, int64_t rows, int32_t cols);

/*
 * Second MLP layer + LayerNorm + conditional affine (+ residual) in one kernel (bf16 operands, tcgen05):
 *   out = LayerNorm(a . w^T + bias) * scale + offset (+ residual)
 * i.e. the tail of MLPWithNormConditioning (common/mlp.py:121-147) and the residual add of
 * common/deep_typed_graph_net.py:569-581 without the pre-norm activations ever leaving the SM: every tile holds whole
 * rows (128 x cols fp32) in tensor memory, the epilogue reads them twice (statistics, normalise).  a: [rows, cols] bf16,
 * w: [cols, cols] bf16 stored [out, in]; bias [cols], scale_offset [2 cols] = (1 + s | o) fp32 or NULL;
 * residual [rows, cols] bf16 / fp32 or NULL; out bf16 / fp32.  cols in {128, 256, 512}.
 */
GC_API int gc_linear_ln_cond(void* stream, const void* a, int64_t lda, int64_t rows, const void* w, int64_t ldw,
                             const float* bias, const float* scale_offset, int32_t do_layer_norm, const void* residual,
                             int32_t res_dtype, int64_t ld_res, void* out, int32_t out_dtype, int64_t ldo, int32_t cols);

/*
 * Fused edge update + aggregation for a bipartite graph whose receivers have exactly three incoming edges each,
 * stored receiver-major (edges 3v, 3v+1, 3v+2 -> receiver v; GenCast's mesh2grid decoder,
 * common/grid_mesh_connectivity.py:104, :125-131).  One kernel, no [E, cols] tensor in HBM:
 *   h_e   = act( base[e mod period] + gs[idx_s[e]] + gr[idx_r[e]] )         first edge-MLP layer, split by operand
 *   y_e   = h_e . W2^T + b2                                                 second layer (tcgen05, fp32 accumulate)
 *   out_v = scale * sum_{e in 3v..3v+2} LayerNorm(y_e) + 3 * offset         LN + conditional affine + segment sum
 * Replaces EdgeWrapper / MLPWithNormConditioning of the edge update (common/typed_graph_net.py:134-159, :295-305;
 * common/mlp.py:115-147) together with jraph.segment_sum (common/typed_graph_net.py:161-195,
 * common/deep_typed_graph_net.py:396-410).  base / gs / gr / w2 are bf16; w2 is [cols, cols] stored [out, in];
 * b2 [cols] and scale_offset [2 cols] = (1 + s | o) are fp32 (either may be NULL); out is bf16 or fp32,
 * [num_receivers, cols]; idx_s / idx_r have 3 * num_receivers entries.  cols in {128, 256, 512}.
 * idx_r == NULL means "the receiver's own row": gr[e / 3] (the mesh2grid case).  Then, when `period` is a multiple of
 * 30, the base rows and the receiver rows of a tile are contiguous and are moved by TMA (only gs is gathered).
 */
GC_API int gc_edge_mlp_sum3(void* stream, const void* base, int64_t ld_base, int64_t period, const void* gs,
                            const int32_t* idx_s, int64_t ld_gs, const void* gr, const int32_t* idx_r, int64_t ld_gr,
                            int32_t act, const void* w2, int64_t ld_w2, const float* b2, const float* scale_offset,
                            int32_t do_layer_norm, void* out, int32_t out_dtype, int64_t ldo, int64_t num_receivers,
                            int32_t cols);

/*
 * Edge MLP whose first layer is known up to one gathered operand, rows out (no [E, cols] hidden tensor in HBM):
 *   y_e = act( base[e mod period] + gs[idx_s[e]] ) . W2^T + b2              bf16 [num_rows, cols]
 * for edge lists where the receiver's contribution to the first layer is static and already part of `base`
 * (GenCast's grid2mesh encoder: the receivers are mesh nodes, whose embedding depends on the noise level only,
 * gencast/denoiser.py:662-675; the table is then e' W1e' + b1 + (m0 W1r)[receivers]).  Same operators as
 * gc_edge_mlp_sum3 (common/typed_graph_net.py:134-159, :295-305; common/mlp.py:115-147 up to the LayerNorm), same
 * kernel: tiles of 128 consecutive edges of one member, base rows by TMA, gs gathered, W2 by TMA, whole rows in tensor
 * memory; LayerNorm + aggregation follow in gc_ln_cond_segment_sum.  num_rows is a multiple of period (members).
 * row_stats (optional, fp32 [num_rows, 4], 16-byte aligned): per row {sum, sum of squares} of y over the first and over the
 * second half of the columns, taken from the fp32 accumulator; gc_ln_cond_segment_sum_stats then skips its own reductions.
 */
GC_API int gc_edge_mlp_rows(void* stream, const void* base, int64_t ld_base, int64_t period, const void* gs,
                            const int32_t* idx_s, int64_t ld_gs, int32_t act, const void* w2, int64_t ld_w2,
                            const float* b2, void* out, int64_t ldo, int64_t num_rows, int32_t cols, float* row_stats);

/*
 * Inverse real spherical-harmonic transform of random coefficients -> isotropic white noise fields in the sampler's
 * state layout.  Replaces dinosaur's RealSphericalHarmonics.to_nodal as called by the reference's noise generator
 * (gencast/samplers_utils.py:99-118, :250-346; the per-l amplitudes of :316-322 are folded into `table`):
 *   out[(member * n_lat + lat) * n_lon + lon, ch] =
 *       sum_m sum_{l >= m} table[m, l, lat] * ( coef[0, m, f, l] cos(2 pi m lon / n_lon) + coef[1, m, f, l] sin(...) ),
 *   f = member * channels + ch.
 * coef: [2, wavenumbers, members * channels, wavenumbers] fp32 (cos | sin coefficients; entries with l < m and the
 * m = 0 sine row are ignored); table: [wavenumbers (m), wavenumbers (l), n_lat] fp32; spec: scratch
 * [members * n_lat * channels * wavenumbers * 2] fp32; out: [members * n_lat * n_lon, channels] fp32.
 * 2 * wavenumbers <= n_lon; channels * wavenumbers * 8 bytes must fit shared memory.
 */
GC_API int gc_sh_synthesis(void* stream, const float* coef, const float* table, float* spec, float* out,
                           int32_t wavenumbers, int32_t members, int32_t channels, int32_t n_lat, int32_t n_lon);

/*
 * Fair CRPS of an M-member ensemble per grid point and channel (no reference implementation
 * exists; defined in DESIGN.md / parallel.py):
 *   crps[i] = mean_m |x_m[i] - y[i]|  -  sum_{j<k} |x_j[i] - x_k[i]| / (M (M - 1))
 * evaluated with the sorted-sample identity sum_{j<k} |x_j - x_k| = sum_k (2k - M + 1) x_(k).
 * members: [M, n] fp32 (row stride ld_members), truth: [n]; out[i] = weight[i / channels] * crps[i]
 * (weights may be NULL = 1).  2 <= M <= 64.  gc_column_sums then reduces out over grid points in a
 * fixed order: sums[c] = sum_g x[g * channels + c].
 */
GC_API int gc_fair_crps(void* stream, const float* members, int64_t ld_members, int32_t num_members,
                 const float* truth, const float* weights, int32_t channels, float* out, int64_t n);
GC_API int gc_column_sums(void* stream, const float* x, int64_t rows, int32_t cols, float* sums);

/*
 * Ensemble statistics accumulation (no reference implementation exists; defined
 * in DESIGN.md): sum[i] += x[i]; sumsq[i] += x[i]^2 over n elements.
 */
GC_API int gc_ensemble_accumulate(void* stream, const float* x, float* sum, float* sumsq, int64_t n);

/*
 * ---------------------------------------------------------------------------------------------------------------
 * One whole network evaluation in a single call.
 *
 * gc_denoiser_forward sequences the ~135 kernel launches of F(c_in x, sigma) -- DenoiserArchitecture.__call__,
 * gencast/denoiser.py:303-341: grid2mesh GNN (:602-688), mesh transformer (:691-728; gencast/sparse_transformer.py:
 * 486-525, :624-634), mesh2grid GNN (:730-768) -- on `stream`, from descriptors of device pointers that the host
 * fills once.  This is the entry point an XLA FFI custom call (or any non-Python host) binds when it wants the
 * denoiser as one operator; csrc/xla_ffi_shim.cc wraps it.  Enqueue-only, CUDA-graph capturable, no allocation.
 * All matrices are row-major and contiguous (leading dimension = columns); weights are stored [out, in] (the
 * Linear kernel transposed, K contiguous, K padded to a multiple of 64) in the operand dtype `dtype`.
 * ---------------------------------------------------------------------------------------------------------------
 */
#define GC_ATTENTION_CSR 0
#define GC_ATTENTION_TILES 1
#define GC_ATTENTION_GATHER 2

/* rows of gc_sigma_context.table (one (1 + s | o) vector of 2 * latent floats per conditional LayerNorm) */
#define GC_COND_G2M_EDGE_EMBED 0
#define GC_COND_G2M_GRID_EMBED 1
#define GC_COND_G2M_MESH_EMBED 2
#define GC_COND_G2M_EDGE_UPDATE 3
#define GC_COND_G2M_GRID_UPDATE 4
#define GC_COND_G2M_MESH_UPDATE 5
#define GC_COND_TRANSFORMER0 6   /* + 2 i: attention norm of block i, + 2 i + 1: its FFW norm; then the final norm,
                                    the mesh2grid edge embedding, edge update and grid update */

#define GC_FORWARD_FUSE_M2G 1    /* mesh2grid edge update + aggregation through gc_edge_mlp_sum3 */
#define GC_FORWARD_FUSE_LN 2     /* second MLP layer + LayerNorm + affine + residual through gc_linear_ln_cond */
#define GC_FORWARD_FUSE_G2M 4    /* sigma->g2m_base already contains the receivers' part (base[e] + m_p[receivers[e]]):
                                    the grid2mesh edge MLP runs through gc_edge_mlp_rows */

/* Linear -> swish -> Linear (common/mlp.py:152-203); the first layer may be split in K-segments (concatenated
 * operands of the reference, common/typed_graph_net.py:301-305, :315-326) */
typedef struct gc_mlp2 {
  const void* w1[GC_MAX_SEGMENTS];   /* [latent, k1[s]] */
  int32_t k1[GC_MAX_SEGMENTS];
  int32_t num_segments;
  const float* b1;                   /* [latent] */
  const void* w2;                    /* [n2, latent] */
  const float* b2;                   /* [n2] */
} gc_mlp2;

/* one transformer block (gencast/sparse_transformer.py:252-307, :458-525) */
typedef struct gc_transformer_layer {
  const void* wqkv;                  /* [3 latent, latent]: q | k | v projections, no bias (:281) */
  const void* wo; const float* bo;   /* [latent, latent] output projection with bias (:305) */
  const void* w1; const float* b1;   /* [ffw_hidden, latent] */
  const void* w2; const float* b2;   /* [latent, ffw_hidden] */
} gc_transformer_layer;

typedef struct gc_denoiser_model {
  int32_t dtype;                     /* GC_BF16 (tensor cores) or GC_F32 */
  int32_t latent, heads, head_dim, ffw_hidden, num_layers, n_out_padded, reserved;
  gc_mlp2 grid_embed;                /* segments: c_in * noisy targets [KN], per-step constants [KC] */
  const void* g2m_w1s;               /* sender block of the grid2mesh edge MLP's first layer, [latent, latent] */
  const void* g2m_w2; const float* g2m_b2;
  gc_mlp2 mesh_update;               /* segments: embedded mesh nodes, aggregated messages */
  gc_mlp2 grid_update;
  const gc_transformer_layer* layers;   /* host array [num_layers] */
  const void* m2g_w1s; const void* m2g_w1r;   /* sender / receiver blocks of the mesh2grid edge MLP's first layer */
  const void* m2g_w2; const float* m2g_b2;
  gc_mlp2 m2g_grid_update;           /* segments: grid latents, aggregated messages */
  gc_mlp2 output;                    /* latent -> latent -> n_out_padded, no LayerNorm (deep_typed_graph_net.py:469-485) */
} gc_denoiser_model;

typedef struct gc_denoiser_graph {
  int64_t grid_rows, mesh_rows, g2m_edges, m2g_edges;    /* totals over the ensemble members evaluated together */
  const int32_t* g2m_senders; const int32_t* g2m_receivers; const int32_t* g2m_row_ptr; const int32_t* g2m_perm;
  const int32_t* m2g_senders; const int32_t* m2g_receivers; const int32_t* m2g_row_ptr;
  const int32_t* m2g_perm;           /* NULL: edges receiver-major, three per grid node */
  const void* g2m_edge_ln; const void* m2g_edge_ln;      /* LayerNorm'ed static edge embeddings [edges, latent] */
  int32_t attention_kind, max_degree, num_q_tiles, mask_period;
  const int32_t* step_ptr; const int32_t* keys; const uint32_t* step_mask; const int32_t* work;   /* GC_ATTENTION_GATHER */
  const int32_t* tile_ptr; const int32_t* tile_kv; const uint32_t* tile_mask;                      /* GC_ATTENTION_TILES */
  const int32_t* nbr_ptr; const int32_t* nbr_idx;                                                  /* GC_ATTENTION_CSR */
} gc_denoiser_graph;

/* everything that depends on the noise level (and the weights) only */
typedef struct gc_sigma_context {
  const float* table;                /* [num conditional norms, 2 latent], gc_cond_tables */
  const void* g2m_w1e; const float* g2m_b1;   /* edge block of the edge MLPs' first layer with the edge embedding's */
  const void* m2g_w1e; const float* m2g_b1;   /* conditional affine folded in (gc_fold_affine_into_linear) */
  const void* g2m_base; const void* m2g_base; /* e' W1e' + b1 tabulated for this level, or NULL */
  int64_t g2m_base_rows, m2g_base_rows;
  const void* m0; const void* m_p;   /* embedded mesh nodes and their product with the receiver block, [mesh_rows, latent] */
} gc_sigma_context;

typedef struct gc_denoiser_workspace {
  void* xin; void* a_const;          /* inputs: [grid_rows, KN], [grid_rows, KC] */
  void* g_h; void* g_y; void* g_h2; void* g_y2; void* g0; void* g_lat; void* g2; void* g_p; void* g_p2; void* g_agg;
  void* m_h; void* m_y; void* m_p; void* m_agg; void* m_out; void* t_h; void* t_o;   /* [mesh_rows, latent] */
  float* x;                          /* [mesh_rows, latent] fp32 residual stream of the transformer */
  void* t_qkv; void* t_f;            /* [mesh_rows, 3 latent], [mesh_rows, ffw_hidden] */
  void* e_h; void* e_y;              /* [max(g2m_edges, m2g_edges), latent] */
  float* f_out;                      /* output: raw network prediction [grid_rows, n_out_padded] fp32 */
  void* branch_stream; void* fork_event; void* join_event;   /* optional parallel branch (cudaStream_t / cudaEvent_t) */
  int32_t flags; int32_t reserved;
} gc_denoiser_workspace;

GC_API int gc_denoiser_forward(void* stream, const gc_denoiser_model* model, const gc_denoiser_graph* graph,
                               const gc_sigma_context* sigma, const gc_denoiser_workspace* workspace);
/* sizeof of the structs above as compiled into the library (0 model, 1 graph, 2 sigma context, 3 workspace,
 * 4 gc_mlp2, 5 gc_transformer_layer), for binding-side layout checks */
GC_API int gc_sizeof_forward_structs(int32_t which);

#ifdef __cplusplus
}
#endif
#endif  /* GENCAST_B200_H_ */
