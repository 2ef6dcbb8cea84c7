// Fused edge update + aggregation of a bipartite interaction network whose receivers have exactly three incoming
// edges each, stored receiver-major (GenCast's mesh2grid decoder: every grid node receives from the three vertices
// of the mesh triangle that contains it, common/grid_mesh_connectivity.py:104, :125-131).
//
// Reference operators replaced, end to end and without any [E, L] tensor in HBM:
//   EdgeWrapper / MLPWithNormConditioning of the edge update        common/typed_graph_net.py:134-159, :295-305,
//                                                                    common/mlp.py:115-147
//   aggregate_fn = jraph.segment_sum over the receivers             common/typed_graph_net.py:161-195,
//                                                                    common/deep_typed_graph_net.py:396-410
// i.e. for every receiver v with edges e = 3v, 3v+1, 3v+2
//   h_e   = act( base[e mod period] + P_s[senders[e]] + P_r[receivers[e]] )      first MLP layer, split by operand
//   y_e   = h_e W2^T + b2                                                        second MLP layer (tensor cores)
//   out_v = (1 + s) * sum_e LayerNorm(y_e) + 3 o                                 LN + conditional affine + segment sum
//
// One persistent CTA per SM walks tiles of 40 receivers = 120 edges.  A tile is one tcgen05 accumulator of
// 128 rows x L columns in tensor memory (the whole row of every edge: LayerNorm needs it): TMEM lane quarter q
// holds the 30 edges of receivers 10 q .. 10 q + 9 (lanes 30, 31 of each quarter idle), so the three rows of a
// receiver always belong to one epilogue warp.
//   warp 0      TMA producer of W2 k-blocks (static weights: never waits for the predecessor grid)
//   warp 1      MMA issuer (M = 128, N = min(L, 256) per instruction, L / 256 instructions per K step), TMEM owner
//   warps 2-9   A producers: sum the three operand rows in fp32, activation, bf16, write the k-block into a
//               128B-swizzled K-major stage (the layout TMA would have produced)
//   warp 18     (FAST) TMA loader of the two operands that are contiguous per tile, see below
//   warps 10-17 epilogue, two per TMEM lane quarter (half of the columns each): pass 1 row statistics over
//               acc + b2, pass 2 normalise, sum the three rows of each receiver with two lane shuffles per element,
//               conditional affine, 64-byte row segments stored by the first lane of each receiver
// The A ring and the W ring (3 x 32 KB) run ahead of the tensor core across tile boundaries, so the gathers of tile
// t+1 proceed under the epilogue of tile t (the accumulator itself cannot be double buffered: 128 x 512 fp32 is all of
// tensor memory).
//
// Operand feed.  With all three operands fetched through registers (12 x 16 B per thread and k-block, consumed before
// the next k-block's loads are issued) a k-block costs one full load latency, ~2 000 clk under load, and 8 of them per
// tile are more than the tile's MMA + epilogue time: 48 KB in flight per SM is what bounded the kernel.  In the FAST
// instantiation (receivers' own rows: idx_r == null, and period a multiple of 30 edges) the `base` rows of a tile
// quarter (30 consecutive table rows) and the receiver rows of the tile (40 consecutive rows) arrive by TMA straight
// into the stage - base rows in the final swizzled position, where the producers finish them in place - up to four
// k-blocks ahead, and only the sender rows are gathered through registers, two k-blocks ahead of their use.
#include "common.cuh"
#include "sm100.cuh"

namespace gc {

long long* g_edge_fused_trace = nullptr;   // debug: clock stamps of CTA 0 (null in normal use)

namespace {

// Debug timeline: CTA 0 records clock64() at pipeline events into trace[role * 512 + index].
#define GC_ETR(role, index)                                                                                       \
  do {                                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && (index) < 512) p.trace[(role) * 512 + (index)] = clock64();      \
  } while (0)

constexpr int EF_THREADS = 608;
constexpr int EF_PRODUCER_WARPS = 8;
constexpr int EF_EPI_WARPS = 8;
constexpr int EF_RECV_PER_TILE = 40;
constexpr int EF_A_STAGES = 5;
constexpr int EF_GR_BYTES = EF_RECV_PER_TILE * 128;     // FAST: receiver rows of one k-block (40 x 64 bf16)
constexpr int EF_W_STAGES = 3;
constexpr int EF_A_STAGE_BYTES = 128 * 64 * 2;
constexpr int EF_PATCH_STRIDE = 36;                     // floats per row of the transposition patch (32 + pad, 16 B aligned)
constexpr int EF_PATCH_BYTES = 32 * EF_PATCH_STRIDE * 4;
constexpr float EF_LN_EPS = 1e-6f;

// MODE 0: all operands through registers; 1 (FAST): base / receiver rows by TMA (degree-3 tiles of 40 receivers);
// 2 (ROWS): tiles of 128 consecutive edges of one member, base rows by TMA, one gather, plain row output (gc_edge_mlp_rows)
template <int L, int MODE = 0>
struct EFCfg {
  static constexpr bool FAST = MODE != 0;
  static constexpr int NI = L < 256 ? L : 256;          // columns per MMA instruction
  static constexpr int NH = L / NI;                     // instructions per K step
  static constexpr int KB = L / 64;                     // k-blocks (the hidden layer is L wide)
  static constexpr int W_STAGE_BYTES = NI * 64 * 2;
  static constexpr int A_STAGES = EF_A_STAGES;
  static constexpr int A_STRIDE = EF_A_STAGE_BYTES + (MODE == 1 ? EF_GR_BYTES : 0);   // multiple of 1024 either way
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = A_STAGES * A_STRIDE;
  static constexpr int PATCH_OFF = W_OFF + EF_W_STAGES * W_STAGE_BYTES;
  // MODE 0: the transposition patches of linear_ln_cond_kernel; MODE 2: one 4 KB TMA-store staging buffer per epilogue warp
  static constexpr int VEC_OFF = PATCH_OFF + (MODE == 0 ? EF_EPI_WARPS * EF_PATCH_BYTES : MODE == 2 ? EF_EPI_WARPS * 4096 : 0);   // b2 [L] | scale [L] | offset [L] floats
  static constexpr int STAT_OFF = VEC_OFF + 3 * L * 4;                            // [2 halves][128 rows] float2
  static constexpr int BAR_OFF = STAT_OFF + 2 * 128 * 8;
  static constexpr int SMEM = BAR_OFF + 256 + 1024;
  static constexpr uint32_t TMEM_COLS = L;              // 128 / 256 / 512: powers of two
};

struct EdgeFusedParams {
  const __nv_bfloat16* base; int64_t ld_base; int64_t period;
  const __nv_bfloat16* gs; const int32_t* idx_s; int64_t ld_gs;
  const __nv_bfloat16* gr; const int32_t* idx_r; int64_t ld_gr;
  int act;
  const float* b2;
  const float* scale_offset;     // [2 L] = (1 + s | o), or null
  int do_ln;
  void* out; int out_dtype; int64_t ldo;
  int64_t num_receivers;
  int num_tiles;
  int members, tiles_per_member;   // tile order: see tile_of()
  int64_t num_rows;                // ROWS mode: members * period edges
  float* row_stats;                // ROWS mode, optional: [num_rows][2 halves] (sum, sum of squares) of y over the half's columns
  long long* trace;
};

// Ensemble members evaluated together share the `base` table (period = one member's edges).  Walking the tiles member by
// member streams the table from HBM once per member; interleaving the members (consecutive tile slots = the same base
// rows for member 0, 1, ...) makes all but the first read of a base row an L2 hit.  Falls back to the plain order when
// the member blocks are not a whole number of tiles.
__device__ __forceinline__ int tile_of(const EdgeFusedParams& p, int slot) {
  if (p.members <= 1) return slot;
  return (slot % p.members) * p.tiles_per_member + slot / p.members;
}

template <int L, int MODE>
__global__ void __launch_bounds__(EF_THREADS, 1)
edge_mlp_sum3_kernel(const __grid_constant__ CUtensorMap w_map, const __grid_constant__ CUtensorMap base_map,
                     const __grid_constant__ CUtensorMap gr_map, const EdgeFusedParams p) {
  using namespace sm100;
  using C = EFCfg<L, MODE>;
  constexpr bool FAST = MODE != 0, ROWS = MODE == 2;
  constexpr int A_STAGES = C::A_STAGES;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_smem = smem_base + C::A_OFF;
  const uint32_t w_smem = smem_base + C::W_OFF;
  const uint32_t bars = smem_base + C::BAR_OFF;
  float* vec_s = reinterpret_cast<float*>(smem_gen + C::VEC_OFF);
  float2* stat_s = reinterpret_cast<float2*>(smem_gen + C::STAT_OFF);
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (A_STAGES + s); };
  auto w_full = [&](int s) { return bars + 8u * (2 * A_STAGES + s); };
  auto w_empty = [&](int s) { return bars + 8u * (2 * A_STAGES + EF_W_STAGES + s); };
  const uint32_t acc_full = bars + 8u * (2 * A_STAGES + 2 * EF_W_STAGES);
  const uint32_t acc_empty = acc_full + 8u;
  const uint32_t tmem_ptr_smem = acc_full + 16u;
  auto raw_full = [&](int s) { return acc_full + 24u + 8u * s; };      // FAST: TMA-fed operands of stage s have landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&w_map);
    if (FAST) prefetch_tensormap(&base_map);
    if (MODE != 0) prefetch_tensormap(&gr_map);     // MODE 2: this slot carries the tensor map of the output rows
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(a_full(s), EF_PRODUCER_WARPS);
      mbar_init(a_empty(s), 1);
      mbar_init(raw_full(s), 1);
    }
    for (int s = 0; s < EF_W_STAGES; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EF_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // warps 0 and 1 run warp-uniform code and elect one lane for the TMA / tcgen05 instructions (operands stay in
  // uniform registers; see gemm_tcgen05.cu)
  if (warp == 0) {
    // ---------------- W2 k-blocks: the same L x L weight for every tile, streamed from L2
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < C::KB; ++kb) {
        for (int h = 0; h < C::NH; ++h) {
          mbar_wait(w_empty(stage), phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(w_full(stage), C::W_STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < C::NI / 128; ++j)
              tma_load_2d(w_smem + stage * C::W_STAGE_BYTES + j * (128 * 128), &w_map, w_full(stage), kb * 64, h * C::NI + j * 128);
          }
          __syncwarp();
          if (++stage == EF_W_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    constexpr uint32_t idesc = idesc_bf16_f32(128, C::NI, 0, 0);
    int sa = 0, sw = 0;
    uint32_t pa = 0, pw = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      GC_ETR(0, 3 * lt);
      mbar_wait(acc_empty, (static_cast<uint32_t>(lt) & 1u) ^ 1u);      // the epilogue has drained the accumulator
      GC_ETR(0, 3 * lt + 1);
      tc_fence_after();
      for (int kb = 0; kb < C::KB; ++kb) {
        mbar_wait(a_full(sa), pa);
        GC_ETR(1, lt * C::KB + kb);
        if (FAST) fence_proxy_async_smem();        // the producers' generic-proxy stores -> the tensor core's async-proxy reads
        const uint64_t da = desc_kmajor_sw128(a_smem + sa * C::A_STRIDE);
        for (int h = 0; h < C::NH; ++h) {
          mbar_wait(w_full(sw), pw);
          tc_fence_after();
          const uint64_t dw = desc_kmajor_sw128(w_smem + sw * C::W_STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + h * C::NI, da + 2u * k, dw + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(w_empty(sw));
          }
          __syncwarp();
          if (++sw == EF_W_STAGES) { sw = 0; pw ^= 1u; }
        }
        if (elect_one()) umma_commit(a_empty(sa));
        __syncwarp();
        if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
      GC_ETR(0, 3 * lt + 2);
    }
  } else if (warp < 2 + EF_PRODUCER_WARPS) {
    // ---------------- A producers: thread = (16-byte unit u of the k-block row, rows i, 32 + i, 64 + i, 96 + i)
    const int pt = threadIdx.x - 64;               // 0 .. 255
    const int u = pt & 7;
    const int i = pt >> 3;                         // lane of the TMEM quarter this row lands in (30, 31: padding)
    pdl_wait();
    int sa = 0;
    uint32_t pa = 0;
    if constexpr (FAST) {
      // base and receiver rows are in the stage (TMA); sender rows come through registers, two k-blocks ahead
      const int grow = i / 3;                      // receiver of this row inside its quarter
      auto setup = [&](int slot, const __nv_bfloat16* (&ps)[4], uint32_t& vmask) {
        vmask = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) ps[j] = p.gs;
        if (slot < p.num_tiles) {
          const int tile = tile_of(p, slot);
          if constexpr (ROWS) {
            // row 32 j + i of the tile = edge (tile % tiles_per_member) * 128 + 32 j + i of member tile / tiles_per_member
            const int64_t member = tile / p.tiles_per_member;
            const int64_t local0 = static_cast<int64_t>(tile % p.tiles_per_member) * 128 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int64_t local = local0 + 32 * j;
              if (local < p.period) {
                const __nv_bfloat16* row = p.gs + static_cast<int64_t>(__ldg(p.idx_s + member * p.period + local)) * p.ld_gs;
                ps[j] = row + u * 8;
                vmask |= 1u << j;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int64_t recv = static_cast<int64_t>(tile) * EF_RECV_PER_TILE + 10 * j + grow;
              if (i < 30 && recv < p.num_receivers) {
                const int64_t e = static_cast<int64_t>(tile) * (3 * EF_RECV_PER_TILE) + 30 * j + i;
                ps[j] = p.gs + static_cast<int64_t>(__ldg(p.idx_s + e)) * p.ld_gs + u * 8;
                vmask |= 1u << j;
              }
            }
          }
        }
      };
      auto issue = [&](const __nv_bfloat16* const (&ps)[4], uint32_t vmask, int kb, uint4 (&x)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (vmask & (1u << j)) x[j] = __ldg(reinterpret_cast<const uint4*>(ps[j] + kb * 64));
      };
      const __nv_bfloat16* ps[4];
      uint32_t vm;
      uint4 xs[2][4];
      setup(blockIdx.x, ps, vm);
      issue(ps, vm, 0, xs[0]);
      issue(ps, vm, 1, xs[1]);
      for (int slot = blockIdx.x; slot < p.num_tiles; slot += gridDim.x) {
        const __nv_bfloat16* pn[4];
        uint32_t vn;
        setup(slot + static_cast<int>(gridDim.x), pn, vn);
        // unrolled by two only (the register slot alternates): eight copies of this body are 100 KB of SASS and the
        // warps then stall on instruction fetch ('no_inst' in the source-level profile)
#pragma unroll 2
        for (int kb = 0; kb < C::KB; ++kb) {
          uint4 (&cur)[4] = xs[kb & 1];
          if (threadIdx.x == 64) GC_ETR(2, 3 * ((slot / gridDim.x) * C::KB + kb));
          mbar_wait(raw_full(sa), pa);
          if (threadIdx.x == 64) GC_ETR(2, 3 * ((slot / gridDim.x) * C::KB + kb) + 1);
          const uint32_t stage = a_smem + sa * C::A_STRIDE;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t row = static_cast<uint32_t>(32 * j + i);
            const uint32_t addr = stage + row * 128u + ((static_cast<uint32_t>(u) ^ (row & 7u)) << 4);
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (vm & (1u << j)) {
              uint4 xb, xr = make_uint4(0u, 0u, 0u, 0u);      // ROWS: no receiver operand (bf16 zeros)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xb.x), "=r"(xb.y), "=r"(xb.z), "=r"(xb.w) : "r"(addr));
              if constexpr (!ROWS) {
                const uint32_t rr = static_cast<uint32_t>(10 * j + grow);
                const uint32_t raddr = stage + EF_A_STAGE_BYTES + rr * 128u + ((static_cast<uint32_t>(u) ^ (rr & 7u)) << 4);
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xr.x), "=r"(xr.y), "=r"(xr.z), "=r"(xr.w) : "r"(raddr));
              }
              const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&xb);
              const __nv_bfloat162* hs = reinterpret_cast<const __nv_bfloat162*>(&cur[j]);
              const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&xr);
              __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 fb = __bfloat1622float2(hb[k]), fs = __bfloat1622float2(hs[k]), fr = __bfloat1622float2(hr[k]);
                float v0 = fb.x + fs.x, v1 = fb.y + fs.y;
                if constexpr (!ROWS) { v0 += fr.x; v1 += fr.y; }
                v0 = apply_act<true>(v0, p.act);
                v1 = apply_act<true>(v1, p.act);
                ho[k] = __floats2bfloat162_rn(v0, v1);
              }
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
          // No proxy fence here: fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR would wait for
          // this thread's prefetched sender rows (measured: 35 % of all stall samples, the prefetch bought nothing).  The
          // MMA warp, which has no loads in flight, fences after its wait on a_full (release / acquire through the barrier).
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full(sa));
          if (threadIdx.x == 64) GC_ETR(2, 3 * ((slot / gridDim.x) * C::KB + kb) + 2);
          if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
          // refill the register slot just consumed with the sender rows of k-block kb + 2 (of the next tile at the end)
          if (kb + 2 < C::KB) issue(ps, vm, kb + 2, xs[kb & 1]);
          else issue(pn, vn, kb + 2 - C::KB, xs[kb & 1]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ps[j] = pn[j];
        vm = vn;
      }
    } else {
      for (int slot = blockIdx.x; slot < p.num_tiles; slot += gridDim.x) {
        const int tile = tile_of(p, slot);
        // rows of this thread: quarter j holds receivers 10 j .. 10 j + 9 of the tile
        const __nv_bfloat16* pb[4];
        const __nv_bfloat16* ps[4];
        const __nv_bfloat16* pr[4];
        bool valid[4];
  #pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t recv = static_cast<int64_t>(tile) * EF_RECV_PER_TILE + 10 * j + i / 3;
          valid[j] = i < 30 && recv < p.num_receivers;
          const int64_t e = static_cast<int64_t>(tile) * (3 * EF_RECV_PER_TILE) + 30 * j + i;
          if (valid[j]) {
            pb[j] = p.base + (e % p.period) * p.ld_base + u * 8;
            ps[j] = p.gs + static_cast<int64_t>(__ldg(p.idx_s + e)) * p.ld_gs + u * 8;
            pr[j] = p.gr + (p.idx_r != nullptr ? static_cast<int64_t>(__ldg(p.idx_r + e)) : e / 3) * p.ld_gr + u * 8;
          } else {
            pb[j] = ps[j] = pr[j] = p.base;
          }
        }
        for (int kb = 0; kb < C::KB; ++kb) {
          uint4 xb[4], xs[4], xr[4];
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (valid[j]) {
              xb[j] = __ldg(reinterpret_cast<const uint4*>(pb[j] + kb * 64));
              xs[j] = __ldg(reinterpret_cast<const uint4*>(ps[j] + kb * 64));
              xr[j] = __ldg(reinterpret_cast<const uint4*>(pr[j] + kb * 64));
            }
          }
          if (lane == 0) mbar_wait(a_empty(sa), pa ^ 1u);
          __syncwarp();
          const uint32_t stage = a_smem + sa * C::A_STRIDE;
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (valid[j]) {
              const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&xb[j]);
              const __nv_bfloat162* hs = reinterpret_cast<const __nv_bfloat162*>(&xs[j]);
              const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&xr[j]);
              __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
  #pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 fb = __bfloat1622float2(hb[k]), fs = __bfloat1622float2(hs[k]), fr = __bfloat1622float2(hr[k]);
                float v0 = fb.x + fs.x + fr.x, v1 = fb.y + fs.y + fr.y;
                v0 = apply_act<true>(v0, p.act);
                v1 = apply_act<true>(v1, p.act);
                ho[k] = __floats2bfloat162_rn(v0, v1);
              }
            }
            const uint32_t row = static_cast<uint32_t>(32 * j + i);
            const uint32_t addr = stage + row * 128u + ((static_cast<uint32_t>(u) ^ (row & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
          fence_proxy_async_smem();                // generic-proxy stores -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full(sa));
          if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
        }
      }
    }
  } else if (warp == 2 + EF_PRODUCER_WARPS + EF_EPI_WARPS) {
    // ---------------- (FAST) TMA loader: per k-block the base rows of the four tile quarters into their final place
    // (quarter j = rows 32 j .. 32 j + 29 of the stage) and the 40 receiver rows behind the stage
    if constexpr (FAST) {
      pdl_wait();                                  // the receiver rows are the predecessor's output
      int sa = 0;
      uint32_t pa = 0;
      for (int slot = blockIdx.x; slot < p.num_tiles; slot += gridDim.x) {
        const int tile = tile_of(p, slot);
        int b_row[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          b_row[j] = static_cast<int>((static_cast<int64_t>(tile) * (3 * EF_RECV_PER_TILE) + 30 * j) % p.period);
        const int r_row = ROWS ? (tile % p.tiles_per_member) * 128 : tile * EF_RECV_PER_TILE;
        for (int kb = 0; kb < C::KB; ++kb) {
          mbar_wait(a_empty(sa), pa ^ 1u);
          if (elect_one()) {
            const uint32_t stage = a_smem + sa * C::A_STRIDE;
            if constexpr (ROWS) {
              // the tile's 128 base rows in one box (rows past the member's last edge are zero-filled)
              mbar_arrive_expect_tx(raw_full(sa), EF_A_STAGE_BYTES);
              tma_load_2d(stage, &base_map, raw_full(sa), kb * 64, r_row);
            } else {
              mbar_arrive_expect_tx(raw_full(sa), 4 * 30 * 128 + EF_GR_BYTES);
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_2d(stage + j * 4096, &base_map, raw_full(sa), kb * 64, b_row[j]);
              tma_load_2d(stage + EF_A_STAGE_BYTES, &gr_map, raw_full(sa), kb * 64, r_row);
            }
          }
          __syncwarp();
          if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
        }
      }
    }
  } else {
    // ---------------- epilogue: LayerNorm over the whole row, sum of the three rows of each receiver, affine, store
    const int ew = warp - (2 + EF_PRODUCER_WARPS);       // 0 .. 7
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int half = ew >> 2;                            // which half of the columns (warps 10-13: q = 2,3,0,1; 14-17 again)
    constexpr int CH = L / 2;                            // columns per warp
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * CH;
    const int et = threadIdx.x - 32 * (2 + EF_PRODUCER_WARPS);     // 0 .. 255
    pdl_wait();
    for (int c = et; c < 3 * L; c += 32 * EF_EPI_WARPS) {
      float v;
      if (c < L) v = p.b2 != nullptr ? __ldg(p.b2 + c) : 0.0f;
      else if (p.scale_offset != nullptr) v = __ldg(p.scale_offset + (c - L));
      else v = c < 2 * L ? 1.0f : 0.0f;
      vec_s[c] = v;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float* b2s = vec_s + half * CH;
    const float* scs = vec_s + L + half * CH;
    const float* ofs = vec_s + 2 * L + half * CH;
    const float inv_n = 1.0f / static_cast<float>(L);
    int lt = 0;
    if constexpr (ROWS) {
      // ---- plain row output: y = acc + b2 as bf16, one pass.  Each lane owns a row; 64-column groups are staged in a
      // 128B-swizzled 4 KB buffer per warp and leave by TMA (a lane-per-row st.global touches 32 lines per instruction:
      // 13 800 clk per tile in the first version, most of it LSU time).  The last, partial tile of a member stores
      // directly (a TMA box would run into the next member's rows).
      const uint32_t stage_buf = smem_base + C::PATCH_OFF + static_cast<uint32_t>(ew) * 4096u;
      const uint32_t sw = static_cast<uint32_t>(lane & 7);
      for (int slot = blockIdx.x; slot < p.num_tiles; slot += gridDim.x, ++lt) {
        const int tile = tile_of(p, slot);
        const int64_t member = tile / p.tiles_per_member;
        const int64_t local0 = static_cast<int64_t>(tile % p.tiles_per_member) * 128;
        const int64_t local = local0 + q * 32 + lane;
        const bool whole = local0 + 128 <= p.period;
        const bool store = local < p.period;
        const int64_t grow0 = member * p.period + local0 + q * 32;
        __nv_bfloat16* dst_row = reinterpret_cast<__nv_bfloat16*>(p.out) + (grow0 + lane) * p.ldo + half * CH;
        if (et == 0) GC_ETR(3, 4 * lt);
        mbar_wait(acc_full, static_cast<uint32_t>(lt) & 1u);
        if (et == 0) GC_ETR(3, 4 * lt + 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr, r);
        float rs = 0.0f, rss = 0.0f;                   // this row's sum / sum of squares of y over this warp's columns
#pragma unroll 1
        for (int c = 0; c < CH; c += 32) {
          float v[32];
          tc_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
          if (c + 32 < CH) {
            tmem_ld_32x32b_x32(taddr + c + 32, r);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
          }
          const int hc = (c >> 5) & 1;                 // which half of the 64-column group
          if (whole && hc == 0) {
            if (lane == 0) bulk_wait_group_read<0>();  // the previous group's store has finished reading the buffer
            __syncwarp();
          }
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(b2s + c + k);
            const float4 b1 = *reinterpret_cast<const float4*>(b2s + c + k + 4);
            const float y0 = v[k] + b0.x, y1 = v[k + 1] + b0.y, y2 = v[k + 2] + b0.z, y3 = v[k + 3] + b0.w;
            const float y4 = v[k + 4] + b1.x, y5 = v[k + 5] + b1.y, y6 = v[k + 6] + b1.z, y7 = v[k + 7] + b1.w;
            if (p.row_stats != nullptr) {              // warp-uniform; statistics of the fp32 values, before rounding
              rs += ((y0 + y1) + (y2 + y3)) + ((y4 + y5) + (y6 + y7));
              rss = fmaf(y0, y0, rss); rss = fmaf(y1, y1, rss); rss = fmaf(y2, y2, rss); rss = fmaf(y3, y3, rss);
              rss = fmaf(y4, y4, rss); rss = fmaf(y5, y5, rss); rss = fmaf(y6, y6, rss); rss = fmaf(y7, y7, rss);
            }
            uint4 o;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
            h[0] = __floats2bfloat162_rn(y0, y1);
            h[1] = __floats2bfloat162_rn(y2, y3);
            h[2] = __floats2bfloat162_rn(y4, y5);
            h[3] = __floats2bfloat162_rn(y6, y7);
            if (whole) {
              const uint32_t unit = static_cast<uint32_t>(hc * 4 + (k >> 3)) ^ sw;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_buf + static_cast<uint32_t>(lane) * 128u + unit * 16u),
                           "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
            } else if (store) {
              *reinterpret_cast<uint4*>(dst_row + c + k) = o;
            }
          }
          if (whole && hc == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&gr_map, stage_buf, half * CH + (c & ~63), static_cast<int>(grow0));
              bulk_commit_group();
            }
          }
        }
        if (p.row_stats != nullptr && store)
          *reinterpret_cast<float2*>(p.row_stats + (grow0 + lane) * 4 + half * 2) = make_float2(rs, rss);
        if (et == 0) GC_ETR(3, 4 * lt + 3);
      }
      if (lane == 0) bulk_wait_group_all();
    } else
    for (int slot = blockIdx.x; slot < p.num_tiles; slot += gridDim.x, ++lt) {
      const int tile = tile_of(p, slot);
      if (et == 0) GC_ETR(3, 4 * lt);
      mbar_wait(acc_full, static_cast<uint32_t>(lt) & 1u);
      if (et == 0) GC_ETR(3, 4 * lt + 1);
      tc_fence_after();
      // ---- pass 1: row statistics of y = acc + b2 over this warp's half of the columns
      float s = 0.0f, ss = 0.0f;
      {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr, r);
#pragma unroll 1
        for (int c = 0; c < CH; c += 32) {
          float v[32];
          tc_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
          if (c + 32 < CH) tmem_ld_32x32b_x32(taddr + c + 32, r);
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 b = *reinterpret_cast<const float4*>(b2s + c + k);
            const float y0 = v[k] + b.x, y1 = v[k + 1] + b.y, y2 = v[k + 2] + b.z, y3 = v[k + 3] + b.w;
            s += (y0 + y1) + (y2 + y3);
            ss = fmaf(y0, y0, ss); ss = fmaf(y1, y1, ss); ss = fmaf(y2, y2, ss); ss = fmaf(y3, y3, ss);
          }
        }
      }
      stat_s[half * 128 + q * 32 + lane] = make_float2(s, ss);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) GC_ETR(3, 4 * lt + 2);
      float mean = 0.0f, rstd = 1.0f;
      if (p.do_ln) {
        const float2 o = stat_s[(half ^ 1) * 128 + q * 32 + lane];
        mean = (s + o.x) * inv_n;
        rstd = rsqrtf(fmaxf((ss + o.y) * inv_n - mean * mean, 0.0f) + EF_LN_EPS);
      }
      // ---- pass 2: normalise, sum the three rows of each receiver across lanes (rows 3 v .. 3 v + 2 sit in lanes
      // 3 v .. 3 v + 2 of this warp), affine, store.  Two shuffles per element instead of a trip through a shared-memory
      // patch: every element is independent, so the warp is never latency-bound on a store -> sync -> load chain.
      const bool leader = lane < 30 && lane % 3 == 0;
      const int64_t recv = static_cast<int64_t>(tile) * EF_RECV_PER_TILE + 10 * q + lane / 3;
      const bool store = leader && recv < p.num_receivers;
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr, r);
#pragma unroll 1
      for (int c = 0; c < CH; c += 32) {
        float v[32];
        if (et == 0 && lt == 2) GC_ETR(4, 4 * (c >> 5));
        tc_wait_ld();
        if (et == 0 && lt == 2) GC_ETR(4, 4 * (c >> 5) + 1);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
        if (c + 32 < CH) {
          tmem_ld_32x32b_x32(taddr + c + 32, r);
        } else {
          // last read of the accumulator: hand it back to the tensor core
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 b = *reinterpret_cast<const float4*>(b2s + c + k);
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = (v[k + j] + bb[j] - mean) * rstd;
            const float x1 = __shfl_down_sync(0xffffffffu, x, 1);
            const float x2 = __shfl_down_sync(0xffffffffu, x, 2);
            v[k + j] = (x + x1) + x2;
          }
        }
        if (et == 0 && lt == 2) GC_ETR(4, 4 * (c >> 5) + 2);
        if (store) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(scs + c + k);
            const float4 of = *reinterpret_cast<const float4*>(ofs + c + k);
            v[k] = fmaf(v[k], sc.x, 3.0f * of.x); v[k + 1] = fmaf(v[k + 1], sc.y, 3.0f * of.y);
            v[k + 2] = fmaf(v[k + 2], sc.z, 3.0f * of.z); v[k + 3] = fmaf(v[k + 3], sc.w, 3.0f * of.w);
          }
          const int64_t off = recv * p.ldo + half * CH + c;
          if (p.out_dtype == GC_BF16) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              uint4 o;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[k + 2 * j], v[k + 2 * j + 1]);
              dst[k >> 3] = o;
            }
          } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
#pragma unroll
            for (int k = 0; k < 32; k += 4) dst[k >> 2] = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
          }
        }
        if (et == 0 && lt == 2) GC_ETR(4, 4 * (c >> 5) + 3);
      }
      if (et == 0) GC_ETR(3, 4 * lt + 3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// Column-split CTA pair for gc_edge_mlp_sum3 (FAST preconditions).  In the single-CTA kernel above the accumulator of a
// tile is 128 rows x L = 512 fp32 columns, i.e. all of tensor memory: the epilogue (two passes over it at the TMEM read
// rate of 64 B/clk/SM = 2 x 4 096 clk, plus two shuffles per element) and the tile's MMAs cannot overlap, and the
// tensor pipe is busy a quarter of the time.  Here a cluster of two CTAs owns a tile: both hold the same hidden-layer
// tile A (128 x L bf16, k-block by k-block) and CTA r computes the output columns [r L/2, (r+1) L/2), so an accumulator is
// L/2 = 256 columns and two of them fit: the MMAs of tile t+1 run under the epilogue of tile t.
//   * A k-block kb is produced by CTA kb % 2 (its producers gather / add / swish exactly as above) and forwarded to the
//     peer's ring by a shared-memory -> peer-shared-memory bulk copy that completes on the peer's `x_full` barrier.
//     The ring has an even number of stages, so stage s always holds k-blocks of parity s % 2: every barrier is used
//     exactly once per ring cycle and phases are (use index / stages) & 1.  Stages of the CTA's own parity (with room
//     for the TMA-fed receiver rows) and stages filled by the peer are two arrays: [own 0..2 | peer 0..2] in both CTAs.
//   * A CTA tells its peer that a peer-filled stage is free (and its x_full armed) by a remote arrive on `p_empty`.
//   * LayerNorm: each CTA reduces its half of the columns, writes the per-row (sum, sum of squares) into the peer's
//     shared memory, and both add the two partials in rank order (the same bits on both sides).
// Warps (640 threads): 0 W2 TMA, 1 MMA, 2-9 producers, 10-17 epilogue, 18 operand TMA + flow control, 19 forwarder.
// ---------------------------------------------------------------------------------------------------------------
constexpr int EP_THREADS = 640;
constexpr int EP_A_STAGES = 6;                           // 3 filled locally (with receiver rows), 3 by the peer

template <int L>
struct EPCfg {
  static constexpr int NC = L / 2;                      // output columns per CTA
  static constexpr int KB = L / 64;
  static constexpr int OWN_STRIDE = EF_A_STAGE_BYTES + EF_GR_BYTES;     // locally produced: A k-block + receiver rows
  static constexpr int PEER_OFF = (EP_A_STAGES / 2) * OWN_STRIDE;       // stages the peer fills: A k-block only
  static constexpr int W_STAGE_BYTES = NC * 64 * 2;
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = PEER_OFF + (EP_A_STAGES / 2) * EF_A_STAGE_BYTES;
  static constexpr int VEC_OFF = W_OFF + EF_W_STAGES * W_STAGE_BYTES;      // b2 | scale | offset of this CTA's columns
  static constexpr int STAT_OFF = VEC_OFF + 3 * NC * 4;                     // [2 buffers][2 halves][128] float2
  static constexpr int XSTAT_OFF = STAT_OFF + 2 * 2 * 128 * 8;              // [2 buffers][128] float2, written by the peer
  static constexpr int BAR_OFF = XSTAT_OFF + 2 * 128 * 8;
  static constexpr int SMEM = BAR_OFF + 512 + 1024;
  static_assert(KB % 2 == 0 && EP_A_STAGES % 2 == 0, "k-block parity = stage parity");
  static_assert(SMEM <= 232448, "edge_mlp_sum3_pair_kernel: shared memory plan does not fit");
};

__device__ __forceinline__ uint32_t map_to_peer(uint32_t local_smem_addr, uint32_t peer_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(peer_rank));
  return r;
}

template <int L>
__global__ void __launch_bounds__(EP_THREADS, 1)
edge_mlp_sum3_pair_kernel(const __grid_constant__ CUtensorMap w_map, const __grid_constant__ CUtensorMap base_map,
                          const __grid_constant__ CUtensorMap gr_map, const EdgeFusedParams p) {
  using namespace sm100;
  using C = EPCfg<L>;
  constexpr int S = EP_A_STAGES, KB = C::KB, NC = C::NC;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_smem = smem_base + C::A_OFF;
  const uint32_t w_smem = smem_base + C::W_OFF;
  const uint32_t bars = smem_base + C::BAR_OFF;
  float* vec_s = reinterpret_cast<float*>(smem_gen + C::VEC_OFF);
  float2* stat_s = reinterpret_cast<float2*>(smem_gen + C::STAT_OFF);
  float2* xstat_s = reinterpret_cast<float2*>(smem_gen + C::XSTAT_OFF);
  auto a_full = [&](int s) { return bars + 8u * s; };                      // local producers done (8 warps)
  auto x_full = [&](int s) { return bars + 8u * (S + s); };                // peer's copy landed (1 arrive + bytes)
  auto a_empty = [&](int s) { return bars + 8u * (2 * S + s); };           // local MMAs have read the stage
  auto p_empty = [&](int s) { return bars + 8u * (3 * S + s); };           // peer's stage s is free and armed (remote arrive)
  auto raw_full = [&](int s) { return bars + 8u * (4 * S + s); };          // TMA-fed operands of an own stage landed
  auto w_full = [&](int s) { return bars + 8u * (5 * S + s); };
  auto w_empty = [&](int s) { return bars + 8u * (5 * S + EF_W_STAGES + s); };
  auto acc_full = [&](int b) { return bars + 8u * (5 * S + 2 * EF_W_STAGES + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (5 * S + 2 * EF_W_STAGES + 2 + b); };
  auto stat_full = [&](int b) { return bars + 8u * (5 * S + 2 * EF_W_STAGES + 4 + b); };   // peer's row statistics arrived
  const uint32_t tmem_ptr_smem = bars + 8u * (5 * S + 2 * EF_W_STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t peer = rank ^ 1u;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // shared-memory address of ring stage s in this CTA: produced here (s % 2 == rank) or filled by the peer
  auto stage_addr = [&](int s) {
    return (static_cast<uint32_t>(s) & 1u) == rank ? a_smem + static_cast<uint32_t>(s >> 1) * C::OWN_STRIDE
                                                    : a_smem + C::PEER_OFF + static_cast<uint32_t>(s >> 1) * EF_A_STAGE_BYTES;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&w_map);
    prefetch_tensormap(&base_map);
    prefetch_tensormap(&gr_map);
    for (int s = 0; s < S; ++s) {
      mbar_init(a_full(s), EF_PRODUCER_WARPS);
      mbar_init(x_full(s), 1);
      mbar_init(a_empty(s), 1);
      mbar_init(p_empty(s), 1);
      mbar_init(raw_full(s), 1);
    }
    for (int s = 0; s < EF_W_STAGES; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), EF_EPI_WARPS); mbar_init(stat_full(b), 4); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 2 * NC);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // both CTAs' barriers exist before anything remote happens
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ---------------- this CTA's half of W2 (rows rank * NC .. + NC), k-block by k-block
    int stage = 0;
    uint32_t phase = 0;
    for (int slot = pair_id; slot < p.num_tiles; slot += num_pairs) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(w_empty(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(w_full(stage), C::W_STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < NC / 128; ++j)
            tma_load_2d(w_smem + stage * C::W_STAGE_BYTES + j * (128 * 128), &w_map, w_full(stage), kb * 64,
                        static_cast<int>(rank) * NC + j * 128);
        }
        __syncwarp();
        if (++stage == EF_W_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: M = 128, N = NC, accumulator buffer lt & 1
    constexpr uint32_t idesc = idesc_bf16_f32(128, NC, 0, 0);
    int sw = 0;
    uint32_t pw = 0;
    int lt = 0;
    int64_t g = 0;                                // running k-block index: stage g % S, phase (g / S) & 1
    for (int slot = pair_id; slot < p.num_tiles; slot += num_pairs, ++lt) {
      const int b = lt & 1;
      GC_ETR(0, 3 * lt);
      mbar_wait(acc_empty(b), ((static_cast<uint32_t>(lt) >> 1) & 1u) ^ 1u);
      GC_ETR(0, 3 * lt + 1);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int sa = static_cast<int>(g % S);
        const uint32_t pa = static_cast<uint32_t>(g / S) & 1u;
        const bool own = (static_cast<uint32_t>(kb) & 1u) == rank;
        mbar_wait(own ? a_full(sa) : x_full(sa), pa);
        GC_ETR(1, lt * KB + kb);
        if (own) fence_proxy_async_smem();        // local generic-proxy stores -> async-proxy reads (peer copies are async already)
        const uint64_t da = desc_kmajor_sw128(stage_addr(sa));
        mbar_wait(w_full(sw), pw);
        tc_fence_after();
        const uint64_t dw = desc_kmajor_sw128(w_smem + sw * C::W_STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + b * NC, da + 2u * k, dw + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(w_empty(sw));
          umma_commit(a_empty(sa));
        }
        __syncwarp();
        if (++sw == EF_W_STAGES) { sw = 0; pw ^= 1u; }
      }
      if (elect_one()) umma_commit(acc_full(b));
      __syncwarp();
      GC_ETR(0, 3 * lt + 2);
    }
  } else if (warp < 2 + EF_PRODUCER_WARPS) {
    // ---------------- A producers for the k-blocks of this CTA's parity (as in the single-CTA FAST path)
    const int pt = threadIdx.x - 64;
    const int u = pt & 7;
    const int i = pt >> 3;
    const int grow = i / 3;
    pdl_wait();
    auto setup = [&](int slot, const __nv_bfloat16* (&ps)[4], uint32_t& vmask) {
      vmask = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) ps[j] = p.gs;
      if (slot < p.num_tiles) {
        const int tile = tile_of(p, slot);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t recv = static_cast<int64_t>(tile) * EF_RECV_PER_TILE + 10 * j + grow;
          if (i < 30 && recv < p.num_receivers) {
            const int64_t e = static_cast<int64_t>(tile) * (3 * EF_RECV_PER_TILE) + 30 * j + i;
            ps[j] = p.gs + static_cast<int64_t>(__ldg(p.idx_s + e)) * p.ld_gs + u * 8;
            vmask |= 1u << j;
          }
        }
      }
    };
    auto issue = [&](const __nv_bfloat16* const (&ps)[4], uint32_t vmask, int kb, uint4 (&x)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (vmask & (1u << j)) x[j] = __ldg(reinterpret_cast<const uint4*>(ps[j] + kb * 64));
    };
    const __nv_bfloat16* ps[4];
    uint32_t vm;
    uint4 xs[2][4];
    const int kb0 = static_cast<int>(rank);       // own k-blocks: kb0, kb0 + 2, ...
    setup(pair_id, ps, vm);
    issue(ps, vm, kb0, xs[0]);
    issue(ps, vm, kb0 + 2, xs[1]);
    int lt = 0;
    for (int slot = pair_id; slot < p.num_tiles; slot += num_pairs, ++lt) {
      const __nv_bfloat16* pn[4];
      uint32_t vn;
      setup(slot + num_pairs, pn, vn);
#pragma unroll 2
      for (int n = 0; n < KB / 2; ++n) {           // n-th own k-block of the tile
        const int kb = kb0 + 2 * n;
        const int64_t g = static_cast<int64_t>(lt) * KB + kb;
        const int sa = static_cast<int>(g % S);
        const uint32_t pa = static_cast<uint32_t>(g / S) & 1u;
        uint4 (&cur)[4] = xs[n & 1];
        if (threadIdx.x == 64) GC_ETR(2, 3 * (lt * (KB / 2) + n));
        mbar_wait(raw_full(sa), pa);
        if (threadIdx.x == 64) GC_ETR(2, 3 * (lt * (KB / 2) + n) + 1);
        const uint32_t stage = stage_addr(sa);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t row = static_cast<uint32_t>(32 * j + i);
          const uint32_t addr = stage + row * 128u + ((static_cast<uint32_t>(u) ^ (row & 7u)) << 4);
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (vm & (1u << j)) {
            const uint32_t rr = static_cast<uint32_t>(10 * j + grow);
            const uint32_t raddr = stage + EF_A_STAGE_BYTES + rr * 128u + ((static_cast<uint32_t>(u) ^ (rr & 7u)) << 4);
            uint4 xb, xr;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xb.x), "=r"(xb.y), "=r"(xb.z), "=r"(xb.w) : "r"(addr));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xr.x), "=r"(xr.y), "=r"(xr.z), "=r"(xr.w) : "r"(raddr));
            const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&xb);
            const __nv_bfloat162* hs = reinterpret_cast<const __nv_bfloat162*>(&cur[j]);
            const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&xr);
            __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 fb = __bfloat1622float2(hb[k]), fs = __bfloat1622float2(hs[k]), fr = __bfloat1622float2(hr[k]);
              float v0 = fb.x + fs.x + fr.x, v1 = fb.y + fs.y + fr.y;
              v0 = apply_act<true>(v0, p.act);
              v1 = apply_act<true>(v1, p.act);
              ho[k] = __floats2bfloat162_rn(v0, v1);
            }
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(sa));
        if (threadIdx.x == 64) GC_ETR(2, 3 * (lt * (KB / 2) + n) + 2);
        // refill the register slot with the sender rows of the own k-block after next (of the next tile at the end)
        if (n + 2 < KB / 2) issue(ps, vm, kb + 4, xs[n & 1]);
        else issue(pn, vn, kb + 4 - KB, xs[n & 1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) ps[j] = pn[j];
      vm = vn;
    }
  } else if (warp < 2 + EF_PRODUCER_WARPS + EF_EPI_WARPS) {
    // ---------------- epilogue on this CTA's NC columns: statistics (exchanged with the peer), normalise, 3-row sums
    const int ew = warp - (2 + EF_PRODUCER_WARPS);
    const int q = warp & 3;
    const int half = ew >> 2;
    constexpr int CH = NC / 2;                           // columns per warp
    const int et = threadIdx.x - 32 * (2 + EF_PRODUCER_WARPS);
    const int col0 = static_cast<int>(rank) * NC;        // first output column of this CTA
    pdl_wait();
    for (int c = et; c < 3 * NC; c += 32 * EF_EPI_WARPS) {
      const int which = c / NC, cc = c - which * NC;
      float v;
      if (which == 0) v = p.b2 != nullptr ? __ldg(p.b2 + col0 + cc) : 0.0f;
      else if (p.scale_offset != nullptr) v = __ldg(p.scale_offset + (which - 1) * L + col0 + cc);
      else v = which == 1 ? 1.0f : 0.0f;
      vec_s[c] = v;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float* b2s = vec_s + half * CH;
    const float* scs = vec_s + NC + half * CH;
    const float* ofs = vec_s + 2 * NC + half * CH;
    const float inv_n = 1.0f / static_cast<float>(L);
    const uint32_t xstat_peer = map_to_peer(smem_base + C::XSTAT_OFF, peer);
    int lt = 0;
    for (int slot = pair_id; slot < p.num_tiles; slot += num_pairs, ++lt) {
      const int tile = tile_of(p, slot);
      const int b = lt & 1;
      const uint32_t ph = (static_cast<uint32_t>(lt) >> 1) & 1u;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * NC + half * CH;
      const int rrow = q * 32 + lane;
      if (et == 0) GC_ETR(3, 4 * lt);
      mbar_wait(acc_full(b), ph);
      if (et == 0) GC_ETR(3, 4 * lt + 1);
      tc_fence_after();
      // ---- pass 1: row statistics over this warp's columns
      float s = 0.0f, ss = 0.0f;
      {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr, r);
#pragma unroll 1
        for (int c = 0; c < CH; c += 32) {
          float v[32];
          tc_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
          if (c + 32 < CH) tmem_ld_32x32b_x32(taddr + c + 32, r);
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(b2s + c + k);
            const float y0 = v[k] + bb.x, y1 = v[k + 1] + bb.y, y2 = v[k + 2] + bb.z, y3 = v[k + 3] + bb.w;
            s += (y0 + y1) + (y2 + y3);
            ss = fmaf(y0, y0, ss); ss = fmaf(y1, y1, ss); ss = fmaf(y2, y2, ss); ss = fmaf(y3, y3, ss);
          }
        }
      }
      stat_s[(b * 2 + half) * 128 + rrow] = make_float2(s, ss);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // this CTA's partial over its NC columns (halves in fixed order: the same value in both half-warps)
      const float2 h0 = stat_s[(b * 2 + 0) * 128 + rrow], h1 = stat_s[(b * 2 + 1) * 128 + rrow];
      const float own_s = h0.x + h1.x, own_ss = h0.y + h1.y;
      if (half == 0) {
        // hand it to the peer: store into its shared memory, then arrive (release at cluster scope) on its barrier
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(xstat_peer + static_cast<uint32_t>(b * 128 + rrow) * 8u),
                     "f"(own_s), "f"(own_ss) : "memory");
        __syncwarp();
        if (lane == 0) {
          const uint32_t rb = map_to_peer(stat_full(b), peer);
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb) : "memory");
        }
      }
      {
        // wait for the peer's partial (acquire at cluster scope)
        uint32_t ok = 0;
        const long long t0 = clock64();
        while (!ok) {
          asm volatile(
              "{\n\t.reg .pred P;\n\t"
              "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2, %3;\n\t"
              "selp.u32 %0, 1, 0, P;\n\t}"
              : "=r"(ok) : "r"(stat_full(b)), "r"(ph), "r"(0x989680u) : "memory");
          if (!ok && clock64() - t0 > 6000000000LL) __trap();
        }
      }
      if (et == 0) GC_ETR(3, 4 * lt + 2);
      float mean = 0.0f, rstd = 1.0f;
      if (p.do_ln) {
        const float2 o = xstat_s[b * 128 + rrow];
        // rank order, so that both CTAs form the same sums
        const float ts = rank == 0 ? own_s + o.x : o.x + own_s;
        const float tss = rank == 0 ? own_ss + o.y : o.y + own_ss;
        mean = ts * inv_n;
        rstd = rsqrtf(fmaxf(tss * inv_n - mean * mean, 0.0f) + EF_LN_EPS);
      }
      // ---- pass 2: normalise, 3-row sums across lanes, affine, store (as in the single-CTA kernel)
      const bool leader = lane < 30 && lane % 3 == 0;
      const int64_t recv = static_cast<int64_t>(tile) * EF_RECV_PER_TILE + 10 * q + lane / 3;
      const bool store = leader && recv < p.num_receivers;
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr, r);
#pragma unroll 1
      for (int c = 0; c < CH; c += 32) {
        float v[32];
        tc_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
        if (c + 32 < CH) {
          tmem_ld_32x32b_x32(taddr + c + 32, r);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(b));
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 bq = *reinterpret_cast<const float4*>(b2s + c + k);
          const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = (v[k + j] + bb[j] - mean) * rstd;
            const float x1 = __shfl_down_sync(0xffffffffu, x, 1);
            const float x2 = __shfl_down_sync(0xffffffffu, x, 2);
            v[k + j] = (x + x1) + x2;
          }
        }
        if (store) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(scs + c + k);
            const float4 of = *reinterpret_cast<const float4*>(ofs + c + k);
            v[k] = fmaf(v[k], sc.x, 3.0f * of.x); v[k + 1] = fmaf(v[k + 1], sc.y, 3.0f * of.y);
            v[k + 2] = fmaf(v[k + 2], sc.z, 3.0f * of.z); v[k + 3] = fmaf(v[k + 3], sc.w, 3.0f * of.w);
          }
          const int64_t off = recv * p.ldo + col0 + half * CH + c;
          if (p.out_dtype == GC_BF16) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              uint4 o;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[k + 2 * j], v[k + 2 * j + 1]);
              dst[k >> 3] = o;
            }
          } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
#pragma unroll
            for (int k = 0; k < 32; k += 4) dst[k >> 2] = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
          }
        }
      }
      if (et == 0) GC_ETR(3, 4 * lt + 3);
    }
  } else if (warp == 2 + EF_PRODUCER_WARPS + EF_EPI_WARPS) {
    // ---------------- operand TMA + flow control, every k-block in order: once the local stage is free, either request
    // the base / receiver rows of an own k-block, or arm x_full for the peer's copy and tell the peer to go ahead
    pdl_wait();
    int lt = 0;
    for (int slot = pair_id; slot < p.num_tiles; slot += num_pairs, ++lt) {
      const int tile = tile_of(p, slot);
      int b_row[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b_row[j] = static_cast<int>((static_cast<int64_t>(tile) * (3 * EF_RECV_PER_TILE) + 30 * j) % p.period);
      const int r_row = tile * EF_RECV_PER_TILE;
      for (int kb = 0; kb < KB; ++kb) {
        const int64_t g = static_cast<int64_t>(lt) * KB + kb;
        const int sa = static_cast<int>(g % S);
        const uint32_t pa = static_cast<uint32_t>(g / S) & 1u;
        const bool own = (static_cast<uint32_t>(kb) & 1u) == rank;
        mbar_wait(a_empty(sa), pa ^ 1u);
        // An own stage is also read by the forwarder's copy of the previous cycle, whose completion is only visible at the
        // peer.  The peer's go-ahead for THIS cycle (p_empty) comes after its MMAs consumed that copy: wait for it too
        // before the stage is overwritten.
        if (own) mbar_wait(p_empty(sa), pa);
        if (elect_one()) {
          const uint32_t stage = stage_addr(sa);
          if (own) {
            mbar_arrive_expect_tx(raw_full(sa), 4 * 30 * 128 + EF_GR_BYTES);
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(stage + j * 4096, &base_map, raw_full(sa), kb * 64, b_row[j]);
            tma_load_2d(stage + EF_A_STAGE_BYTES, &gr_map, raw_full(sa), kb * 64, r_row);
          } else {
            mbar_arrive_expect_tx(x_full(sa), EF_A_STAGE_BYTES);
            mbar_arrive_cluster(p_empty(sa), peer);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- forwarder: a finished own k-block goes to the same stage of the peer's ring
    int lt = 0;
    for (int slot = pair_id; slot < p.num_tiles; slot += num_pairs, ++lt) {
      for (int kb = static_cast<int>(rank); kb < KB; kb += 2) {
        const int64_t g = static_cast<int64_t>(lt) * KB + kb;
        const int sa = static_cast<int>(g % S);
        const uint32_t pa = static_cast<uint32_t>(g / S) & 1u;
        mbar_wait(a_full(sa), pa);                 // local producers are done with the stage
        if (lane == 0) GC_ETR(5, 2 * (lt * (KB / 2) + (kb >> 1)));
        mbar_wait(p_empty(sa), pa);                // the peer's stage is free and its x_full armed
        if (lane == 0) GC_ETR(5, 2 * (lt * (KB / 2) + (kb >> 1)) + 1);
        fence_proxy_async_smem();                  // producers' generic-proxy stores -> the bulk copy's async-proxy reads
        if (elect_one()) {
          const uint32_t src = stage_addr(sa);
          // in the peer this stage is one of the peer-filled ones: same layout in both CTAs, [own | peer-filled]
          const uint32_t dst = map_to_peer(a_smem + C::PEER_OFF + static_cast<uint32_t>(sa >> 1) * EF_A_STAGE_BYTES, peer);
          const uint32_t bar = map_to_peer(x_full(sa), peer);
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "r"(src), "r"(static_cast<uint32_t>(EF_A_STAGE_BYTES)), "r"(bar) : "memory");
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // the peer may still be copying into / reading from this CTA
  if (warp == 1) tmem_dealloc(tmem_base, 2 * NC);
}

// ---------------------------------------------------------------------------------------------------------------
// Second MLP layer + LayerNorm + conditional affine (+ residual) in one kernel, for the node MLPs:
//     out = LN(A W2^T + b2) * (1 + s) + o (+ residual)            MLPWithNormConditioning, common/mlp.py:115-147,
//                                                                  residual of common/deep_typed_graph_net.py:569-581
// Same whole-row accumulator as above (128 rows x L columns of TMEM per tile, persistent CTA per SM), but the A operand
// is the hidden layer in HBM and arrives by TMA, so there are no producer warps: warp 0 streams A and W2 k-blocks,
// warp 1 issues the MMAs, eight epilogue warps (two per TMEM lane quarter, half of the columns each) compute the row
// statistics in a first pass over the accumulator and normalise / transpose / add the residual / store in a second.
// It replaces a GEMM that writes y, and a LayerNorm kernel that reads y and the residual and writes out: y never
// exists in HBM.
// ---------------------------------------------------------------------------------------------------------------
constexpr int LN_THREADS = 320;

struct LinearLnParams {
  const float* b2;
  const float* scale_offset;
  int do_ln;
  const void* residual; int res_dtype; int64_t ld_res;
  void* out; int out_dtype; int64_t ldo;
  int64_t rows;
  int num_tiles;
};

template <int L>
__global__ void __launch_bounds__(LN_THREADS, 1)
linear_ln_cond_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map, const LinearLnParams p) {
  using namespace sm100;
  using C = EFCfg<L>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_smem = smem_base + C::A_OFF;
  const uint32_t w_smem = smem_base + C::W_OFF;
  const uint32_t bars = smem_base + C::BAR_OFF;
  float* vec_s = reinterpret_cast<float*>(smem_gen + C::VEC_OFF);
  float2* stat_s = reinterpret_cast<float2*>(smem_gen + C::STAT_OFF);
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (EF_A_STAGES + s); };
  auto w_full = [&](int s) { return bars + 8u * (2 * EF_A_STAGES + s); };
  auto w_empty = [&](int s) { return bars + 8u * (2 * EF_A_STAGES + EF_W_STAGES + s); };
  const uint32_t acc_full = bars + 8u * (2 * EF_A_STAGES + 2 * EF_W_STAGES);
  const uint32_t acc_empty = acc_full + 8u;
  const uint32_t tmem_ptr_smem = acc_full + 16u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&a_map);
    prefetch_tensormap(&w_map);
    for (int s = 0; s < EF_A_STAGES; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < EF_W_STAGES; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EF_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ---------------- TMA producer: W2 k-blocks (static weights: the first ring pass is requested before the wait for
    // the predecessor grid) and the A k-blocks of this CTA's tiles
    int sa = 0, sw = 0;
    uint32_t pa = 0, pw = 0;
    auto load_w = [&](int kb, int h) {
      mbar_wait(w_empty(sw), pw ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full(sw), C::W_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < C::NI / 128; ++j)
          tma_load_2d(w_smem + sw * C::W_STAGE_BYTES + j * (128 * 128), &w_map, w_full(sw), kb * 64, h * C::NI + j * 128);
      }
      __syncwarp();
      if (++sw == EF_W_STAGES) { sw = 0; pw ^= 1u; }
    };
    // W stages needed before the first A k-block can be consumed: (kb, h) pairs in issue order
    int pre = 0;
    const int first_total = C::KB * C::NH;
    if (static_cast<int>(blockIdx.x) < p.num_tiles)
      for (; pre < EF_W_STAGES && pre < first_total; ++pre) load_w(pre / C::NH, pre % C::NH);
    pdl_wait();
    bool first = true;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < C::KB; ++kb) {
        mbar_wait(a_empty(sa), pa ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full(sa), EF_A_STAGE_BYTES);
          tma_load_2d(a_smem + sa * EF_A_STAGE_BYTES, &a_map, a_full(sa), kb * 64, tile * 128);
        }
        __syncwarp();
        if (++sa == EF_A_STAGES) { sa = 0; pa ^= 1u; }
        for (int h = 0; h < C::NH; ++h) {
          if (first && kb * C::NH + h < pre) continue;        // already requested above
          load_w(kb, h);
        }
      }
      first = false;
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    constexpr uint32_t idesc = idesc_bf16_f32(128, C::NI, 0, 0);
    int sa = 0, sw = 0;
    uint32_t pa = 0, pw = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      mbar_wait(acc_empty, (static_cast<uint32_t>(lt) & 1u) ^ 1u);
      tc_fence_after();
      for (int kb = 0; kb < C::KB; ++kb) {
        mbar_wait(a_full(sa), pa);
        const uint64_t da = desc_kmajor_sw128(a_smem + sa * EF_A_STAGE_BYTES);
        for (int h = 0; h < C::NH; ++h) {
          mbar_wait(w_full(sw), pw);
          tc_fence_after();
          const uint64_t dw = desc_kmajor_sw128(w_smem + sw * C::W_STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + h * C::NI, da + 2u * k, dw + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(w_empty(sw));
          }
          __syncwarp();
          if (++sw == EF_W_STAGES) { sw = 0; pw ^= 1u; }
        }
        if (elect_one()) umma_commit(a_empty(sa));
        __syncwarp();
        if (++sa == EF_A_STAGES) { sa = 0; pa ^= 1u; }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue
    const int ew = warp - 2;                             // 0 .. 7
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int half = ew >> 2;
    constexpr int CH = L / 2;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * CH;
    float* patch = reinterpret_cast<float*>(smem_gen + C::PATCH_OFF + ew * EF_PATCH_BYTES);
    const int et = threadIdx.x - 64;                     // 0 .. 255
    pdl_wait();
    for (int c = et; c < 3 * L; c += 32 * EF_EPI_WARPS) {
      float v;
      if (c < L) v = p.b2 != nullptr ? __ldg(p.b2 + c) : 0.0f;
      else if (p.scale_offset != nullptr) v = __ldg(p.scale_offset + (c - L));
      else v = c < 2 * L ? 1.0f : 0.0f;
      vec_s[c] = v;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float* b2s = vec_s + half * CH;
    const float* scs = vec_s + L + half * CH;
    const float* ofs = vec_s + 2 * L + half * CH;
    const float inv_n = 1.0f / static_cast<float>(L);
    const int cp = lane & 15;                            // column pair of a 32-column chunk
    const int rsel = lane >> 4;                          // row parity handled in the store loop
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      mbar_wait(acc_full, static_cast<uint32_t>(lt) & 1u);
      tc_fence_after();
      float s = 0.0f, ss = 0.0f;
      {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr, r);
#pragma unroll 1
        for (int c = 0; c < CH; c += 32) {
          float v[32];
          tc_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
          if (c + 32 < CH) tmem_ld_32x32b_x32(taddr + c + 32, r);
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 b = *reinterpret_cast<const float4*>(b2s + c + k);
            const float y0 = v[k] + b.x, y1 = v[k + 1] + b.y, y2 = v[k + 2] + b.z, y3 = v[k + 3] + b.w;
            s += (y0 + y1) + (y2 + y3);
            ss = fmaf(y0, y0, ss); ss = fmaf(y1, y1, ss); ss = fmaf(y2, y2, ss); ss = fmaf(y3, y3, ss);
          }
        }
      }
      stat_s[half * 128 + q * 32 + lane] = make_float2(s, ss);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float mean = 0.0f, rstd = 1.0f;
      if (p.do_ln) {
        const float2 o = stat_s[(half ^ 1) * 128 + q * 32 + lane];
        mean = (s + o.x) * inv_n;
        rstd = rsqrtf(fmaxf((ss + o.y) * inv_n - mean * mean, 0.0f) + EF_LN_EPS);
      }
      const int64_t row0 = static_cast<int64_t>(tile) * 128 + q * 32;
      // The residual rows of a 32-column chunk (16 loads per lane: 2 columns x rows rsel, rsel + 2, ...) are requested one
      // chunk ahead, before the chunk in hand is normalised: a load per store-loop iteration would put one L2 / HBM
      // latency on the critical path of every 4 iterations (measured: 80 000 clk per tile instead of ~12 000).
      const bool has_res = p.residual != nullptr;
      const bool res_bf16 = p.res_dtype == GC_BF16;
      uint32_t rcur[16][2], rnxt[16][2];
      auto load_res = [&](int c, uint32_t (&dst)[16][2]) {
        if (!has_res) return;
        const int col = half * CH + c + 2 * cp;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int64_t row = row0 + 2 * it + rsel;
          dst[it][0] = 0u; dst[it][1] = 0u;
          if (row < p.rows) {
            if (res_bf16) {
              dst[it][0] = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + row * p.ld_res + col));
            } else {
              const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const float*>(p.residual) + row * p.ld_res + col));
              dst[it][0] = t.x; dst[it][1] = t.y;
            }
          }
        }
      };
      load_res(0, rcur);
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr, r);
#pragma unroll 1
      for (int c = 0; c < CH; c += 32) {
        float v[32];
        if (c + 32 < CH) load_res(c + 32, rnxt);
        tc_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
        if (c + 32 < CH) {
          tmem_ld_32x32b_x32(taddr + c + 32, r);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 b = *reinterpret_cast<const float4*>(b2s + c + k);
          const float4 sc = *reinterpret_cast<const float4*>(scs + c + k);
          const float4 of = *reinterpret_cast<const float4*>(ofs + c + k);
          float4 x;
          x.x = fmaf((v[k] + b.x - mean) * rstd, sc.x, of.x); x.y = fmaf((v[k + 1] + b.y - mean) * rstd, sc.y, of.y);
          x.z = fmaf((v[k + 2] + b.z - mean) * rstd, sc.z, of.z); x.w = fmaf((v[k + 3] + b.w - mean) * rstd, sc.w, of.w);
          *reinterpret_cast<float4*>(patch + lane * EF_PATCH_STRIDE + k) = x;
        }
        __syncwarp();
        // rows of the patch leave 16 lanes at a time (64 B of bf16 / 128 B of fp32 per row and instruction)
        const int col = half * CH + c + 2 * cp;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int rr = 2 * it + rsel;
          const int64_t row = row0 + rr;
          float2 x = *reinterpret_cast<const float2*>(patch + rr * EF_PATCH_STRIDE + 2 * cp);
          if (has_res) {
            if (res_bf16) {
              const float2 rsd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rcur[it][0]));
              x.x += rsd.x; x.y += rsd.y;
            } else {
              x.x += __uint_as_float(rcur[it][0]); x.y += __uint_as_float(rcur[it][1]);
            }
          }
          if (row < p.rows) {
            if (p.out_dtype == GC_BF16) {
              *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + col) = __floats2bfloat162_rn(x.x, x.y);
            } else {
              *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + row * p.ldo + col) = x;
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 16; ++it) { rcur[it][0] = rnxt[it][0]; rcur[it][1] = rnxt[it][1]; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int L>
int launch_linear_ln(cudaStream_t st, const CUtensorMap& a_map, const CUtensorMap& w_map, const LinearLnParams& p) {
  using C = EFCfg<L>;
  GC_CHECK_CUDA(cudaFuncSetAttribute(linear_ln_cond_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM),
                "cudaFuncSetAttribute(linear_ln_cond_kernel)");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = static_cast<unsigned>(p.num_tiles < sms ? p.num_tiles : sms);
  GC_CHECK_CUDA(launch_kernel(linear_ln_cond_kernel<L>, dim3(grid), dim3(LN_THREADS), (size_t)C::SMEM, st, a_map, w_map, p),
                "linear_ln_cond_kernel");
  return GC_OK;
}

template <int L>
int launch_edge_pair(cudaStream_t st, const CUtensorMap& w_map, const CUtensorMap& base_map, const CUtensorMap& gr_map,
                     const EdgeFusedParams& p) {
  using C = EPCfg<L>;
  GC_CHECK_CUDA(cudaFuncSetAttribute(edge_mlp_sum3_pair_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM),
                "cudaFuncSetAttribute(edge_mlp_sum3_pair_kernel)");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = sms / 2;
  if (p.num_tiles < pairs) pairs = p.num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(EP_THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  GC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, edge_mlp_sum3_pair_kernel<L>, w_map, base_map, gr_map, p), "edge_mlp_sum3_pair_kernel");
  return GC_OK;
}

// GENCAST_EDGE_PAIR=1: column-split CTA pair (two accumulators per CTA), for L in {256, 512}.  Opt-in: measured equal to
// the single-CTA kernel (733 vs 737 us at 1 deg x 4): with the phases overlapped, producers and epilogue both slow
// down - what saturates is the SM's LSU data pipe (61 % of peak averaged over the single-CTA kernel, ~100 % while its
// epilogue runs: broadcast LDS.128 of the bias / scale / offset vectors = 4 wavefronts each, two SHFL per element),
// and that work per tile is the same in both variants.  Read at every call so that tests can switch it.
bool edge_pair_enabled() {
  const char* v = getenv("GENCAST_EDGE_PAIR");
  return v != nullptr && v[0] == '1';
}

// GENCAST_EDGE_TMA=0: all three operands through registers (the round-2 first version), for A/B measurements
bool edge_tma_enabled() {
  static const bool on = []() {
    const char* v = getenv("GENCAST_EDGE_TMA");
    return !(v != nullptr && v[0] == '0');
  }();
  return on;
}

template <int L, int MODE>
int launch_edge_fused(cudaStream_t st, const CUtensorMap& w_map, const CUtensorMap& base_map, const CUtensorMap& gr_map,
                      const EdgeFusedParams& p) {
  using C = EFCfg<L, MODE>;
  static_assert(C::SMEM <= 232448, "edge_mlp_sum3_kernel: shared memory plan does not fit");
  GC_CHECK_CUDA(cudaFuncSetAttribute(edge_mlp_sum3_kernel<L, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM),
                "cudaFuncSetAttribute(edge_mlp_sum3_kernel)");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = static_cast<unsigned>(p.num_tiles < sms ? p.num_tiles : sms);
  GC_CHECK_CUDA(launch_kernel(edge_mlp_sum3_kernel<L, MODE>, dim3(grid), dim3(EF_THREADS), (size_t)C::SMEM, st, w_map, base_map,
                              gr_map, p),
                "edge_mlp_sum3_kernel");
  return GC_OK;
}

}  // namespace
}  // namespace gc

extern "C" __attribute__((visibility("default"))) void gc_debug_set_edge_fused_trace(void* ptr) {
  gc::g_edge_fused_trace = reinterpret_cast<long long*>(ptr);
}

extern "C" int gc_linear_ln_cond(void* stream, const void* a, int64_t lda, int64_t rows, const void* w, int64_t ldw, const float* bias,
                                 const float* scale_offset, int32_t do_layer_norm, const void* residual, int32_t res_dtype,
                                 int64_t ld_res, void* out, int32_t out_dtype, int64_t ldo, int32_t cols) {
  using namespace gc;
  GC_REQUIRE(a && w && out, "gc_linear_ln_cond: null buffer");
  GC_REQUIRE(cols == 128 || cols == 256 || cols == 512, "gc_linear_ln_cond: cols=%d (supported: 128, 256, 512)", cols);
  GC_REQUIRE(rows >= 0 && rows < (1LL << 31), "gc_linear_ln_cond: rows=%lld", (long long)rows);
  GC_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && lda >= cols && ldw >= cols && ldo >= cols && aligned16(a) &&
                 aligned16(w) && aligned16(out), "gc_linear_ln_cond: alignment");
  GC_REQUIRE(out_dtype == GC_BF16 || out_dtype == GC_F32, "gc_linear_ln_cond: bad out dtype");
  if (residual != nullptr)
    GC_REQUIRE((res_dtype == GC_BF16 || res_dtype == GC_F32) && ld_res % 8 == 0 && aligned16(residual), "gc_linear_ln_cond: residual");
  if (rows == 0) return GC_OK;
  CUtensorMap a_map, w_map;
  int rc = make_tmap_bf16_2d(&a_map, a, (uint64_t)rows, (uint64_t)cols, (uint64_t)lda, 64, 128);
  if (rc != GC_OK) return rc;
  rc = make_tmap_bf16_2d(&w_map, w, (uint64_t)cols, (uint64_t)cols, (uint64_t)ldw, 64, 128);
  if (rc != GC_OK) return rc;
  LinearLnParams p;
  p.b2 = bias; p.scale_offset = scale_offset; p.do_ln = do_layer_norm;
  p.residual = residual; p.res_dtype = res_dtype; p.ld_res = ld_res;
  p.out = out; p.out_dtype = out_dtype; p.ldo = ldo; p.rows = rows;
  p.num_tiles = (int)((rows + 127) / 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cols == 128) return launch_linear_ln<128>(st, a_map, w_map, p);
  if (cols == 256) return launch_linear_ln<256>(st, a_map, w_map, p);
  return launch_linear_ln<512>(st, a_map, w_map, p);
}

extern "C" int gc_edge_mlp_sum3(void* stream, const void* base, int64_t ld_base, int64_t period, const void* gs,
                                const int32_t* idx_s, int64_t ld_gs, const void* gr, const int32_t* idx_r, int64_t ld_gr,
                                int32_t act, const void* w2, int64_t ld_w2, const float* b2, const float* scale_offset,
                                int32_t do_layer_norm, void* out, int32_t out_dtype, int64_t ldo, int64_t num_receivers,
                                int32_t cols) {
  using namespace gc;
  GC_REQUIRE(base && gs && idx_s && gr && w2 && out, "gc_edge_mlp_sum3: null buffer");
  GC_REQUIRE(cols == 128 || cols == 256 || cols == 512, "gc_edge_mlp_sum3: cols=%d (supported: 128, 256, 512)", cols);
  GC_REQUIRE(period > 0 && num_receivers > 0 && num_receivers < (1LL << 31) / 3, "gc_edge_mlp_sum3: bad sizes");
  GC_REQUIRE(ld_base % 8 == 0 && ld_gs % 8 == 0 && ld_gr % 8 == 0 && ld_w2 % 8 == 0 && ldo % 8 == 0 && aligned16(base) &&
                 aligned16(gs) && aligned16(gr) && aligned16(w2) && aligned16(out),
             "gc_edge_mlp_sum3: alignment");
  GC_REQUIRE(out_dtype == GC_BF16 || out_dtype == GC_F32, "gc_edge_mlp_sum3: bad out dtype");
  GC_REQUIRE(act == GC_ACT_NONE || act == GC_ACT_SWISH || act == GC_ACT_GELU_TANH, "gc_edge_mlp_sum3: act=%d", act);
  CUtensorMap w_map;
  int rc = make_tmap_bf16_2d(&w_map, w2, (uint64_t)cols, (uint64_t)cols, (uint64_t)ld_w2, 64, 128);
  if (rc != GC_OK) return rc;
  EdgeFusedParams p;
  p.base = reinterpret_cast<const __nv_bfloat16*>(base); p.ld_base = ld_base; p.period = period;
  p.gs = reinterpret_cast<const __nv_bfloat16*>(gs); p.idx_s = idx_s; p.ld_gs = ld_gs;
  p.gr = reinterpret_cast<const __nv_bfloat16*>(gr); p.idx_r = idx_r; p.ld_gr = ld_gr;
  p.act = act; p.b2 = b2; p.scale_offset = scale_offset; p.do_ln = do_layer_norm;
  p.out = out; p.out_dtype = out_dtype; p.ldo = ldo; p.num_receivers = num_receivers;
  p.num_tiles = (int)((num_receivers + EF_RECV_PER_TILE - 1) / EF_RECV_PER_TILE);
  p.members = 1; p.tiles_per_member = p.num_tiles; p.num_rows = 0; p.row_stats = nullptr;
  p.trace = g_edge_fused_trace;
  {
    const int64_t edges = 3 * num_receivers, per_tile = 3 * EF_RECV_PER_TILE;
    if (period < edges && edges % period == 0 && period % per_tile == 0) {
      p.members = (int)(edges / period);
      p.tiles_per_member = (int)(period / per_tile);
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // FAST: the tile's base rows and receiver rows are contiguous, so they travel by TMA (see the kernel's header)
  const bool fast = idx_r == nullptr && period % 30 == 0 && edge_tma_enabled();
  if (fast) {
    CUtensorMap base_map, gr_map;
    rc = make_tmap_bf16_2d(&base_map, base, (uint64_t)period, (uint64_t)cols, (uint64_t)ld_base, 64, 30);
    if (rc != GC_OK) return rc;
    rc = make_tmap_bf16_2d(&gr_map, gr, (uint64_t)num_receivers, (uint64_t)cols, (uint64_t)ld_gr, 64, EF_RECV_PER_TILE);
    if (rc != GC_OK) return rc;
    if (edge_pair_enabled() && cols >= 256) {
      // W2 boxes of 128 rows as above; each CTA of a pair loads its half of the rows
      if (cols == 256) return launch_edge_pair<256>(st, w_map, base_map, gr_map, p);
      return launch_edge_pair<512>(st, w_map, base_map, gr_map, p);
    }
    if (cols == 128) return launch_edge_fused<128, 1>(st, w_map, base_map, gr_map, p);
    if (cols == 256) return launch_edge_fused<256, 1>(st, w_map, base_map, gr_map, p);
    return launch_edge_fused<512, 1>(st, w_map, base_map, gr_map, p);
  }
  if (cols == 128) return launch_edge_fused<128, 0>(st, w_map, w_map, w_map, p);
  if (cols == 256) return launch_edge_fused<256, 0>(st, w_map, w_map, w_map, p);
  return launch_edge_fused<512, 0>(st, w_map, w_map, w_map, p);
}

extern "C" int gc_edge_mlp_rows(void* stream, const void* base, int64_t ld_base, int64_t period, const void* gs,
                                const int32_t* idx_s, int64_t ld_gs, int32_t act, const void* w2, int64_t ld_w2,
                                const float* b2, void* out, int64_t ldo, int64_t num_rows, int32_t cols, float* row_stats) {
  using namespace gc;
  GC_REQUIRE(base && gs && idx_s && w2 && out, "gc_edge_mlp_rows: null buffer");
  GC_REQUIRE(cols == 128 || cols == 256 || cols == 512, "gc_edge_mlp_rows: cols=%d (supported: 128, 256, 512)", cols);
  GC_REQUIRE(period > 0 && num_rows >= 0 && num_rows % period == 0 && num_rows < (1LL << 31),
             "gc_edge_mlp_rows: num_rows must be a multiple of period");
  GC_REQUIRE(ld_base % 8 == 0 && ld_gs % 8 == 0 && ld_w2 % 8 == 0 && ldo % 8 == 0 && aligned16(base) && aligned16(gs) &&
                 aligned16(w2) && aligned16(out), "gc_edge_mlp_rows: alignment");
  GC_REQUIRE(act == GC_ACT_NONE || act == GC_ACT_SWISH || act == GC_ACT_GELU_TANH, "gc_edge_mlp_rows: act=%d", act);
  GC_REQUIRE(row_stats == nullptr || aligned16(row_stats), "gc_edge_mlp_rows: row_stats alignment");
  if (num_rows == 0) return GC_OK;
  CUtensorMap w_map, base_map, out_map;
  int rc = make_tmap_bf16_2d(&w_map, w2, (uint64_t)cols, (uint64_t)cols, (uint64_t)ld_w2, 64, 128);
  if (rc != GC_OK) return rc;
  rc = make_tmap_bf16_2d(&base_map, base, (uint64_t)period, (uint64_t)cols, (uint64_t)ld_base, 64, 128);
  if (rc != GC_OK) return rc;
  rc = make_tmap_bf16_2d(&out_map, out, (uint64_t)num_rows, (uint64_t)cols, (uint64_t)ldo, 64, 32);
  if (rc != GC_OK) return rc;
  EdgeFusedParams p;
  p.base = reinterpret_cast<const __nv_bfloat16*>(base); p.ld_base = ld_base; p.period = period;
  p.gs = reinterpret_cast<const __nv_bfloat16*>(gs); p.idx_s = idx_s; p.ld_gs = ld_gs;
  p.gr = nullptr; p.idx_r = nullptr; p.ld_gr = 0;
  p.act = act; p.b2 = b2; p.scale_offset = nullptr; p.do_ln = 0;
  p.out = out; p.out_dtype = GC_BF16; p.ldo = ldo; p.num_receivers = 0; p.num_rows = num_rows;
  p.row_stats = row_stats;
  p.members = (int)(num_rows / period);
  p.tiles_per_member = (int)((period + 127) / 128);
  p.num_tiles = p.members * p.tiles_per_member;
  p.trace = g_edge_fused_trace;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // the kernel's third tensor-map slot (receiver rows in the degree-3 variant) carries the output map here
  if (cols == 128) return launch_edge_fused<128, 2>(st, w_map, base_map, out_map, p);
  if (cols == 256) return launch_edge_fused<256, 2>(st, w_map, base_map, out_map, p);
  return launch_edge_fused<512, 2>(st, w_map, base_map, out_map, p);
}
