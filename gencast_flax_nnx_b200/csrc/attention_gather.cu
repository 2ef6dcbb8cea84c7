// k-hop attention on tcgen05 tensor cores over per-query-tile COMPACTED key lists (bf16 in, fp32 accumulate).
//
// Same operator as attention_tc.cu (TriblockdiagMHA without its projections,
// gencast/sparse_transformer.py:309-354; softmax :100-125; mask :163-201), different work decomposition.
// The (query tile, key tile) list of attention_tc.cu walks 11.3 key tiles of 128 per 128-query patch at
// 1 deg although the patch attends to only 690 distinct keys on average (14.7 % of the listed tile entries are
// live).  Here the host lists, per query tile, the sorted union of its keys (graph.khop_compact_steps), cut in
// steps of 64; the kernel gathers those K / V rows itself into dense 128B-swizzled operand tiles, so a patch
// takes 11 steps of 64 keys (5.6 tiles' worth) and every tensor-core column is a key somebody attends to.
//
// One CTA = (128-query tile, head), two CTAs per SM (96 KB of shared memory and 256 TMEM columns each), so the
// prologue / epilogue of one overlaps the main loop of the other and the SM's tensor pipe, TMEM read port and
// MUFU see two independent instruction streams.  Per step t (64 keys):
//     S_t = Q K_t^T        tcgen05.mma M=128 N=64 K=d, Q TMA-staged once, K_t gathered (cp.async 16 B, manual
//                          128-byte swizzle: the layout cp.async.bulk.tensor would have produced)
//     P_t = exp2(S_t c - m c)  four softmax warps (one per TMEM lane quarter) read S from TMEM once, apply the
//                          64-bit row masks and write bf16 P back over the S columns it came from
//     O  += P_t V_t        TS-form tcgen05.mma (A = P from TMEM, B = gathered V tile, MN-major)
// S is double buffered (2 x 64 columns), so S_{t+1} is in TMEM when the softmax warps finish step t, and the
// in-order tensor pipe lets S_{t+2} overwrite P_t right behind P_t V_t without further synchronisation.
//
// Softmax is exact, online and SINGLE PASS over S: TMEM is read at 64 B/clk per SM, one pass over a 128 x 64
// fp32 tile costs 512 clk, as much as the tile's MMAs, so the usual max-then-exp double read would make the TMEM
// port the bottleneck.  A row's offset m is fixed the first time the row meets a neighbour (that step is read
// twice: masked maximum, then exponentials) and afterwards exponentials are taken against that stale offset,
// which is exact after the final division by l as long as nothing overflows.  The guard is the step's own row
// sum: if it exceeds 2^40 (a logit more than 40 / c above the offset) the row's offset is raised to the step's
// maximum, l and the row of O in TMEM are rescaled (after the preceding P V has completed) and the step is redone.
// Masked logits contribute exactly 0, which is what the reference's where(mask, logits, -1e30) + softmax gives.
// 32-key half-steps in which none of a warp's 32 queries has a neighbour are neither read nor exponentiated.
//
// Warps (288 threads): 0 = MMA issuer / TMEM owner / Q load, 1-4 = softmax + epilogue, 5-8 = K / V row gather.
#include "common.cuh"
#include "sm100.cuh"

namespace gc {

long long* g_attention_gather_trace = nullptr;   // debug: clock stamps of CTA 0 (null in normal use)

namespace {

// Debug timeline: CTA 0 records clock64() at pipeline events into trace[role * 512 + index].
#define GC_GTR(role, index)                                                                                       \
  do {                                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && (index) < 512) p.trace[(role) * 512 + (index)] = clock64();      \
  } while (0)

constexpr int GQ = 128;            // queries per tile
constexpr int GS = 64;             // keys per step
constexpr int G_THREADS = 288;
constexpr int G_LOADERS = 4;
constexpr int G_MAX_STAGED_STEPS = 32;   // key lists up to this many steps are staged in shared memory (8 KB)

template <int D>
struct GCfg {
  static constexpr int CHUNKS = D / 64;                 // 64-element (128 B) operand chunks along d
  static constexpr int SLOT_BYTES = GS * D * 2;         // one gathered K or V step tile
  static constexpr int Q_BYTES = GQ * D * 2;
  static constexpr int NK = D == 64 ? 4 : 2;            // K ring: freed as soon as S_t has been formed
  static constexpr int NV = D == 64 ? 4 : 2;            // V ring: freed when P_t V_t has completed
  static constexpr int KEYS_BYTES = G_MAX_STAGED_STEPS * GS * 4;
  static constexpr int SMEM = Q_BYTES + (NK + NV) * SLOT_BYTES + KEYS_BYTES + 256 /*barriers*/ + 1024 /*align*/;
  static constexpr uint32_t TMEM_COLS = 256;            // S_0, S_1 (64 each) | O (D) at column 128
};

struct GatherAttParams {
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
  const int32_t* step_ptr;     // [num_q_tiles + 1]
  const int32_t* keys;         // [num_steps * 64] row of qkv holding the key / value of each compacted column
  const uint2* mask;           // [mask steps][128 rows] 64 bits: bit j = column j of the step is a neighbour of the row
  const int32_t* work;         // [num_q_tiles] launch order of the query tiles (longest first)
  int mask_period;             // masks repeat with this period over the step index (members share them); 0 = no repeat
  __nv_bfloat16* out;
  int64_t ldo;
  int nodes;
  int heads;
  int hd;                      // heads * head_dim: column offset of K inside a qkv row (V at 2 * hd)
  float scale_log2e;           // head_dim^-0.5 * log2(e)
  long long* trace;
};

__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// The executing thread arrives on `bar` once all of its earlier cp.async copies have landed (asynchronously: the
// thread itself does not wait; .noinc: the arrival is counted in the barrier's initial count).
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 32 logits of one row -> 16 packed bf16 pairs of P = exp2(s c + noff) (0 where the mask bit is clear);
// partial row sums go to ls[0..1].
__device__ __forceinline__ void exp_chunk(const uint32_t (&v)[32], uint32_t mw, float c2, float noff, uint32_t (&packed)[16],
                                          float (&ls)[2]) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float e0 = ex2_approx(fmaf(__uint_as_float(v[i]), c2, noff));
    const float e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), c2, noff));
    const float p0 = (mw & (1u << i)) ? e0 : 0.0f;
    const float p1 = (mw & (1u << (i + 1))) ? e1 : 0.0f;
    ls[(i >> 1) & 1] += p0 + p1;
    const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
    packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
  }
}

// TMA_ROWS: the K / V rows of a step are fetched by TMA row gathers (cp.async.bulk.tensor tile::gather4, four rows per
// instruction, issued by one warp per stream) instead of 16-byte cp.async copies by four warps.
template <int D, int ROWS>
__global__ void __launch_bounds__(G_THREADS, 2)
khop_attention_gather_kernel(const __grid_constant__ CUtensorMap q_map, const __grid_constant__ CUtensorMap row_map,
                             const __grid_constant__ CUtensorMap out_map, const GatherAttParams p) {
  using namespace sm100;
  using C = GCfg<D>;
  constexpr bool TMA_K = ROWS != 0, TMA_V = ROWS == 1;     // which streams use TMA row gathers
  constexpr bool TMA_ROWS = TMA_K;
  pdl_launch_dependents();
  if (threadIdx.x == 0) GC_GTR(6, 0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;
  const uint32_t k_smem = q_smem + C::Q_BYTES;
  const uint32_t v_smem = k_smem + C::NK * C::SLOT_BYTES;
  const uint32_t keys_smem = v_smem + C::NV * C::SLOT_BYTES;
  const uint32_t bars = keys_smem + C::KEYS_BYTES;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (1 + C::NK + s); };
  auto v_full = [&](int s) { return bars + 8u * (1 + 2 * C::NK + s); };
  auto v_empty = [&](int s) { return bars + 8u * (1 + 2 * C::NK + C::NV + s); };
  const uint32_t b2 = bars + 8u * (1 + 2 * C::NK + 2 * C::NV);
  auto s_full = [&](int b) { return b2 + 8u * b; };
  auto p_full = [&](int b) { return b2 + 8u * (2 + b); };
  const uint32_t pv_done = b2 + 8u * 4;
  const uint32_t o_full = b2 + 8u * 5;
  const uint32_t tmem_ptr_smem = b2 + 8u * 6;
  int32_t* keys_s = reinterpret_cast<int32_t*>(smem_raw + (keys_smem - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  // static tables (never written by queued work): read before the wait for the predecessor grid
  const int qt = __ldg(p.work + blockIdx.x / p.heads);
  const int s_beg = __ldg(p.step_ptr + qt);
  const int T = __ldg(p.step_ptr + qt + 1) - s_beg;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&q_map);
    prefetch_tensormap(&out_map);
    if (TMA_ROWS) prefetch_tensormap(&row_map);
    mbar_init(q_full, 1);
    // "full" barriers: one arrive.expect_tx (TMA), or one arrival per copying lane (cp.async)
    for (int s = 0; s < C::NK; ++s) { mbar_init(k_full(s), TMA_K ? 1u : (G_LOADERS / 2) * 32); mbar_init(k_empty(s), 1); }
    for (int s = 0; s < C::NV; ++s) { mbar_init(v_full(s), TMA_V ? 1u : (G_LOADERS / 2) * 32); mbar_init(v_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(p_full(b), 4); }
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  {
    // this query tile's key list -> shared memory (the gather warps read every entry twice: K and V)
    const int n_stage = min(T, G_MAX_STAGED_STEPS) * GS;
    for (int i = threadIdx.x; i < n_stage; i += G_THREADS) keys_s[i] = __ldg(p.keys + static_cast<int64_t>(s_beg) * GS + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tmem_s = tmem_base;            // S_0 at +0, S_1 at +64
  const uint32_t tmem_o = tmem_base + 128;      // O at +128 .. +128 + D

  if (warp == 0) {
    // ---------------- Q load + MMA issuer.  The whole warp runs this code and one elected lane issues the
    // tcgen05 / TMA instructions: inside an `if (lane == 0)` region the compiler treats every operand as divergent
    // and wraps each tcgen05.mma in an ELECT / R2UR.BROADCAST x5 / BRA.U.ANY loop (~15 dependent instructions, ~100 clk
    // per MMA for an instruction that computes for 32-64 clk); warp-uniform control flow keeps descriptors in
    // uniform registers.
    pdl_wait();
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
      for (int c = 0; c < C::CHUNKS; ++c) tma_load_2d(q_smem + c * (GQ * 128), &q_map, q_full, head * D + 64 * c, qt * GQ);
    }
    __syncwarp();
    constexpr uint32_t idesc_s = idesc_bf16_f32(GQ, GS, 0, 0);
    constexpr uint32_t idesc_o = idesc_bf16_f32(GQ, D, 0, 1);     // B = V tile, MN-major
    const uint64_t dq0 = desc_kmajor_sw128(q_smem);
    int kslot = 0, vslot = 0;
    uint32_t kphase = 0, vphase = 0;
    auto issue_s = [&](int t) {
      const int b = t & 1;
      GC_GTR(1, 2 * t);
      mbar_wait(k_full(kslot), kphase);
      GC_GTR(1, 2 * t + 1);
      if (!TMA_K) fence_proxy_async_smem();      // rows written by cp.async (generic proxy), read by the MMA (async proxy)
      tc_fence_after();
      const uint64_t dk0 = desc_kmajor_sw128(k_smem + kslot * C::SLOT_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < D / 16; ++j) {
          // descriptor start-address field is in 16-byte units: 64-wide d chunk j / 4, 32 bytes per K = 16 step inside it
          const uint32_t off_q = ((j >> 2) * (GQ * 128) + (j & 3) * 32) >> 4;
          const uint32_t off_k = ((j >> 2) * (GS * 128) + (j & 3) * 32) >> 4;
          umma_f16(tmem_s + b * 64, dq0 + off_q, dk0 + off_k, idesc_s, j > 0);
        }
        umma_commit(k_empty(kslot));
        umma_commit(s_full(b));
      }
      __syncwarp();
      if (++kslot == C::NK) { kslot = 0; kphase ^= 1u; }
    };
    mbar_wait(q_full, 0);
    for (int t = 0; t < 2 && t < T; ++t) issue_s(t);
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      GC_GTR(2, 3 * t);
      mbar_wait(v_full(vslot), vphase);
      GC_GTR(2, 3 * t + 1);
      mbar_wait(p_full(b), (t >> 1) & 1);
      GC_GTR(2, 3 * t + 2);
      if (!TMA_V) fence_proxy_async_smem();
      tc_fence_after();
      // A: P[128 x 16 keys] from tensor memory (lane = query row, 8 columns of two bf16 each);
      // B: V[16 keys x D], MN-major: 16 key rows of 128 B start at j * 2048, 64-wide d chunks GS * 128 B apart
      const uint64_t dv0 = desc_mnmajor_sw128(v_smem + vslot * C::SLOT_BYTES, GS * 128, 1024);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < GS / 16; ++j)
          umma_f16_ts(tmem_o, tmem_s + b * 64 + 8 * j, dv0 + static_cast<uint32_t>(j * (2048 >> 4)), idesc_o, (t > 0 || j > 0) ? 1u : 0u);
        umma_commit(v_empty(vslot));
        umma_commit(pv_done);
      }
      __syncwarp();
      if (++vslot == C::NV) { vslot = 0; vphase ^= 1u; }
      // S_{t+2} overwrites the columns of S_t / P_t: in issue order behind P_t V_t, which read them
      if (t + 2 < T) issue_s(t + 2);
    }
    if (elect_one()) umma_commit(o_full);
    __syncwarp();
  } else if (warp <= 4) {
    // ---------------- softmax + epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 = query rows of the tile
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = p.scale_log2e;
    const uint32_t o_addr = tmem_o + lane_addr;
    // masks repeat with period p.mask_period over the global step index (members share one copy); a query tile's
    // steps are contiguous inside one period, so the reduction is done once (a per-step modulo is a ~30-instruction
    // sequence through the XU pipe on the softmax warps' critical path)
    const int gs0 = p.mask_period > 0 ? s_beg % p.mask_period : s_beg;
    const uint2* mask_row = p.mask + static_cast<int64_t>(gs0) * GQ + r;
    auto mask_of = [&](int t) { return __ldg(mask_row + static_cast<int64_t>(t) * GQ); };
    float m = -INFINITY;                                 // the row's offset, in raw logit units
    float l0 = 0.0f, l1 = 0.0f;
    uint2 mk_next = make_uint2(0u, 0u);
    if (T > 0) mk_next = mask_of(0);
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      const uint2 mk = mk_next;
      if (t + 1 < T) mk_next = mask_of(t + 1);
      const uint32_t mw[2] = {mk.x, mk.y};
      bool live[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) live[c] = __any_sync(0xffffffffu, mw[c] != 0u);
      const uint32_t s_addr = tmem_s + b * 64 + lane_addr;
      if (threadIdx.x == 32) GC_GTR(3, 3 * t);
      mbar_wait(s_full(b), (t >> 1) & 1);
      if (threadIdx.x == 32) GC_GTR(3, 3 * t + 1);
      tc_fence_after();
      uint32_t packed[2][16];
      float ls[2];
      auto pass = [&](float noff) {
        ls[0] = 0.0f; ls[1] = 0.0f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (live[c]) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(s_addr + c * 32, v);
            tc_wait_ld();
            exp_chunk(v, mw[c], c2, noff, packed[c], ls);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[c][i] = 0u;
          }
        }
      };
      pass(m == -INFINITY ? 0.0f : -m * c2);
      // first neighbour of the row, or exponentials that left the safe range: fix the offset and redo the step
      const bool bad = (m == -INFINITY && (mw[0] | mw[1]) != 0u) || !(ls[0] + ls[1] <= 1.0995116e12f);
      if (__any_sync(0xffffffffu, bad)) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (live[c]) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(s_addr + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (mw[c] & (1u << i)) ? __uint_as_float(v[i]) : -INFINITY);
          }
        }
        const float m_new = bad ? fmaxf(m, mx) : m;
        const bool resc = bad && m != -INFINITY;           // rows that already carry a sum and an O row
        if (__any_sync(0xffffffffu, resc)) {
          const float alpha = resc ? exp2f(m * c2 - m_new * c2) : 1.0f;
          l0 *= alpha; l1 *= alpha;
          if (t > 0) {
            mbar_wait(pv_done, (t - 1) & 1);                // P_{t-1} V_{t-1} (and everything before it) has completed
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < D; c += 32) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(o_addr + c, o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x32(o_addr + c, o);
            }
            tc_wait_st();
          }
        }
        m = m_new;
        pass(m == -INFINITY ? 0.0f : -m * c2);
      }
      l0 += ls[0]; l1 += ls[1];
      // keys 32c .. 32c+31 of the row -> columns 16c .. 16c+15 of the step's TMEM region, over S columns that have
      // been read (chunk c of S sits at columns 32c .. 32c+31)
      tmem_st_32x32b_x16(s_addr, packed[0]);
      tmem_st_32x32b_x16(s_addr + 16, packed[1]);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(b));
      if (threadIdx.x == 32) GC_GTR(3, 3 * t + 2);
    }
    // ---- epilogue: O / l
    const float l = l0 + l1;
    const float inv_l = l > 0.0f ? 1.0f / l : 0.0f;
    if (threadIdx.x == 32) GC_GTR(4, 0);
    if (T > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 32) GC_GTR(4, 1);
    const int64_t row = static_cast<int64_t>(qt) * GQ + r;
    if (T > 0) {
      // O / l as bf16 into the Q tile's shared memory (every MMA that read Q has completed: o_full), in the 128B-swizzled
      // layout of a {64 columns, 32 rows} TMA box per warp and 64-column block, then out by TMA: a lane-per-row st.global
      // touches 32 lines per instruction (2 700 clk of the CTA's 4 500-clk epilogue in the trace).  Rows past the last
      // node are clipped by the tensor map.
      const uint32_t sw = static_cast<uint32_t>(r & 7);
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(o_addr + c, v);
        tc_wait_ld();
        const uint32_t blk = q_smem + static_cast<uint32_t>(c >> 6) * (GQ * 128);
        const int hc = (c >> 5) & 1;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            h[j] = __floats2bfloat162_rn(__uint_as_float(v[i + 2 * j]) * inv_l, __uint_as_float(v[i + 2 * j + 1]) * inv_l);
          const uint32_t unit = static_cast<uint32_t>(hc * 4 + (i >> 3)) ^ sw;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + static_cast<uint32_t>(r) * 128u + unit * 16u), "r"(o.x),
                       "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
        if (hc == 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&out_map, blk + static_cast<uint32_t>(q * 32) * 128u, head * D + (c & ~63), qt * GQ + q * 32);
            bulk_commit_group();
          }
        }
      }
      if (lane == 0) bulk_wait_group_read<0>();      // the stores have read this CTA's shared memory before it exits
    } else if (row < p.nodes) {
      // a query tile without keys: zeros (the Q load may still be in flight, so shared memory is left alone)
      __nv_bfloat16* dst = p.out + row * p.ldo + head * D;
#pragma unroll
      for (int i = 0; i < D; i += 8) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0u, 0u, 0u, 0u);
    }
  } else {
    // ---------------- K / V row gather: two independent streams, so a K tile is requested the moment its slot frees
    // (S_t formed) and never queues behind a V tile whose slot only frees when P_{t-2} V_{t-2} has completed.
    // cp.async variant: warps 5, 6 fetch the K tiles (32 rows each), warps 7, 8 the V tiles.
    const bool v_stream = warp >= 7;
    pdl_wait();                                  // qkv is the predecessor's output
    int ld_ev = 0;
    const int k_col = p.hd + head * D;
    const int v_col = 2 * p.hd + head * D;
    if ((v_stream && TMA_V) || (!v_stream && TMA_K)) {
      // Warp 5 feeds the K ring, warp 7 the V ring (6 and 8 have nothing to do).  Lane = (64-wide d chunk, group of four
      // consecutive key slots): one gather4 per lane puts its four rows at rows 4 g .. 4 g + 3 of the chunk's block, in
      // the same swizzled layout the MMA descriptors expect, so a whole tile is requested by one warp-wide instruction.
      if (warp == 5 || warp == 7) {
        constexpr int GROUPS = GS / 4;             // 16
        const int grp = lane % GROUPS, chunk = lane / GROUPS;
        const bool active = chunk < C::CHUNKS;
        const int col = (v_stream ? v_col : k_col) + 64 * chunk;
        const uint32_t dst_off = static_cast<uint32_t>(chunk) * (GS * 128) + static_cast<uint32_t>(grp) * 512u;
        int slot = 0;
        uint32_t phase = 0;
        constexpr int NSLOT = C::NK;               // NK == NV
        for (int t = 0; t < T; ++t) {
          int4 kk = make_int4(0, 0, 0, 0);
          if (active) {
            if (t < G_MAX_STAGED_STEPS) kk = *reinterpret_cast<const int4*>(keys_s + t * GS + 4 * grp);
            else kk = __ldg(reinterpret_cast<const int4*>(p.keys + (static_cast<int64_t>(s_beg) + t) * GS) + grp);
          }
          const uint32_t empty = v_stream ? v_empty(slot) : k_empty(slot);
          const uint32_t full = v_stream ? v_full(slot) : k_full(slot);
          const uint32_t base_s = (v_stream ? v_smem : k_smem) + slot * C::SLOT_BYTES;
          mbar_wait(empty, phase ^ 1u);
          if (lane == 0) { GC_GTR(v_stream ? 5 : 0, ld_ev); ++ld_ev; }
          if (lane == 0) mbar_arrive_expect_tx(full, C::SLOT_BYTES);
          __syncwarp();
          if (active) tma_gather4_2d(base_s + dst_off, &row_map, full, col, kk.x, kk.y, kk.z, kk.w);
          if (++slot == NSLOT) { slot = 0; phase ^= 1u; }
        }
      }
    } else {
      constexpr int UNITS = D / 8;                 // 16-byte units per row
      constexpr int RPI = 32 / UNITS;              // rows per warp instruction
      constexpr int ITER = 32 / RPI;
      const int lw = (warp - 5) & 1;               // which half of the tile's rows
      const int unit = lane % UNITS;
      const int row_in = lane / UNITS;
      const uint32_t dst_unit = static_cast<uint32_t>(unit >> 3) * (GS * 128);   // 64-wide d chunk of this unit
      const uint32_t u8 = static_cast<uint32_t>(unit & 7);
      auto key_of = [&](int t, int rl) -> int64_t {
        if (t < G_MAX_STAGED_STEPS) return keys_s[t * GS + rl];
        return __ldg(p.keys + (static_cast<int64_t>(s_beg) + t) * GS + rl);
      };
      // A gather warp never waits for its own copies: every lane's cp.async.mbarrier.arrive.noinc makes the tile's
      // "full" barrier count that lane's copies when they land.  (Waiting with cp.async.wait_group and arriving with a
      // plain mbarrier.arrive costs a MEMBAR.ALL that also waits for the NEXT tile's copies issued just before: one tile
      // in flight per warp and 2 700 clk from issue to signal, which made the softmax warps wait for S half of the time.)
      auto produce = [&](uint32_t empty_bar, uint32_t parity, uint32_t full_bar, uint32_t slot_base, int col0, int t) {
        if (lane == 0) mbar_wait(empty_bar, parity);
        __syncwarp();
        if (lane == 0 && lw == 0) { GC_GTR(v_stream ? 5 : 0, ld_ev); ++ld_ev; }
  #pragma unroll
        for (int i = 0; i < ITER; ++i) {
          const int rl = lw * 32 + i * RPI + row_in;
          const int64_t key = key_of(t, rl);
          const __nv_bfloat16* src = p.qkv + key * p.ld_qkv + col0 + unit * 8;
          const uint32_t dst = slot_base + dst_unit + static_cast<uint32_t>(rl) * 128u + ((u8 ^ (static_cast<uint32_t>(rl) & 7u)) << 4);
          cp_async_16(dst, src);
        }
        cp_async_arrive_noinc(full_bar);
      };
      int kslot = 0, vslot = 0;
      uint32_t kphase = 0, vphase = 0;
      auto load_k = [&](int t) {
        produce(k_empty(kslot), kphase ^ 1u, k_full(kslot), k_smem + kslot * C::SLOT_BYTES, k_col, t);
        if (++kslot == C::NK) { kslot = 0; kphase ^= 1u; }
      };
      auto load_v = [&](int t) {
        produce(v_empty(vslot), vphase ^ 1u, v_full(vslot), v_smem + vslot * C::SLOT_BYTES, v_col, t);
        if (++vslot == C::NV) { vslot = 0; vphase ^= 1u; }
      };
      if (v_stream) {
        for (int t = 0; t < T; ++t) load_v(t);
      } else {
        for (int t = 0; t < T; ++t) load_k(t);
      }
    }
  }
  if (threadIdx.x == 32) GC_GTR(4, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
  if (threadIdx.x == 0) GC_GTR(4, 3);
}

// How the K / V rows of a step reach shared memory.  GENCAST_ATT_ROWS = cpasync (default): 16-byte cp.async copies by
// four gather warps; tma: TMA row gathers (tile::gather4) for both streams, one warp each; split: K by TMA, V by cp.async.
// Measured on B200 at 1 deg x 4 members (3 628 steps x 4 heads): 115.6 / 117.7 / 121.2 us - the same within noise.  A
// gather4 costs its issuing warp ~80 clk (32 of them = one K tile block the loader for ~2 500 clk), i.e. the TMA unit
// takes ~20 clk per 128-byte row; either way a 16 KB tile of 256-byte row pieces arrives ~3 000 clk after it is
// requested and the kernel moves 475 MB of such pieces per launch (4 TB/s out of L2): that traffic is the floor.
int rows_mode() {
  static const int mode = []() {
    const char* v = getenv("GENCAST_ATT_ROWS");
    if (v != nullptr && v[0] == 't') return 1;
    if (v != nullptr && v[0] == 's') return 2;
    return 0;
  }();
  return mode;
}

template <int D, int ROWS>
int launch_gather(cudaStream_t st, const CUtensorMap& map, const CUtensorMap& row_map, const CUtensorMap& out_map,
                  const GatherAttParams& p, int num_q_tiles) {
  using C = GCfg<D>;
  cudaError_t e = cudaFuncSetAttribute(khop_attention_gather_kernel<D, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(khop_attention_gather_kernel)");
  GC_CHECK_CUDA(launch_kernel(khop_attention_gather_kernel<D, ROWS>, dim3(num_q_tiles * p.heads), dim3(G_THREADS), (size_t)C::SMEM,
                              st, map, row_map, out_map, p), "khop_attention_gather_kernel");
  return GC_OK;
}

}  // namespace
}  // namespace gc

// Debug hook (not part of the public header): clock-stamp buffer of at least 8 * 512 int64.
extern "C" __attribute__((visibility("default"))) void gc_debug_set_attention_gather_trace(void* ptr) {
  gc::g_attention_gather_trace = reinterpret_cast<long long*>(ptr);
}

extern "C" int gc_khop_attention_gather(void* stream, const void* qkv, int64_t ld_qkv, const int32_t* step_ptr,
                                        const int32_t* keys, const uint32_t* mask, const int32_t* work,
                                        int32_t num_q_tiles, int32_t mask_period, void* out, int64_t ldo, int64_t nodes,
                                        int32_t heads, int32_t head_dim) {
  using namespace gc;
  GC_REQUIRE(qkv && step_ptr && keys && mask && work && out, "gc_khop_attention_gather: null buffer");
  GC_REQUIRE(head_dim == 64 || head_dim == 128, "gc_khop_attention_gather: head_dim=%d (supported: 64, 128)", head_dim);
  GC_REQUIRE(heads >= 1 && ld_qkv >= 3LL * heads * head_dim && ldo >= 1LL * heads * head_dim,
             "gc_khop_attention_gather: bad sizes");
  GC_REQUIRE(aligned16(qkv) && aligned16(out) && (reinterpret_cast<uintptr_t>(mask) & 7u) == 0 && ld_qkv % 8 == 0 && ldo % 8 == 0,
             "gc_khop_attention_gather: alignment");
  GC_REQUIRE(nodes > 0 && nodes < (1LL << 31), "gc_khop_attention_gather: nodes=%lld", (long long)nodes);
  GC_REQUIRE(num_q_tiles >= 1 && static_cast<int64_t>(num_q_tiles) * 128 >= nodes && mask_period >= 0,
             "gc_khop_attention_gather: num_q_tiles=%d does not cover %lld nodes", num_q_tiles, (long long)nodes);
  CUtensorMap map;
  int rc = make_tmap_bf16_2d(&map, qkv, (uint64_t)nodes, (uint64_t)(3LL * heads * head_dim), (uint64_t)ld_qkv, 64, 128);
  if (rc != GC_OK) return rc;
  GatherAttParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); p.ld_qkv = ld_qkv;
  p.step_ptr = step_ptr; p.keys = keys; p.mask = reinterpret_cast<const uint2*>(mask); p.work = work;
  p.mask_period = mask_period;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.nodes = (int)nodes; p.heads = heads;
  p.hd = heads * head_dim;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)head_dim);
  p.trace = g_attention_gather_trace;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap out_map;      // {64 columns, 32 rows}: one warp's share of a 64-column block of the output tile
  rc = make_tmap_bf16_2d(&out_map, out, (uint64_t)nodes, (uint64_t)(1LL * heads * head_dim), (uint64_t)ldo, 64, 32);
  if (rc != GC_OK) return rc;
  const int mode = rows_mode();
  CUtensorMap row_map = map;
  if (mode != 0) {           // one row of 64 columns per box: the unit of a gather4 request
    rc = make_tmap_bf16_2d(&row_map, qkv, (uint64_t)nodes, (uint64_t)(3LL * heads * head_dim), (uint64_t)ld_qkv, 64, 1);
    if (rc != GC_OK) return rc;
  }
  if (head_dim == 64) {
    if (mode == 0) return launch_gather<64, 0>(st, map, row_map, out_map, p, num_q_tiles);
    if (mode == 1) return launch_gather<64, 1>(st, map, row_map, out_map, p, num_q_tiles);
    return launch_gather<64, 2>(st, map, row_map, out_map, p, num_q_tiles);
  }
  if (mode == 0) return launch_gather<128, 0>(st, map, row_map, out_map, p, num_q_tiles);
  if (mode == 1) return launch_gather<128, 1>(st, map, row_map, out_map, p, num_q_tiles);
  return launch_gather<128, 2>(st, map, row_map, out_map, p, num_q_tiles);
}
