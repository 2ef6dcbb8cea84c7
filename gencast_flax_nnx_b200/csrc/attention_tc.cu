// Block-sparse k-hop attention on tcgen05 tensor cores (bf16 in, fp32 accumulate).
//
// Work unit: one CTA = (128-query tile, head).  The k-hop pattern is given as a list
// of non-empty 128 x 128 (query tile, key tile) pairs with a 128 x 128 bit mask each
// (built once on the host from adj^k, gencast/transformer.py:21-47 /
// gencast/sparse_transformer.py:555).  For each listed key tile:
//     S = Q K^T          tcgen05.mma  M=128 N=128 K=d      (operands TMA-staged, 128B swizzle)
//     P = exp2(S - max)  softmax warps read S from TMEM, apply the bit mask and write bf16 P
//                        back into TENSOR MEMORY, over the columns of the S tile they came from
//     O += P V           tcgen05.mma  M=128 N=d K=128, A = P read from tensor memory (TS form),
//                        B = V consumed MN-major straight from its TMA tile
// Softmax is exact, single pass and online: every row keeps a running offset m and sum l;
// a tile's masked maximum only replaces m when it exceeds it by more than 2^8 (then l and
// the row of O in TMEM are rescaled by the softmax warps themselves, which is rare after
// the first tile), otherwise exponentials are taken against the stale offset, which is
// exact after the final division by l.  Masked logits contribute exactly 0, which is what
// the reference's where(mask, logits, -1e30) + softmax evaluates to
// (gencast/sparse_transformer.py:100-125, :340-347).
//
// Why P lives in tensor memory: with 128 x 128 tiles every MMA instruction reads 4 KB of A and
// 4 KB of B from shared memory per 64 clk, i.e. the full 128 B/clk of the SM; with P in shared
// memory the clock-stamp trace showed 110 clk per MMA instead of 64 (tile period 2 000 clk for
// 1 024 clk of tensor work), the MMA-issuing thread being the critical path.  Reading P from TMEM
// halves the shared-memory traffic of the PV product and removes the P stores and their
// proxy fences; the 64 KB of P buffers become two more K/V slots.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
// warps 2-9 softmax / epilogue in two groups of four (one warp per TMEM lane quarter each):
// group g owns the key tiles t = g, g+2, ... with its own S/P columns, O accumulator and (m, l)
// state, so the two groups are independent online-softmax streams that are merged once at the
// end (O = (a0 O_0 + a1 O_1) / (a0 l_0 + a1 l_1), a_g = 2^(m_g - m)).  The tensor pipe executes
// MMAs in issue order, so S(t+2), issued right after PV(t), may overwrite the columns P(t) was
// read from without any further synchronisation, and its completion implies that of PV(t-2):
// the group may rescale O_g as soon as S(t) has arrived.
// 32 x 32 sub-blocks of a tile whose mask bits are all zero (about half of them with the
// hierarchical patch ordering of the mesh) are neither read nor exponentiated.
// TMEM: S_0 / P_0, S_1 / P_1 (2 x 128 columns, P in the first 64) + O_0, O_1 (2 x d columns).
#include "common.cuh"
#include "sm100.cuh"

namespace gc {

long long* g_attention_trace = nullptr;

namespace {

constexpr int TQ = 128;          // queries per tile
constexpr int TK = 128;          // keys per tile
constexpr int ATT_THREADS = 320;   // TMA warp, MMA warp, 8 softmax warps

template <int D>
struct AttCfg {
  static constexpr int CHUNKS = D / 64;                  // 64-element (128 B) operand chunks along d
  static constexpr int SLOT_BYTES = TK * D * 2;          // one K or V tile
  static constexpr int Q_BYTES = TQ * D * 2;
  // K / V tile ring: a K tile and the V tile listed with it are consumed back to back (PV(t), S(t+2))
  static constexpr int NSLOT = D == 64 ? 10 : 6;
  static constexpr int SMEM = Q_BYTES + NSLOT * SLOT_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 64 /*tile_kv ring*/;
};

struct AttParams {
  const int32_t* tile_ptr;    // [num_q_tiles + 1]
  const int32_t* tile_kv;     // [num_tiles] key-tile index of each listed pair
  const uint4* tile_mask;     // [num_tiles][128 rows] 128 bits: bit j of row r = key (kv*128 + j) is a neighbour
  __nv_bfloat16* out;
  int64_t ldo;
  int nodes;
  int heads;
  int hd;                     // heads * head_dim = column offset of K inside a qkv row (V at 2 * hd)
  float scale_log2e;          // head_dim^-0.5 * log2(e)
  long long* trace;           // debug: per-role clock stamps of CTA 0 (null in normal use)
};

// Debug timeline: CTA 0 records clock64() at pipeline events into trace[role * 512 + index].
#define GC_TRACE(role, index)                                                         \
  do {                                                                                \
    if (p.trace != nullptr && blockIdx.x == 0 && (index) < 512) p.trace[(role) * 512 + (index)] = clock64(); \
  } while (0)


template <int D>
__global__ void __launch_bounds__(ATT_THREADS, 1)
khop_attention_tc_kernel(const __grid_constant__ CUtensorMap qkv_map, const AttParams p) {
  using namespace sm100;
  pdl_launch_dependents();     // let the next kernel's launch and prologue overlap this one
  using C = AttCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;
  const uint32_t slot_smem = q_smem + C::Q_BYTES;
  const uint32_t bars = slot_smem + C::NSLOT * C::SLOT_BYTES;
  // barrier block
  const uint32_t q_full = bars;
  auto slot_full = [&](int s) { return bars + 8u * (1 + s); };
  auto slot_empty = [&](int s) { return bars + 8u * (1 + C::NSLOT + s); };
  const uint32_t b2 = bars + 8u * (1 + 2 * C::NSLOT);
  auto s_full = [&](int b) { return b2 + 8u * b; };
  auto p_full = [&](int b) { return b2 + 8u * (2 + b); };
  const uint32_t o_full = b2 + 8u * 4;
  const uint32_t tmem_ptr_smem = b2 + 8u * 5;
  // tile_kv entries of the tiles in flight, written by the TMA producer before it requests K_t and read by
  // the MMA issuer after the tile has landed (the mbarrier orders them): no global load on the issuer's path
  volatile uint32_t* kv_ring = reinterpret_cast<volatile uint32_t*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  const int qt = blockIdx.x / p.heads;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&qkv_map);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::NSLOT; ++s) { mbar_init(slot_full(s), 1); mbar_init(slot_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(p_full(b), 4);
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  pdl_wait();                                    // everything below may read what a predecessor wrote
  const int t_beg = __ldg(p.tile_ptr + qt);
  const int T = __ldg(p.tile_ptr + qt + 1) - t_beg;
  const uint32_t tmem_s0 = tmem_base;            // S buffers at columns 0 and 128
  const uint32_t tmem_o = tmem_base + 256;       // O at columns 256 .. 256 + D

  const int q_col = head * D;
  const int k_col = p.hd + head * D;
  const int v_col = 2 * p.hd + head * D;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: loads in exactly the order the MMA warp consumes them
      pdl_wait();
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
      for (int c = 0; c < C::CHUNKS; ++c) tma_load_2d(q_smem + c * (TQ * 128), &qkv_map, q_full, q_col + 64 * c, qt * TQ);
      int slot = 0;
      uint32_t phase = 0;
      int tma_ev = 0;
      auto load_tile = [&](int col, int kv) {
        mbar_wait(slot_empty(slot), phase ^ 1u);
        GC_TRACE(0, tma_ev); ++tma_ev;
        mbar_arrive_expect_tx(slot_full(slot), C::SLOT_BYTES);
        for (int c = 0; c < C::CHUNKS; ++c)
          tma_load_2d(slot_smem + slot * C::SLOT_BYTES + c * (TK * 128), &qkv_map, slot_full(slot), col + 64 * c, kv * TK);
        if (++slot == C::NSLOT) { slot = 0; phase ^= 1u; }
      };
      // consumption order: K_0, K_1, then per tile t: V_t, K_{t+2}
      auto load_k = [&](int t) {
        const uint32_t kv = static_cast<uint32_t>(__ldg(p.tile_kv + t_beg + t));
        kv_ring[t & 15] = kv;
        load_tile(k_col, static_cast<int>(kv & 0xffffffu));
      };
      for (int t = 0; t < 2 && t < T; ++t) load_k(t);
      for (int t = 0; t < T; ++t) {
        load_tile(v_col, static_cast<int>(kv_ring[t & 15] & 0xffffffu));
        if (t + 2 < T) load_k(t + 2);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      constexpr uint32_t idesc_o = idesc_bf16_f32(TQ, D, 0, 1);     // B = V tile, MN-major
      // live key range of a pair, in 32-key sub-blocks: [lo, lo + cnt)
      auto key_range = [&](int t, int& lo, int& cnt) {       // call after K_t has landed
        const uint32_t kv = kv_ring[t & 15];
        lo = static_cast<int>((kv >> 24) & 3u);
        cnt = 4 - lo - static_cast<int>((kv >> 26) & 3u);
      };
      int slot = 0;
      uint32_t slot_phase = 0;
      int g = 0;     // S iterations issued so far (buffer g & 1, use count g >> 1)
      int mma_ev = 0;
      auto issue_s = [&]() {
        const int b = g & 1;
        mbar_wait(slot_full(slot), slot_phase);
        int lo, cnt;
        key_range(g, lo, cnt);
        // N = 32 cnt keys starting at key 32 lo: K rows 32 lo .. of the tile (4 KB per 32 rows and 64-wide d
        // chunk), S columns 32 lo .. of the group's buffer, so that column = key as the softmax expects
        const uint32_t idesc_s = idesc_bf16_f32(TQ, 32 * cnt, 0, 0);
        GC_TRACE(1, 2 * mma_ev);
        GC_TRACE(1, 2 * mma_ev + 1); ++mma_ev;
        tc_fence_after();
        const uint32_t k_base = slot_smem + slot * C::SLOT_BYTES + static_cast<uint32_t>(lo) * 4096u;
#pragma unroll
        for (int j = 0; j < D / 16; ++j) {
          const uint32_t off = (j >> 2) * (TQ * 128) + (j & 3) * 32;
          umma_f16(tmem_s0 + b * 128 + 32 * lo, desc_kmajor_sw128(q_smem + off), desc_kmajor_sw128(k_base + off), idesc_s, j > 0);
        }
        umma_commit(slot_empty(slot));
        umma_commit(s_full(b));
        if (++slot == C::NSLOT) { slot = 0; slot_phase ^= 1u; }
        ++g;
      };
      mbar_wait(q_full, 0);
      for (int t = 0; t < 2 && t < T; ++t) issue_s();
      for (int t = 0; t < T; ++t) {
        const int pb = t & 1;
        mbar_wait(slot_full(slot), slot_phase);
        GC_TRACE(2, 2 * t);
        mbar_wait(p_full(pb), (t >> 1) & 1);
        GC_TRACE(2, 2 * t + 1);
        tc_fence_after();
        const uint32_t v_base = slot_smem + slot * C::SLOT_BYTES;
        int lo, cnt;
        key_range(t, lo, cnt);
        // two 16-key MMA steps per live 32-key sub-block
        for (int j = 2 * lo; j < 2 * (lo + cnt); ++j) {
          // A: P[128 x 16 keys] from tensor memory: lane = query row, 8 columns of two bf16 each.
          // B: V[16 keys x D], MN-major: 16 key rows of 128 B start at j * 2048; 64-wide d chunks are
          // TK * 128 bytes apart.
          const uint64_t db = desc_mnmajor_sw128(v_base + j * 2048, TK * 128, 1024);
          umma_f16_ts(tmem_o + (t & 1) * D, tmem_s0 + pb * 128 + 8 * j, db, idesc_o, (t > 1 || j > 2 * lo) ? 1u : 0u);
        }
        umma_commit(slot_empty(slot));
        if (++slot == C::NSLOT) { slot = 0; slot_phase ^= 1u; }
        // S(t+2) overwrites the columns of S(t) / P(t): in issue order behind PV(t), which read them
        if (t + 2 < T) issue_s();
      }
      umma_commit(o_full);
    }
  } else {
    // ---------------- softmax + epilogue warps
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;                   // softmax group = parity of the key tiles it owns
    const int r = q * 32 + lane;                       // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    // [2 groups][m, l][128] floats, in the Q tile's shared memory: every S MMA (the only reader of Q) has
    // completed before any softmax warp leaves its tile loop
    float* xch = reinterpret_cast<float*>(smem_raw + (q_smem - smem_u32(smem_raw)));
    const float c2 = p.scale_log2e;
    const uint32_t s_addr = tmem_s0 + grp * 128 + lane_addr;
    const uint32_t o_addr = tmem_o + grp * D + lane_addr;
    float m = -INFINITY;                               // running offset (raw logit units)
    float ls[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint4 mk_next = make_uint4(0u, 0u, 0u, 0u);
    if (grp < T) mk_next = __ldg(p.tile_mask + static_cast<int64_t>(t_beg + grp) * TQ + r);
    for (int t = grp; t < T; t += 2) {
      const int use = t >> 1;                          // how many times this group's buffers were used before
      const uint4 mk = mk_next;                        // mask rows are fetched one tile ahead
      if (t + 2 < T) mk_next = __ldg(p.tile_mask + static_cast<int64_t>(t_beg + t + 2) * TQ + r);
      const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
      // 32 x 32 sub-blocks without any neighbour are skipped (warp-uniform)
      bool live[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) live[c] = __any_sync(0xffffffffu, mw[c] != 0u);
      if (threadIdx.x == 64) GC_TRACE(3, 2 * t);
      mbar_wait(s_full(grp), use & 1);
      if (threadIdx.x == 64) GC_TRACE(3, 2 * t + 1);
      tc_fence_after();
      // (a) masked maximum of the row over this tile; the tcgen05.ld of the next sub-block is in
      // flight while the current one is reduced
      // Tensor memory is read at 64 B/clk per SM: a 128 x 128 fp32 S tile costs 1024 clk per pass,
      // as much as both MMAs of the tile together, so sub-blocks without a neighbour are not read at
      // all (neither here nor in the exponential pass).
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (live[c]) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(s_addr + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              mx[i & 3] = fmaxf(mx[i & 3], (mw[c] & (1u << i)) ? __uint_as_float(v[i]) : -INFINITY);
          }
        }
      }
      const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      const float m_new = fmaxf(m, mt);
      // O_grp is at rest: S(t) was issued behind PV(t-2) and has completed
      if (threadIdx.x == 64) GC_TRACE(4, t);
      if (use == 0) {
        m = m_new;                                     // nothing accumulated yet
      } else {
        // raise the offset only when the stale one would let exponentials grow past 2^8
        // (a row that has not met a neighbour yet, m = -inf, has l = 0 and a zero row of O: it just
        // adopts the new offset)
        if (m == -INFINITY) m = m_new;
        const bool need = m_new * c2 > m * c2 + 8.0f;      // false when m_new == m
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? exp2f(m * c2 - m_new * c2) : 1.0f;
          if (need) {
            m = m_new;
#pragma unroll
            for (int i = 0; i < 4; ++i) ls[i] *= alpha;
          }
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
      const float offset = (m == -INFINITY) ? 0.0f : m * c2;
      // (b) P = exp2(S c - offset) on the live sub-blocks, zeros elsewhere -> swizzled K-major operand
      {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t packed[16];
          if (live[c]) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(s_addr + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float e0, e1;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v[i]), c2, -offset)));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v[i + 1]), c2, -offset)));
              const float p0 = (mw[c] & (1u << i)) ? e0 : 0.0f;
              const float p1 = (mw[c] & (1u << (i + 1))) ? e1 : 0.0f;
              ls[(i >> 1) & 3] += p0 + p1;
              const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
              packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = 0u;
          }
          // keys 32c .. 32c+31 of this row -> columns 16c .. 16c+15 of the tile's TMEM region: over S
          // columns that have been read already (chunk c of S sits at columns 32c .. 32c+31)
          tmem_st_32x32b_x16(s_addr + c * 16, packed);
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(grp));
    }
    // ---- merge the two streams: every thread publishes (m, l) of its row
    const float l_own = (ls[0] + ls[1]) + (ls[2] + ls[3]);
    // both groups are past their last S tile: no S MMA is left that reads Q, its memory can be reused
    asm volatile("bar.sync 1, 256;" ::: "memory");
    xch[grp * 256 + r] = m;
    xch[grp * 256 + 128 + r] = l_own;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float m0 = xch[r], l0 = xch[128 + r], m1 = xch[256 + r], l1 = xch[256 + 128 + r];
    const float mm = fmaxf(m0, m1);
    const float a0 = (m0 == -INFINITY) ? 0.0f : exp2f((m0 - mm) * c2);
    const float a1 = (m1 == -INFINITY) ? 0.0f : exp2f((m1 - mm) * c2);
    const float l = l0 * a0 + l1 * a1;
    const bool two = T > 1;                            // O_1 is only defined when group 1 had a tile
    // epilogue: group g stores channel half g of the head
    if (threadIdx.x == 64) GC_TRACE(5, 0);
    mbar_wait(o_full, 0);
    if (threadIdx.x == 64) GC_TRACE(5, 1);
    tc_fence_after();
    const float inv_l = l > 0.0f ? 1.0f / l : 0.0f;
    const float w0 = a0 * inv_l, w1 = a1 * inv_l;
    const int64_t row = static_cast<int64_t>(qt) * TQ + r;
#pragma unroll
    for (int c = grp * (D / 2); c < (grp + 1) * (D / 2); c += 32) {
      uint32_t v[32], u[32];
      tmem_ld_32x32b_x32(tmem_o + lane_addr + c, v);
      if (two) tmem_ld_32x32b_x32(tmem_o + D + lane_addr + c, u);
      tc_wait_ld();
      if (row < p.nodes && T > 0) {
        __nv_bfloat16* dst = p.out + row * p.ldo + head * D + c;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x0 = __uint_as_float(v[i + 2 * j]) * w0, x1 = __uint_as_float(v[i + 2 * j + 1]) * w0;
            if (two) {
              x0 = fmaf(__uint_as_float(u[i + 2 * j]), w1, x0);
              x1 = fmaf(__uint_as_float(u[i + 2 * j + 1]), w1, x1);
            }
            h[j] = __floats2bfloat162_rn(x0, x1);
          }
          *reinterpret_cast<uint4*>(dst + i) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int D>
int launch_tc(cudaStream_t st, const CUtensorMap& map, const AttParams& p, int num_q_tiles) {
  using C = AttCfg<D>;
  cudaError_t e = cudaFuncSetAttribute(khop_attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(khop_attention_tc_kernel)");
  GC_CHECK_CUDA(launch_kernel(khop_attention_tc_kernel<D>, dim3(num_q_tiles * p.heads), dim3(ATT_THREADS), (size_t)C::SMEM, st,
                              map, p), "khop_attention_tc_kernel");
  return GC_OK;
}

}  // namespace
}  // namespace gc

// Debug hook (not part of the public header): clock-stamp buffer of at least 8 * 512 int64.
extern "C" __attribute__((visibility("default"))) void gc_debug_set_attention_trace(void* ptr) {
  gc::g_attention_trace = reinterpret_cast<long long*>(ptr);
}

extern "C" int gc_khop_attention_tiles(void* stream, const void* qkv, int64_t ld_qkv, const int32_t* tile_ptr,
                                       const int32_t* tile_kv, const uint32_t* tile_mask, void* out, int64_t ldo,
                                       int64_t nodes, int32_t heads, int32_t head_dim) {
  using namespace gc;
  GC_REQUIRE(qkv && tile_ptr && tile_kv && tile_mask && out, "gc_khop_attention_tiles: null buffer");
  GC_REQUIRE(head_dim == 64 || head_dim == 128, "gc_khop_attention_tiles: head_dim=%d (supported: 64, 128)", head_dim);
  GC_REQUIRE(heads >= 1 && ld_qkv >= 3LL * heads * head_dim && ldo >= 1LL * heads * head_dim,
             "gc_khop_attention_tiles: bad sizes");
  GC_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(tile_mask) && ld_qkv % 8 == 0 && ldo % 8 == 0,
             "gc_khop_attention_tiles: alignment");
  GC_REQUIRE(nodes > 0 && nodes < (1LL << 31), "gc_khop_attention_tiles: nodes=%lld", (long long)nodes);
  CUtensorMap map;
  int rc = make_tmap_bf16_2d(&map, qkv, (uint64_t)nodes, (uint64_t)(3LL * heads * head_dim), (uint64_t)ld_qkv, 64, 128);
  if (rc != GC_OK) return rc;
  AttParams p;
  p.tile_ptr = tile_ptr; p.tile_kv = tile_kv; p.tile_mask = reinterpret_cast<const uint4*>(tile_mask);
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.nodes = (int)nodes; p.heads = heads;
  p.hd = heads * head_dim;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)head_dim);
  p.trace = g_attention_trace;
  const int num_q_tiles = (int)((nodes + 127) / 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (head_dim == 64) return launch_tc<64>(st, map, p, num_q_tiles);
  return launch_tc<128>(st, map, p, num_q_tiles);
}
