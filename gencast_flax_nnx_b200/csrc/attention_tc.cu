// Block-sparse k-hop attention on tcgen05 tensor cores (bf16 in, fp32 accumulate).
//
// Work unit: one CTA = (128-query tile, head).  The k-hop pattern is given as a list
// of non-empty 128 x 128 (query tile, key tile) pairs with a 128 x 128 bit mask each
// (built once on the host from adj^k, gencast/transformer.py:21-47 /
// gencast/sparse_transformer.py:555).  For each listed key tile:
//     S = Q K^T          tcgen05.mma  M=128 N=128 K=d      (operands TMA-staged, 128B swizzle)
//     P = exp2(S - max)  softmax warps read S from TMEM, apply the bit mask, write bf16 P
//                        to shared memory in the K-major swizzled operand layout
//     O += P V           tcgen05.mma  M=128 N=d K=128      (V consumed MN-major, no transpose)
// Softmax is exact and two-pass: pass 1 sweeps the key tiles for the masked row maximum
// (S only), pass 2 recomputes S, exponentiates against the final maximum and accumulates
// O in TMEM, so O is never rescaled.  Masked logits contribute exactly 0, which is what
// the reference's where(mask, logits, -1e30) + softmax evaluates to
// (gencast/sparse_transformer.py:100-125, :340-347).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
// warps 2-9 softmax / epilogue: two warps per TMEM lane quarter (warp % 4), each owning
// one 64-key half of every S tile, so each SM sub-partition has two warps to interleave.
// TMEM: S double buffer (2 x 128 columns) + O (d columns).
#include "common.cuh"
#include "sm100.cuh"

namespace gc {

namespace {

constexpr int TQ = 128;          // queries per tile
constexpr int TK = 128;          // keys per tile
constexpr int ATT_THREADS = 320;   // TMA warp, MMA warp, 8 softmax warps

template <int D>
struct AttCfg {
  static constexpr int CHUNKS = D / 64;                  // 64-element (128 B) operand chunks along d
  static constexpr int SLOT_BYTES = TK * D * 2;          // one K or V tile
  static constexpr int Q_BYTES = TQ * D * 2;
  static constexpr int P_BYTES = TQ * TK * 2;            // 32 KB, two 64-key chunks
  static constexpr int NPBUF = D == 64 ? 2 : 1;
  static constexpr int NSLOT = D == 64 ? 6 : 4;
  static constexpr int SMEM = Q_BYTES + NSLOT * SLOT_BYTES + NPBUF * P_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*exchange*/;
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
};

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, 1)
khop_attention_tc_kernel(const __grid_constant__ CUtensorMap qkv_map, const AttParams p) {
  using namespace sm100;
  using C = AttCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;
  const uint32_t slot_smem = q_smem + C::Q_BYTES;
  const uint32_t p_smem = slot_smem + C::NSLOT * C::SLOT_BYTES;
  const uint32_t bars = p_smem + C::NPBUF * C::P_BYTES;
  // barrier block
  const uint32_t q_full = bars;
  auto slot_full = [&](int s) { return bars + 8u * (1 + s); };
  auto slot_empty = [&](int s) { return bars + 8u * (1 + C::NSLOT + s); };
  const uint32_t b2 = bars + 8u * (1 + 2 * C::NSLOT);
  auto s_full = [&](int b) { return b2 + 8u * b; };
  auto s_free = [&](int b) { return b2 + 8u * (2 + b); };
  auto p_full = [&](int b) { return b2 + 8u * (4 + b); };
  auto p_empty = [&](int b) { return b2 + 8u * (6 + b); };
  const uint32_t o_full = b2 + 8u * 8;
  const uint32_t tmem_ptr_smem = b2 + 8u * 9;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  const int qt = blockIdx.x / p.heads;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&qkv_map);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::NSLOT; ++s) { mbar_init(slot_full(s), 1); mbar_init(slot_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1); mbar_init(s_free(b), 8);
      mbar_init(p_full(b), 8); mbar_init(p_empty(b), 1);
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
  pdl_launch_dependents();
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
      auto load_tile = [&](int col, int kv) {
        mbar_wait(slot_empty(slot), phase ^ 1u);
        mbar_arrive_expect_tx(slot_full(slot), C::SLOT_BYTES);
        for (int c = 0; c < C::CHUNKS; ++c)
          tma_load_2d(slot_smem + slot * C::SLOT_BYTES + c * (TK * 128), &qkv_map, slot_full(slot), col + 64 * c, kv * TK);
        if (++slot == C::NSLOT) { slot = 0; phase ^= 1u; }
      };
      for (int t = 0; t < T; ++t) load_tile(k_col, __ldg(p.tile_kv + t_beg + t));          // pass 1: K_0 .. K_{T-1}
      // pass 2 consumption order: K_0, K_1, V_0, K_2, V_1, ..., K_{T-1}, V_{T-2}, V_{T-1}
      for (int t = 0; t < T; ++t) {
        if (t == 0) load_tile(k_col, __ldg(p.tile_kv + t_beg));
        if (t + 1 < T) load_tile(k_col, __ldg(p.tile_kv + t_beg + t + 1));
        load_tile(v_col, __ldg(p.tile_kv + t_beg + t));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      constexpr uint32_t idesc_s = idesc_bf16_f32(TQ, TK, 0, 0);
      constexpr uint32_t idesc_o = idesc_bf16_f32(TQ, D, 0, 1);     // B = V tile, MN-major
      int slot = 0;
      uint32_t slot_phase = 0;
      int g = 0;     // S iterations issued so far (buffer g & 1, use count g >> 1)
      auto issue_s = [&]() {
        const int b = g & 1;
        mbar_wait(slot_full(slot), slot_phase);
        mbar_wait(s_free(b), ((g >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t k_base = slot_smem + slot * C::SLOT_BYTES;
#pragma unroll
        for (int j = 0; j < D / 16; ++j) {
          const uint32_t off = (j >> 2) * (TQ * 128) + (j & 3) * 32;
          umma_f16(tmem_s0 + b * 128, desc_kmajor_sw128(q_smem + off), desc_kmajor_sw128(k_base + off), idesc_s, j > 0);
        }
        umma_commit(slot_empty(slot));
        umma_commit(s_full(b));
        if (++slot == C::NSLOT) { slot = 0; slot_phase ^= 1u; }
        ++g;
      };
      mbar_wait(q_full, 0);
      for (int t = 0; t < T; ++t) issue_s();                       // pass 1
      for (int t = 0; t < T; ++t) {                                // pass 2
        if (t == 0) issue_s();
        if (t + 1 < T) issue_s();
        const int pb = t % C::NPBUF;
        mbar_wait(slot_full(slot), slot_phase);
        mbar_wait(p_full(pb), (t / C::NPBUF) & 1);
        tc_fence_after();
        const uint32_t v_base = slot_smem + slot * C::SLOT_BYTES;
        const uint32_t p_base = p_smem + pb * C::P_BYTES;
#pragma unroll
        for (int j = 0; j < TK / 16; ++j) {
          // A: P[128 x 16 keys], K-major, 64-key chunks of 16 KB.  B: V[16 keys x D], MN-major:
          // 16 key rows of 128 B start at j * 2048; 64-wide d chunks are TK * 128 bytes apart.
          const uint64_t da = desc_kmajor_sw128(p_base + (j >> 2) * (TQ * 128) + (j & 3) * 32);
          const uint64_t db = desc_mnmajor_sw128(v_base + j * 2048, TK * 128, 1024);
          umma_f16(tmem_o, da, db, idesc_o, (t > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(slot_empty(slot));
        umma_commit(p_empty(pb));
        if (++slot == C::NSLOT) { slot = 0; slot_phase ^= 1u; }
      }
      umma_commit(o_full);
    }
  } else {
    // ---------------- softmax + epilogue warps
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;                  // which 64-key half of each S tile this warp owns
    const int r = q * 32 + lane;                       // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem_raw + (bars - smem_u32(smem_raw)) + 8 * 32);   // [2][128] exchange
    pdl_wait();
    int g = 0;
    // Each thread walks its half row in two 32-column steps; the tcgen05.ld of the next step is in
    // flight while the current one is reduced, and reductions use independent accumulators.
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int t = 0; t < T; ++t, ++g) {                 // pass 1: masked row maximum
      const int b = g & 1;
      const uint2 mk = __ldg(reinterpret_cast<const uint2*>(p.tile_mask + static_cast<int64_t>(t_beg + t) * TQ + r) + half);
      const uint32_t mw[2] = {mk.x, mk.y};
      mbar_wait(s_full(b), (g >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_s0 + b * 128 + half * 64 + lane_addr;
      uint32_t v[32];
      tmem_ld_32x32b_x32(s_addr, v);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float f[32];
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (c < 1) tmem_ld_32x32b_x32(s_addr + 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          mx[i & 3] = fmaxf(mx[i & 3], (mw[c] & (1u << i)) ? f[i] : -INFINITY);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(b));
    }
    float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    // combine the two halves of each row (named barrier 1 over the 256 softmax threads)
    xch[half * 128 + r] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(m, xch[(half ^ 1) * 128 + r]);
    const float m_scaled = (m == -INFINITY) ? 0.0f : m * p.scale_log2e;
    float ls[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int t = 0; t < T; ++t, ++g) {                 // pass 2: P = exp2(S * c - max * c), row sums
      const int b = g & 1;
      const int pb = t % C::NPBUF;
      const uint2 mk = __ldg(reinterpret_cast<const uint2*>(p.tile_mask + static_cast<int64_t>(t_beg + t) * TQ + r) + half);
      const uint32_t mw[2] = {mk.x, mk.y};
      mbar_wait(s_full(b), (g >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_s0 + b * 128 + half * 64 + lane_addr;
      uint32_t v[32];
      tmem_ld_32x32b_x32(s_addr, v);
      if (t >= C::NPBUF) mbar_wait(p_empty(pb), ((t / C::NPBUF) - 1) & 1);
      // this warp's keys are the 64-key operand chunk `half` of P
      const uint32_t chunk_base = p_smem + pb * C::P_BYTES + half * (TQ * 128) + r * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float f[32];
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (c < 1) tmem_ld_32x32b_x32(s_addr + 32, v);
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0, e1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(f[i], p.scale_log2e, -m_scaled)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(f[i + 1], p.scale_log2e, -m_scaled)));
          const float p0 = (mw[c] & (1u << i)) ? e0 : 0.0f;
          const float p1 = (mw[c] & (1u << (i + 1))) ? e1 : 0.0f;
          ls[(i >> 1) & 3] += p0 + p1;
          const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
          packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
        // keys 32c .. 32c+31 of the chunk -> 16-byte units c * 4 + u, swizzled by row
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t unit = static_cast<uint32_t>(c * 4 + u) ^ static_cast<uint32_t>(r & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(chunk_base + unit * 16), "r"(packed[4 * u]),
                       "r"(packed[4 * u + 1]), "r"(packed[4 * u + 2]), "r"(packed[4 * u + 3])
                       : "memory");
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();       // generic-proxy stores of P -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_free(b));
        mbar_arrive(p_full(pb));
      }
    }
    float l = (ls[0] + ls[1]) + (ls[2] + ls[3]);
    asm volatile("bar.sync 1, 256;" ::: "memory");     // everyone has read the pass-1 exchange
    xch[half * 128 + r] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xch[(half ^ 1) * 128 + r];
    // epilogue: O / l -> bf16 -> global; each warp stores its half of the head's channels
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = l > 0.0f ? 1.0f / l : 0.0f;
    const int64_t row = static_cast<int64_t>(qt) * TQ + r;
#pragma unroll
    for (int c = half * (D / 2); c < (half + 1) * (D / 2); c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_o + lane_addr + c, v);
      tc_wait_ld();
      if (row < p.nodes && T > 0) {
        __nv_bfloat16* dst = p.out + row * p.ldo + head * D + c;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            h[j] = __floats2bfloat162_rn(__uint_as_float(v[i + 2 * j]) * inv_l, __uint_as_float(v[i + 2 * j + 1]) * inv_l);
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
  const int num_q_tiles = (int)((nodes + 127) / 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (head_dim == 64) return launch_tc<64>(st, map, p, num_q_tiles);
  return launch_tc<128>(st, map, p, num_q_tiles);
}
