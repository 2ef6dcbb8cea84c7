// HBM-bound row-wise kernels: LayerNorm + conditional affine, the deterministic
// receiver-sorted segment sum, noise-level conditioning tables, affine folding,
// the DPM-Solver++ 2S update, layout glue and ensemble accumulation.
//
// All of them move each byte once with 128-bit accesses, one warp per row, fp32
// statistics; grids are sized in multiples of the SM count for the large inputs.
#include "common.cuh"

namespace gc {

namespace {

constexpr float LN_EPS = 1e-6f;  // flax.nnx.LayerNorm epsilon used by the reference (common/mlp.py:95-103)

// Each lane owns NV = cols/32 elements of the row, as NV/4 chunks of 4 interleaved
// across the warp so that every access is a fully coalesced 512 B (f32) or 256 B
// (bf16) request.
template <int NV>
__device__ __forceinline__ void load_row(const void* base, int dtype, int64_t row_off, int lane, float (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV / 4; ++j) {
    float t[4];
    load_as_float<4>(base, dtype, row_off + (j * 32 + lane) * 4, t);
    v[j * 4] = t[0]; v[j * 4 + 1] = t[1]; v[j * 4 + 2] = t[2]; v[j * 4 + 3] = t[3];
  }
}

template <int NV>
__device__ __forceinline__ void store_row(void* base, int dtype, int64_t row_off, int lane, const float (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV / 4; ++j) {
    float t[4] = {v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]};
    store_from_float<4>(base, dtype, row_off + (j * 32 + lane) * 4, t);
  }
}

// LayerNorm statistics exactly as flax's fast-variance path: var = E[x^2] - E[x]^2, clipped at 0.
template <int NV>
__device__ __forceinline__ void layer_norm_inplace(float (&v)[NV], int cols) {
  float s = 0.0f, ss = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { s += v[i]; ss = fmaf(v[i], v[i], ss); }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float inv_n = 1.0f / static_cast<float>(cols);
  const float mean = s * inv_n;
  const float var = fmaxf(ss * inv_n - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + LN_EPS);
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = (v[i] - mean) * rstd;
}

// RES: a residual row is added; it is then requested together with the input row (one load latency per row instead of
// two) at the price of 18 more registers, which the variant without residual does not pay.
template <int NV, bool RES>
__global__ void __launch_bounds__(256) ln_cond_kernel(const void* __restrict__ x, int x_dtype, int64_t ldx,
                                                      const float* __restrict__ scale_offset, int do_ln,
                                                      const void* __restrict__ residual, int res_dtype, int64_t ld_res,
                                                      void* __restrict__ out, int out_dtype, int64_t ldo,
                                                      int64_t rows) {
  constexpr int cols = NV * 32;
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  // (1 + s | o) live in shared memory, not in 2 x NV registers per lane: the kernel is bound by the row bytes each SM keeps
  // in flight, i.e. by resident warps (78 -> 46 registers: 3 -> 5 blocks per SM at L = 512)
  __shared__ __align__(16) float so_s[2 * cols];
  pdl_wait();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x)
    so_s[c] = scale_offset != nullptr ? __ldg(scale_offset + c) : (c < cols ? 1.0f : 0.0f);
  __syncthreads();
  for (int64_t row = warp0; row < rows; row += nwarps) {
    float v[NV], r[NV];
    load_row<NV>(x, x_dtype, row * ldx, lane, v);
    if constexpr (RES) load_row<NV>(residual, res_dtype, row * ld_res, lane, r);
    if (do_ln) layer_norm_inplace<NV>(v, cols);
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const float4 sc = *reinterpret_cast<const float4*>(&so_s[(j * 32 + lane) * 4]);
      const float4 of = *reinterpret_cast<const float4*>(&so_s[cols + (j * 32 + lane) * 4]);
      v[j * 4] = fmaf(v[j * 4], sc.x, of.x); v[j * 4 + 1] = fmaf(v[j * 4 + 1], sc.y, of.y);
      v[j * 4 + 2] = fmaf(v[j * 4 + 2], sc.z, of.z); v[j * 4 + 3] = fmaf(v[j * 4 + 3], sc.w, of.w);
    }
    if constexpr (RES) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += r[i];
    }
    store_row<NV>(out, out_dtype, row * ldo, lane, v);
  }
}

// One warp per receiver; a block owns the receivers blockIdx.x + k * gridDim.x (strided: high
// in-degree receivers -- the mesh nodes around the poles, SURVEY.md Appendix A -- cluster in index
// space, striding spreads them over all blocks), warp w takes k = w, w + 8, ...
// Receivers whose in-degree exceeds HEAVY are only listed in that pass and afterwards summed by
// the whole block (each warp a contiguous slice of the edge range, partial sums combined in warp
// order): hundreds to thousands of edges on one warp would be a serial tail.  Summation order is
// a pure function of (row_ptr, edge_perm, grid size), so results are bitwise reproducible.
//
// The kernel is bound by memory latency, not by instructions: what matters is how many row bytes
// each SM keeps in flight.  Hence
//  * rows stay in registers as loaded (bf16 pairs: 8 registers per 512-wide row and lane) until
//    they are consumed, 16-byte loads;
//  * while a warp reduces the rows of receiver k it already has the first SEG_UNROLL rows of its
//    next receiver (and that receiver's row_ptr entries) in flight, so the memory pipe is fed
//    during the LayerNorm shuffles as well;
//  * the conditional affine is applied once per receiver instead of once per edge:
//        sum_e (LN(y_e) (1 + s) + o) = (1 + s) sum_e LN(y_e) + deg o
//    (same value up to fp32 rounding).
constexpr int SEG_WARPS = 8;
constexpr int HEAVY = 32;
constexpr int SEG_UNROLL = 3;
constexpr int SEG_RING = 6;                 // rows per warp in flight in the shared-memory ring variant
constexpr int SEG_HEAVY_PER_WARP = 64;   // heavy receivers a warp can defer to the block (beyond: it sums them itself)

// OCC = resident blocks per SM the register budget is cut for; the grid is one wave of exactly those blocks (receivers
// are strided over the blocks, so a partial last wave - 2.67 waves with the former 8 blocks per SM - was a pure tail and
// left the warps of a block waiting at the barrier before the heavy-receiver pass).  Regular low-degree graphs (mesh2grid: three
// rows per receiver, no permutation) run best with the full budget (2 blocks, no spills: 305 us vs 365 us
// at 1 deg x 4); irregular ones (grid2mesh: degrees 3 .. 594 through edge_perm) gain more from a third
// block of warps than they lose to a few spilled registers (205 us vs 246 us).
// STATS: the producer of y (gc_edge_mlp_rows) also wrote, per row, the sum and the sum of squares of the row's two column
// halves from its fp32 accumulator (row_stats[row] = {s0, ss0, s1, ss1}).  The row's mean and rstd then cost a broadcast
// 16-byte load instead of 32 multiply-adds and two warp reductions, and the kernel moves from instruction-bound (249
// instructions per row) towards the HBM roofline.
// RING (bf16 rows stored receiver-sorted, no edge permutation): pass 1 streams the rows of a warp's receivers through a
// per-warp ring of SEG_RING rows in shared memory filled by cp.async, so a warp always has that many rows in flight,
// whatever the in-degrees, without holding them in registers: the register-batch version exposed one load latency per
// batch of three rows (long_scoreboard 10 stall cycles per issued instruction in ncu, 2.3-2.6 TB/s).
template <int NV, bool BF16, int OCC, bool STATS = false, bool RING = false>
__global__ void __launch_bounds__(SEG_WARPS * 32, OCC) ln_cond_segment_sum_kernel(
    const void* __restrict__ y, int64_t ldy, const float* __restrict__ scale_offset, int do_ln,
    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ edge_perm, void* __restrict__ out,
    int out_dtype, int64_t ldo, int64_t num_segments, const float* __restrict__ row_stats) {
  constexpr int cols = NV * 32;
  constexpr int W = NV % 8 == 0 ? 8 : 4;            // consecutive elements per lane and chunk
  constexpr int NCH = NV / W;
  constexpr int RAWD = BF16 ? NV / 2 : NV;          // 32-bit registers of one row per lane, as loaded
  constexpr int RAW = RAWD + (STATS ? 2 : 0);       // ... followed by the row's (sum, sum of squares), the same in every lane
  __shared__ __align__(16) float partial[SEG_WARPS][cols];
  __shared__ __align__(16) float so_s[2 * cols];
  __shared__ int heavy_list[SEG_WARPS][SEG_HEAVY_PER_WARP];
  __shared__ int heavy_count[SEG_WARPS];
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float inv_n = 1.0f / static_cast<float>(cols);
  pdl_wait();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x)
    so_s[c] = scale_offset != nullptr ? __ldg(scale_offset + c) : (c < cols ? 1.0f : 0.0f);
  __syncthreads();

  auto load_raw = [&](int64_t e, uint32_t (&raw)[RAW]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int64_t off = e * ldy + (j * 32 + lane) * W;
      if constexpr (BF16) {
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(y) + off;
        if constexpr (W == 8) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
          raw[4 * j] = q.x; raw[4 * j + 1] = q.y; raw[4 * j + 2] = q.z; raw[4 * j + 3] = q.w;
        } else {
          const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
          raw[2 * j] = q.x; raw[2 * j + 1] = q.y;
        }
      } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(y) + off);
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
          const float4 q = __ldg(p + i);
          raw[W * j + 4 * i] = __float_as_uint(q.x); raw[W * j + 4 * i + 1] = __float_as_uint(q.y);
          raw[W * j + 4 * i + 2] = __float_as_uint(q.z); raw[W * j + 4 * i + 3] = __float_as_uint(q.w);
        }
      }
    }
    if constexpr (STATS) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(row_stats) + e);
      raw[RAWD] = __float_as_uint(q.x + q.z);          // halves in fixed order
      raw[RAWD + 1] = __float_as_uint(q.y + q.w);
    }
  };
  // rows [j0, min(j0 + SEG_UNROLL, end)) of the edge list -> raw
  auto load_batch = [&](int j0, int end, uint32_t (&raw)[SEG_UNROLL][RAW]) {
#pragma unroll
    for (int u = 0; u < SEG_UNROLL; ++u) {
      if (j0 + u < end) {                        // warp-uniform
        const int64_t e = edge_perm != nullptr ? __ldg(edge_perm + j0 + u) : j0 + u;
        load_raw(e, raw[u]);
      }
    }
  };
  // acc += LN(row) (or row)
  auto consume_row = [&](const uint32_t (&raw)[RAW], float (&acc)[NV]) {
    float v[NV];
    if constexpr (BF16) {
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {         // bf16 -> fp32 is a 16-bit shift
        v[2 * i] = __uint_as_float(raw[i] << 16);
        v[2 * i + 1] = __uint_as_float(raw[i] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = __uint_as_float(raw[i]);
    }
    if (do_ln) {
      float s = 0.0f, ss = 0.0f;
      if constexpr (STATS) {
        s = __uint_as_float(raw[RAWD]);
        ss = __uint_as_float(raw[RAWD + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) { s += v[i]; ss = fmaf(v[i], v[i], ss); }
        s = warp_sum(s);
        ss = warp_sum(ss);
      }
      const float mean = s * inv_n;
      const float rstd = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.0f) + LN_EPS);
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = fmaf(v[i] - mean, rstd, acc[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += v[i];
    }
  };
  // the rows of one batch, in edge order
  auto consume_batch = [&](int j0, int end, const uint32_t (&raw)[SEG_UNROLL][RAW], float (&acc)[NV]) {
#pragma unroll
    for (int u = 0; u < SEG_UNROLL; ++u) {
      if (j0 + u < end) consume_row(raw[u], acc);
    }
  };
  // remaining batches of a range whose first batch is already in `first`
  auto accumulate_rest = [&](int beg, int end, const uint32_t (&first)[SEG_UNROLL][RAW], float (&acc)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
    consume_batch(beg, end, first, acc);
    for (int j0 = beg + SEG_UNROLL; j0 < end; j0 += SEG_UNROLL) {
      uint32_t raw[SEG_UNROLL][RAW];
      load_batch(j0, end, raw);
      consume_batch(j0, end, raw, acc);
    }
  };
  // out[seg] = (1 + s) * acc + deg * o
  auto finish = [&](int64_t seg, int deg, float (&acc)[NV]) {
    const float fdeg = static_cast<float>(deg);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float t[W];
#pragma unroll
      for (int i = 0; i < W / 4; ++i) {
        const float4 sc = *reinterpret_cast<const float4*>(&so_s[(j * 32 + lane) * W + 4 * i]);
        const float4 of = *reinterpret_cast<const float4*>(&so_s[cols + (j * 32 + lane) * W + 4 * i]);
        const float* a = &acc[j * W + 4 * i];
        t[4 * i] = fmaf(a[0], sc.x, fdeg * of.x); t[4 * i + 1] = fmaf(a[1], sc.y, fdeg * of.y);
        t[4 * i + 2] = fmaf(a[2], sc.z, fdeg * of.z); t[4 * i + 3] = fmaf(a[3], sc.w, fdeg * of.w);
      }
      store_from_float<W>(out, out_dtype, seg * ldo + (j * 32 + lane) * W, t);
    }
  };

  // ---- pass 1: light receivers, one warp each.  Software pipeline over the warp's receivers
  // k, k + 8, ...: the row_ptr pair of receiver i + 2 and the first rows of receiver i + 1 are
  // requested before receiver i is reduced, so neither the index load nor the row loads sit on
  // the critical path of the in-order warp.
  const int64_t stride = static_cast<int64_t>(gridDim.x);
  auto seg_of = [&](int64_t k) { return static_cast<int64_t>(blockIdx.x) + k * stride; };
  auto load_bounds = [&](int64_t kk, int& b, int& e) {
    const int64_t sg = seg_of(kk);
    b = 0; e = -1;                                  // e < b marks "no such receiver"
    if (sg < num_segments) { b = __ldg(row_ptr + sg); e = __ldg(row_ptr + sg + 1); }
  };
  int my_heavy = 0;
  if constexpr (RING) {
    // One rolling stream of rows per warp: the rows of its receivers k, k + 8, ... concatenated.  A slot of the ring
    // carries its receiver and, on the receiver's last row, the in-degree; it is refilled as soon as it has been read.
    // Rows are added in row order per receiver, as in the register version: the same bits.
    extern __shared__ __align__(16) uint8_t ring_raw[];
    constexpr int ROWB = cols * 2;                  // bytes of a bf16 row
    constexpr int SLOTB = ROWB + 16;                // + the row's statistics
    constexpr int R = SEG_RING;
    const uint32_t ring = static_cast<uint32_t>(__cvta_generic_to_shared(ring_raw)) + static_cast<uint32_t>(warp) * (R * SLOTB);
    int64_t g_k = static_cast<int64_t>(warp) - SEG_WARPS;
    int g_row = 0, g_end = 0, g_deg = 0;
    int64_t g_seg = 0;
    bool g_done = false;
    int nb0, ne0, nb1, ne1, nb2, ne2;               // bounds of the next three receivers, requested ahead
    load_bounds(warp, nb0, ne0);
    load_bounds(warp + SEG_WARPS, nb1, ne1);
    load_bounds(warp + 2 * SEG_WARPS, nb2, ne2);
    auto next_row = [&](int& row, int64_t& seg, int& last_deg) -> bool {
      while (!g_done && g_row == g_end) {
        g_k += SEG_WARPS;
        const int b = nb0, e = ne0;
        nb0 = nb1; ne0 = ne1; nb1 = nb2; ne1 = ne2;
        load_bounds(g_k + 3 * SEG_WARPS, nb2, ne2);
        if (e < b) { g_done = true; break; }
        const int64_t sg = seg_of(g_k);
        const int deg = e - b;
        if (deg == 0) {                             // receiver without edges: (1 + s) * 0 + 0 * o
          float zero[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i) zero[i] = 0.0f;
          finish(sg, 0, zero);
          continue;
        }
        if (deg > HEAVY && my_heavy < SEG_HEAVY_PER_WARP) {
          if (lane == 0) heavy_list[warp][my_heavy] = static_cast<int>(g_k);
          ++my_heavy;
          continue;
        }
        g_row = b; g_end = e; g_seg = sg; g_deg = deg;
      }
      if (g_done) return false;
      row = g_row++;
      seg = g_seg;
      last_deg = g_row == g_end ? g_deg : 0;
      return true;
    };
    int64_t slot_seg[R];
    int slot_deg[R];                                // -1: empty; 0: inner row; > 0: last row of a receiver of that degree
    auto fill = [&](int u) {
      int row;
      slot_deg[u] = -1;
      if (next_row(row, slot_seg[u], slot_deg[u])) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(y) + static_cast<int64_t>(row) * ldy * 2;
        const uint32_t dst = ring + static_cast<uint32_t>(u) * SLOTB;
#pragma unroll
        for (int j = 0; j < ROWB / 512; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 512 + lane * 16), "l"(src + j * 512 + lane * 16) : "memory");
        if constexpr (ROWB % 512 != 0) {
          if (lane * 16 < ROWB % 512)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (ROWB / 512) * 512 + lane * 16),
                         "l"(src + (ROWB / 512) * 512 + lane * 16) : "memory");
        }
        if constexpr (STATS) {
          if (lane == 0)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ROWB), "l"(row_stats + static_cast<int64_t>(row) * 4) : "memory");
        }
      } else {
        slot_deg[u] = -1;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");     // one group per call, empty or not: wait counts stay uniform
    };
#pragma unroll
    for (int u = 0; u < R; ++u) fill(u);
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
    while (slot_deg[0] >= 0) {                      // slots are filled in order: slot 0 empty = stream exhausted
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (slot_deg[u] >= 0) {                     // warp-uniform
          asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");    // the oldest group = this slot's
          __syncwarp();
          uint32_t raw[RAW];
          const uint32_t src = ring + static_cast<uint32_t>(u) * SLOTB;
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            const uint32_t a = src + static_cast<uint32_t>((j * 32 + lane) * W) * 2u;
            if constexpr (W == 8) {
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw[4 * j]), "=r"(raw[4 * j + 1]), "=r"(raw[4 * j + 2]),
                           "=r"(raw[4 * j + 3]) : "r"(a));
            } else {
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(raw[2 * j]), "=r"(raw[2 * j + 1]) : "r"(a));
            }
          }
          if constexpr (STATS) {
            float4 q;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(src + ROWB));
            raw[RAWD] = __float_as_uint(q.x + q.z);
            raw[RAWD + 1] = __float_as_uint(q.y + q.w);
          }
          consume_row(raw, acc);
          if (slot_deg[u] > 0) {
            finish(slot_seg[u], slot_deg[u], acc);
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
          }
          __syncwarp();                             // every lane has read the slot before it is refilled
          fill(u);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
  // decides what to do with receiver kk given its bounds; light ones get their first rows requested
  auto open_segment = [&](int64_t kk, int b, int e, bool& is_light, uint32_t (&raw)[SEG_UNROLL][RAW]) {
    is_light = false;
    if (e < b) return;
    is_light = (e - b <= HEAVY) || my_heavy >= SEG_HEAVY_PER_WARP;
    if (is_light) {
      load_batch(b, e, raw);
    } else {
      if (lane == 0) heavy_list[warp][my_heavy] = static_cast<int>(kk);
      ++my_heavy;
    }
  };
  int64_t k = warp;
  int beg, end, nbeg, nend;
  bool light = false;
  uint32_t cur[SEG_UNROLL][RAW];
  load_bounds(k, beg, end);
  load_bounds(k + SEG_WARPS, nbeg, nend);
  open_segment(k, beg, end, light, cur);
  while (end >= beg) {
    int n2beg, n2end;
    load_bounds(k + 2 * SEG_WARPS, n2beg, n2end);   // consumed in the next iteration
    bool nlight = false;
    uint32_t nxt[SEG_UNROLL][RAW];
    open_segment(k + SEG_WARPS, nbeg, nend, nlight, nxt);
    if (light) {
      float acc[NV];
      accumulate_rest(beg, end, cur, acc);
      finish(seg_of(k), end - beg, acc);
    }
    k += SEG_WARPS;
    beg = nbeg; end = nend; light = nlight;
    nbeg = n2beg; nend = n2end;
#pragma unroll
    for (int u = 0; u < SEG_UNROLL; ++u)
#pragma unroll
      for (int i = 0; i < RAW; ++i) cur[u][i] = nxt[u][i];
  }
  }
  if (lane == 0) heavy_count[warp] = my_heavy;
  __syncthreads();

  // ---- pass 2: deferred receivers (warp lists in warp order), the whole block each
  for (int w = 0; w < SEG_WARPS; ++w) {
    const int cnt = heavy_count[w];
    for (int h = 0; h < cnt; ++h) {
      const int64_t hseg = seg_of(heavy_list[w][h]);
      const int hb = __ldg(row_ptr + hseg), he = __ldg(row_ptr + hseg + 1);
      const int per = (he - hb + SEG_WARPS - 1) / SEG_WARPS;
      const int b = min(hb + warp * per, he), e = min(b + per, he);
      float acc[NV];
      if constexpr (RING) {
        // this warp's slice of the receiver's rows through its shared-memory ring (rows in flight: SEG_RING instead of one
        // batch of three; a 594-edge receiver was ~27 us of tail for its block)
        extern __shared__ __align__(16) uint8_t ring_raw[];
        constexpr int ROWB = cols * 2, SLOTB = ROWB + 16, R = SEG_RING;
        const uint32_t ring = static_cast<uint32_t>(__cvta_generic_to_shared(ring_raw)) + static_cast<uint32_t>(warp) * (R * SLOTB);
        auto request = [&](int row, int u) {
          if (row < e) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(y) + static_cast<int64_t>(row) * ldy * 2;
            const uint32_t dst = ring + static_cast<uint32_t>(u) * SLOTB;
#pragma unroll
            for (int j = 0; j < ROWB / 512; ++j)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 512 + lane * 16), "l"(src + j * 512 + lane * 16) : "memory");
            if constexpr (ROWB % 512 != 0) {
              if (lane * 16 < ROWB % 512)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (ROWB / 512) * 512 + lane * 16),
                             "l"(src + (ROWB / 512) * 512 + lane * 16) : "memory");
            }
            if constexpr (STATS) {
              if (lane == 0)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ROWB), "l"(row_stats + static_cast<int64_t>(row) * 4) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
#pragma unroll
        for (int u = 0; u < R; ++u) request(b + u, u);
        for (int r0 = b; r0 < e; r0 += R) {
#pragma unroll
          for (int u = 0; u < R; ++u) {
            if (r0 + u < e) {                       // warp-uniform
              asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");
              __syncwarp();
              uint32_t raw[RAW];
              const uint32_t src = ring + static_cast<uint32_t>(u) * SLOTB;
#pragma unroll
              for (int j = 0; j < NCH; ++j) {
                const uint32_t a = src + static_cast<uint32_t>((j * 32 + lane) * W) * 2u;
                if constexpr (W == 8) {
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw[4 * j]), "=r"(raw[4 * j + 1]), "=r"(raw[4 * j + 2]),
                               "=r"(raw[4 * j + 3]) : "r"(a));
                } else {
                  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(raw[2 * j]), "=r"(raw[2 * j + 1]) : "r"(a));
                }
              }
              if constexpr (STATS) {
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(src + ROWB));
                raw[RAWD] = __float_as_uint(q.x + q.z);
                raw[RAWD + 1] = __float_as_uint(q.y + q.w);
              }
              consume_row(raw, acc);
              __syncwarp();
              request(r0 + u + R, u);
            }
          }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        uint32_t first[SEG_UNROLL][RAW];
        load_batch(b, e, first);
        accumulate_rest(b, e, first, acc);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) partial[warp][i * 32 + lane] = acc[i];
      __syncthreads();
      if (warp == 0) {
        float tot[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) tot[i] = 0.0f;
        for (int ww = 0; ww < SEG_WARPS; ++ww)
#pragma unroll
          for (int i = 0; i < NV; ++i) tot[i] += partial[ww][i * 32 + lane];
        finish(hseg, he - hb, tot);
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float gelu_tanh_exact(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}

// grid = (layers, num_sigma); every block recomputes the 16-wide noise encoding
// (64 -> 32 -> 16, a few kFLOP) and then writes its layer's (1 + s | o) row.
__global__ void __launch_bounds__(128) cond_tables_kernel(const float* __restrict__ sigma, const float* __restrict__ w0,
                                                          const float* __restrict__ b0, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, float base_period, int nfreq,
                                                          const float* __restrict__ wc, const float* __restrict__ bc,
                                                          int layers, int width, float* __restrict__ table) {
  __shared__ float feat[128];
  __shared__ float hid[32];
  __shared__ float cond[16];
  pdl_launch_dependents();
  pdl_wait();
  const int layer = blockIdx.x;
  const int si = blockIdx.y;
  const int tid = threadIdx.x;
  const float z = logf(__ldg(sigma + si));
  if (tid < nfreq) {
    // common/model_utils.py:751-757: angular frequency 2 pi k / base_period, k = 1..K
    const float ang = z * (6.283185307179586f * static_cast<float>(tid + 1) / base_period);
    feat[tid] = cosf(ang);
    feat[nfreq + tid] = sinf(ang);
  }
  __syncthreads();
  if (tid < 32) {
    float a = __ldg(b0 + tid);
    for (int i = 0; i < 2 * nfreq; ++i) a = fmaf(feat[i], __ldg(w0 + i * 32 + tid), a);
    hid[tid] = gelu_tanh_exact(a);
  }
  __syncthreads();
  if (tid < 16) {
    float a = __ldg(b1 + tid);
    for (int i = 0; i < 32; ++i) a = fmaf(hid[i], __ldg(w1 + i * 16 + tid), a);
    cond[tid] = a;
  }
  __syncthreads();
  const float* wl = wc + static_cast<int64_t>(layer) * 16 * 2 * width;
  const float* bl = bc + static_cast<int64_t>(layer) * 2 * width;
  float* dst = table + (static_cast<int64_t>(si) * layers + layer) * 2 * width;
  for (int c = tid; c < 2 * width; c += blockDim.x) {
    float a = __ldg(bl + c);
#pragma unroll
    for (int i = 0; i < 16; ++i) a = fmaf(cond[i], __ldg(wl + i * 2 * width + c), a);
    dst[c] = c < width ? a + 1.0f : a;
  }
}

// One block per output row n.
__global__ void __launch_bounds__(128) fold_affine_kernel(const void* __restrict__ w, int dtype, int64_t ldw,
                                                          const float* __restrict__ bias,
                                                          const float* __restrict__ scale_offset, void* __restrict__ w_out,
                                                          int64_t ldw_out, float* __restrict__ bias_out, int k) {
  __shared__ float red[4];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.x;
  float dot = 0.0f;
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    float wv;
    if (dtype == GC_BF16) wv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(w)[n * ldw + c]);
    else wv = reinterpret_cast<const float*>(w)[n * ldw + c];
    const float o = wv * __ldg(scale_offset + c);
    dot = fmaf(wv, __ldg(scale_offset + k + c), dot);
    if (dtype == GC_BF16) reinterpret_cast<__nv_bfloat16*>(w_out)[n * ldw_out + c] = __float2bfloat16_rn(o);
    else reinterpret_cast<float*>(w_out)[n * ldw_out + c] = o;
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) bias_out[n] = (bias != nullptr ? bias[n] : 0.0f) + red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(256) dpm_update_kernel(const float* __restrict__ f, int64_t ldf,
                                                         const float* __restrict__ x_cur, const float* __restrict__ x_base,
                                                         int64_t ldx, const float* __restrict__ sched,
                                                         float* __restrict__ x_out, void* __restrict__ xin_out,
                                                         int xin_dtype, int64_t ld_xin, int64_t rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const float c_out = __ldg(sched), c_skip = __ldg(sched + 1), a = __ldg(sched + 2), c_in_next = __ldg(sched + 3);
  const int64_t total = rows * cols;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // four independent elements per thread and iteration: 12 loads in flight per thread (the kernel is latency-bound)
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * nthreads) {
    float fv[4], xc[4], xb[4];
    int64_t off_x[4], off_in[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * nthreads;
      ok[u] = i < total;
      const int64_t r = ok[u] ? i / cols : 0;
      const int c = ok[u] ? static_cast<int>(i - r * cols) : 0;
      off_x[u] = r * ldx + c;
      off_in[u] = r * ld_xin + c;
      fv[u] = __ldg(f + r * ldf + c);
      xc[u] = __ldg(x_cur + off_x[u]);
      xb[u] = __ldg(x_base + off_x[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const float d = c_out * fv[u] + c_skip * xc[u];
      const float xn = a * xb[u] + (1.0f - a) * d;
      x_out[off_x[u]] = xn;
      if (xin_out != nullptr) {
        const float xi = c_in_next * xn;
        if (xin_dtype == GC_BF16) reinterpret_cast<__nv_bfloat16*>(xin_out)[off_in[u]] = __float2bfloat16_rn(xi);
        else reinterpret_cast<float*>(xin_out)[off_in[u]] = xi;
      }
    }
  }
}

__global__ void __launch_bounds__(256) cast_pad_kernel(const void* __restrict__ src, int src_dtype, int64_t ld_src,
                                                       int cols_src, void* __restrict__ dst, int dst_dtype,
                                                       int64_t ld_dst, int cols_dst, const float* __restrict__ scale_dev,
                                                       int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  const float scale = scale_dev != nullptr ? __ldg(scale_dev) : 1.0f;
  const int64_t total = rows * cols_dst;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols_dst;
    const int c = static_cast<int>(i - r * cols_dst);
    float v = 0.0f;
    if (c < cols_src) {
      if (src_dtype == GC_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[r * ld_src + c]);
      else v = reinterpret_cast<const float*>(src)[r * ld_src + c];
      v *= scale;
    }
    if (dst_dtype == GC_BF16) reinterpret_cast<__nv_bfloat16*>(dst)[r * ld_dst + c] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(dst)[r * ld_dst + c] = v;
  }
}

__global__ void __launch_bounds__(256) select_columns_kernel(const float* __restrict__ s0, int64_t ld0,
                                                             const float* __restrict__ s1, int64_t ld1,
                                                             const float* __restrict__ s2, int64_t ld2,
                                                             const int32_t* __restrict__ table, float* __restrict__ out,
                                                             int64_t ldo, int64_t rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int j = static_cast<int>(i - r * cols);
    const int code = __ldg(table + j);
    const int k = code >> 24, c = code & 0xffffff;
    const float* src = k == 0 ? s0 : (k == 1 ? s1 : s2);
    const int64_t ld = k == 0 ? ld0 : (k == 1 ? ld1 : ld2);
    out[r * ldo + j] = __ldg(src + r * ld + c);
  }
}

// One warp per base row e0 < period: the base row is read once and combined, member by member, with the
// gathered rows of that member's edge e0 + m * period.  16-byte accesses, fp32 sums, MUFU activations (the same
// swish_fast / gelu_tanh_fast as the GEMM epilogues this replaces).
template <int NV>
__global__ void __launch_bounds__(256) edge_hidden_kernel(const __nv_bfloat16* __restrict__ base, int64_t ldb, int64_t period,
                                                          const __nv_bfloat16* __restrict__ g0, const int32_t* __restrict__ idx0, int64_t ld0,
                                                          const __nv_bfloat16* __restrict__ g1, const int32_t* __restrict__ idx1, int64_t ld1,
                                                          int act, __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t rows) {
  constexpr int W = NV % 8 == 0 ? 8 : 4;
  constexpr int NCH = NV / W;
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int members = static_cast<int>(rows / period);
  pdl_wait();
  auto load = [&](const __nv_bfloat16* p, float (&v)[NV]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float t[W];
      load_as_float<W>(p, GC_BF16, (j * 32 + lane) * W, t);
#pragma unroll
      for (int i = 0; i < W; ++i) v[j * W + i] = t[i];
    }
  };
  for (int64_t e0 = warp0; e0 < period; e0 += nwarps) {
    float b[NV];
    load(base + e0 * ldb, b);
    for (int m = 0; m < members; ++m) {
      const int64_t e = e0 + static_cast<int64_t>(m) * period;
      float a[NV], v[NV];
      load(g0 + static_cast<int64_t>(__ldg(idx0 + e)) * ld0, a);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = b[i] + a[i];
      if (g1 != nullptr) {
        load(g1 + static_cast<int64_t>(__ldg(idx1 + e)) * ld1, a);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += a[i];
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = apply_act<true>(v[i], act);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float t[W];
#pragma unroll
        for (int i = 0; i < W; ++i) t[i] = v[j * W + i];
        store_from_float<W>(out, GC_BF16, e * ldo + (j * 32 + lane) * W, t);
      }
    }
  }
}

// One thread per (grid point, channel): its M member values sit in a shared-memory column
// ([M][256] floats, conflict free), are insertion-sorted there, and the two CRPS terms are summed
// in member / rank order (deterministic).
constexpr int CRPS_MAX_MEMBERS = 64;
__global__ void __launch_bounds__(256) fair_crps_kernel(const float* __restrict__ members, int64_t ldm, int M,
                                                        const float* __restrict__ truth,
                                                        const float* __restrict__ weights, int channels,
                                                        float* __restrict__ out, int64_t n) {
  extern __shared__ float col[];                 // [M][256]
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + tid; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float y = __ldg(truth + i);
    float skill = 0.0f;
    for (int m = 0; m < M; ++m) {
      const float x = __ldg(members + m * ldm + i);
      skill += fabsf(x - y);
      int k = m;                                  // insertion into the sorted prefix
      while (k > 0 && col[(k - 1) * 256 + tid] > x) { col[k * 256 + tid] = col[(k - 1) * 256 + tid]; --k; }
      col[k * 256 + tid] = x;
    }
    float pair = 0.0f;
    for (int k = 0; k < M; ++k) pair = fmaf(static_cast<float>(2 * k - M + 1), col[k * 256 + tid], pair);
    const float crps = skill / static_cast<float>(M) - pair / (static_cast<float>(M) * static_cast<float>(M - 1));
    out[i] = (weights != nullptr ? __ldg(weights + i / channels) : 1.0f) * crps;
  }
}

// One block per column: strided partial sums per thread, then a fixed tree.
__global__ void __launch_bounds__(256) column_sums_kernel(const float* __restrict__ x, int64_t rows, int cols,
                                                          float* __restrict__ sums) {
  __shared__ float red[256];
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;
  float acc = 0.0f;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) acc += __ldg(x + r * cols + c);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[c] = red[0];
}

__global__ void __launch_bounds__(256) ensemble_accumulate_kernel(const float* __restrict__ x, float* __restrict__ sum,
                                                                  float* __restrict__ sumsq, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = __ldg(x + i);
    sum[i] += v;
    sumsq[i] = fmaf(v, v, sumsq[i]);
  }
}

// dst[r, c] = cast( fill_post( (fill_pre(src[r, c]) - loc[c]) / scale[c] ) ): per-channel input normalisation with NaN
// cleaning on either side of it, written straight into a (padded) GEMM operand.
__global__ void __launch_bounds__(256) normalize_cast_kernel(const float* __restrict__ src, int64_t ld_src, int cols,
                                                             const float* __restrict__ loc, const float* __restrict__ scale,
                                                             const float* __restrict__ fill_pre, const float* __restrict__ fill_post,
                                                             void* __restrict__ dst, int dst_dtype, int64_t ld_dst, int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    float v = __ldg(src + r * ld_src + c);
    if (fill_pre != nullptr && isnan(v)) v = __ldg(fill_pre + c);          // a NaN fill value keeps the NaN
    if (loc != nullptr) v = __fsub_rn(v, __ldg(loc + c));
    if (scale != nullptr) v = __fdiv_rn(v, __ldg(scale + c));
    if (fill_post != nullptr && isnan(v)) v = __ldg(fill_post + c);
    if (dst_dtype == GC_BF16) reinterpret_cast<__nv_bfloat16*>(dst)[r * ld_dst + c] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(dst)[r * ld_dst + c] = v;
  }
}

// out[r, c] = pred[r, c] * scale[c] (+ loc[c]) (+ window[r, res_col[c]], NaN-filled with res_fill[c]); NaN where one of
// the listed input columns is NaN.  No FMA contraction: bitwise what the reference's three array operations give in fp32.
__global__ void __launch_bounds__(256) unnormalize_residual_kernel(const float* __restrict__ pred, int64_t ld_pred, int cols,
                                                                   const float* __restrict__ scale, const float* __restrict__ loc,
                                                                   const float* __restrict__ window, int64_t ld_window,
                                                                   const int32_t* __restrict__ res_col,
                                                                   const float* __restrict__ res_fill,
                                                                   const int32_t* __restrict__ nan_cols, int nan_per_col,
                                                                   float* __restrict__ out, int64_t ldo, int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    float v = __ldg(pred + r * ld_pred + c);
    if (scale != nullptr) v = __fmul_rn(v, __ldg(scale + c));
    if (loc != nullptr) v = __fadd_rn(v, __ldg(loc + c));
    if (res_col != nullptr) {
      const int rc = __ldg(res_col + c);
      if (rc >= 0) {
        float last = __ldg(window + r * ld_window + rc);
        if (res_fill != nullptr && isnan(last)) last = __ldg(res_fill + c);
        v = __fadd_rn(v, last);
      }
    }
    if (nan_cols != nullptr) {
      for (int k = 0; k < nan_per_col; ++k) {
        const int nc = __ldg(nan_cols + c * nan_per_col + k);
        if (nc >= 0 && isnan(__ldg(window + r * ld_window + nc))) v = __int_as_float(0x7fc00000);
      }
    }
    out[r * ldo + c] = v;
  }
}

int sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

unsigned grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
  const int64_t need = (work_items + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(sm_count()) * blocks_per_sm;
  return static_cast<unsigned>(need < 1 ? 1 : (need < cap ? need : cap));
}

bool dtype_ok(int d) { return d == GC_F32 || d == GC_BF16; }

}  // namespace
}  // namespace gc

extern "C" {

using namespace gc;

int gc_ln_cond(void* stream, const void* x, int32_t x_dtype, int64_t ldx, const float* scale_offset,
               int32_t do_layer_norm, const void* residual, int32_t res_dtype, int64_t ld_res, void* out,
               int32_t out_dtype, int64_t ldo, int64_t rows, int32_t cols) {
  GC_REQUIRE(x && out, "gc_ln_cond: null buffer");
  GC_REQUIRE(cols == 128 || cols == 256 || cols == 512, "gc_ln_cond: cols=%d (supported: 128, 256, 512)", cols);
  GC_REQUIRE(dtype_ok(x_dtype) && dtype_ok(out_dtype), "gc_ln_cond: bad dtype");
  GC_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0 && aligned16(x) && aligned16(out), "gc_ln_cond: alignment");
  if (residual) GC_REQUIRE(dtype_ok(res_dtype) && ld_res % 4 == 0 && aligned16(residual), "gc_ln_cond: residual alignment");
  if (rows <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // grid-stride kernel: one wave of exactly the blocks that are resident at once (a partial second wave is a tail)
  auto resident = [](auto kernel) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, 0) != cudaSuccess || n < 1) n = 4;
    return n;
  };
#define GC_LAUNCH_LN(NV)                                                                                        \
  do {                                                                                                          \
  if (residual != nullptr) {                                                                                    \
    static const int occ = resident(ln_cond_kernel<NV, true>);                                                  \
    GC_CHECK_CUDA(launch_kernel(ln_cond_kernel<NV, true>, dim3(grid_for(rows, 8, occ)), dim3(256), 0, st, x, x_dtype, ldx, \
                                scale_offset, do_layer_norm, residual, res_dtype, ld_res, out, out_dtype, ldo, rows), \
                  "ln_cond_kernel");                                                                            \
  } else {                                                                                                      \
    static const int occ = resident(ln_cond_kernel<NV, false>);                                                 \
    GC_CHECK_CUDA(launch_kernel(ln_cond_kernel<NV, false>, dim3(grid_for(rows, 8, occ)), dim3(256), 0, st, x, x_dtype, ldx, \
                                scale_offset, do_layer_norm, residual, res_dtype, ld_res, out, out_dtype, ldo, rows), \
                  "ln_cond_kernel");                                                                            \
  }                                                                                                             \
  } while (0)
  if (cols == 128) GC_LAUNCH_LN(4);
  else if (cols == 256) GC_LAUNCH_LN(8);
  else GC_LAUNCH_LN(16);
#undef GC_LAUNCH_LN
  GC_CHECK_LAUNCH("ln_cond_kernel");
  return GC_OK;
}

int gc_ln_cond_segment_sum_stats(void* stream, const void* y, int32_t y_dtype, int64_t ldy, const float* scale_offset,
                                 int32_t do_layer_norm, const int32_t* row_ptr, const int32_t* edge_perm, void* out,
                                 int32_t out_dtype, int64_t ldo, int64_t num_segments, int32_t cols, const float* row_stats) {
  GC_REQUIRE(y && out && row_ptr, "gc_ln_cond_segment_sum: null buffer");
  GC_REQUIRE(cols == 128 || cols == 256 || cols == 512, "gc_ln_cond_segment_sum: cols=%d (supported: 128, 256, 512)", cols);
  GC_REQUIRE(dtype_ok(y_dtype) && dtype_ok(out_dtype), "gc_ln_cond_segment_sum: bad dtype");
  GC_REQUIRE(ldy % 8 == 0 && ldo % 8 == 0 && aligned16(y) && aligned16(out), "gc_ln_cond_segment_sum: alignment");
  if (scale_offset) GC_REQUIRE(aligned16(scale_offset), "gc_ln_cond_segment_sum: scale_offset alignment");
  if (row_stats) GC_REQUIRE(aligned16(row_stats) && y_dtype == GC_BF16 && (do_layer_norm & 1),
                            "gc_ln_cond_segment_sum_stats: row statistics need 16-byte alignment, bf16 rows and LayerNorm");
  if (num_segments <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int do_ln = do_layer_norm & 1;
  const bool irregular = edge_perm != nullptr || (do_layer_norm & GC_SEGSUM_IRREGULAR) != 0;
#define GC_LAUNCH_SEG(NV, BF, OCC, ST, RG)                                                                      \
  do {                                                                                                          \
    auto kernel = ln_cond_segment_sum_kernel<NV, BF, OCC, ST, RG>;                                               \
    const size_t ring_bytes = RG ? static_cast<size_t>(SEG_WARPS) * SEG_RING * (NV * 32 * 2 + 16) : 0;          \
    if (RG) GC_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes), \
                          "cudaFuncSetAttribute(ln_cond_segment_sum_kernel)");                                  \
    GC_CHECK_CUDA(launch_kernel(kernel, dim3(grid_for(num_segments, SEG_WARPS, OCC)), dim3(SEG_WARPS * 32), ring_bytes, st, \
                                y, ldy, scale_offset, do_ln, row_ptr, edge_perm, out, out_dtype, ldo, num_segments,    \
                                row_stats), "ln_cond_segment_sum_kernel");                                      \
  } while (0)
  // bf16 rows without an edge permutation stream through the shared-memory ring (GENCAST_SEGSUM_RING=0: register batches)
  static const bool ring_on = []() { const char* v = getenv("GENCAST_SEGSUM_RING"); return !(v != nullptr && v[0] == '0'); }();
  const bool ring = ring_on && y_dtype == GC_BF16 && edge_perm == nullptr;
  if (ring && row_stats != nullptr) {
    if (cols == 128) GC_LAUNCH_SEG(4, true, 3, true, true);
    else if (cols == 256) GC_LAUNCH_SEG(8, true, 3, true, true);
    else GC_LAUNCH_SEG(16, true, 3, true, true);
  } else if (ring) {
    if (cols == 128) GC_LAUNCH_SEG(4, true, 3, false, true);
    else if (cols == 256) GC_LAUNCH_SEG(8, true, 3, false, true);
    else GC_LAUNCH_SEG(16, true, 3, false, true);
  } else if (y_dtype == GC_BF16 && row_stats != nullptr) {
    if (cols == 128) GC_LAUNCH_SEG(4, true, 2, true, false);
    else if (cols == 256) GC_LAUNCH_SEG(8, true, 2, true, false);
    else GC_LAUNCH_SEG(16, true, 3, true, false);
  } else if (y_dtype == GC_BF16) {
    if (cols == 128) GC_LAUNCH_SEG(4, true, 2, false, false);
    else if (cols == 256) GC_LAUNCH_SEG(8, true, 2, false, false);
    else if (irregular) GC_LAUNCH_SEG(16, true, 3, false, false);
    else GC_LAUNCH_SEG(16, true, 2, false, false);
  } else {
    if (cols == 128) GC_LAUNCH_SEG(4, false, 2, false, false);
    else if (cols == 256) GC_LAUNCH_SEG(8, false, 2, false, false);
    else GC_LAUNCH_SEG(16, false, 2, false, false);
  }
#undef GC_LAUNCH_SEG
  GC_CHECK_LAUNCH("ln_cond_segment_sum_kernel");
  return GC_OK;
}

int gc_ln_cond_segment_sum(void* stream, const void* y, int32_t y_dtype, int64_t ldy, const float* scale_offset,
                           int32_t do_layer_norm, const int32_t* row_ptr, const int32_t* edge_perm, void* out,
                           int32_t out_dtype, int64_t ldo, int64_t num_segments, int32_t cols) {
  return gc_ln_cond_segment_sum_stats(stream, y, y_dtype, ldy, scale_offset, do_layer_norm, row_ptr, edge_perm, out, out_dtype,
                                      ldo, num_segments, cols, nullptr);
}

int gc_cond_tables(void* stream, const float* sigma, int32_t num_sigma, const float* w0, const float* b0,
                   const float* w1, const float* b1, float base_period, int32_t num_frequencies, const float* wc,
                   const float* bc, int32_t layers, int32_t width, float* table) {
  GC_REQUIRE(sigma && w0 && b0 && w1 && b1 && wc && bc && table, "gc_cond_tables: null buffer");
  GC_REQUIRE(num_frequencies >= 1 && num_frequencies <= 64, "gc_cond_tables: num_frequencies=%d", num_frequencies);
  GC_REQUIRE(layers >= 1 && width >= 1 && num_sigma >= 1 && num_sigma <= 65535, "gc_cond_tables: bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(cond_tables_kernel, dim3(layers, num_sigma), dim3(128), 0, st, sigma, w0, b0, w1, b1,
                              base_period, num_frequencies, wc, bc, layers, width, table), "cond_tables_kernel");
  GC_CHECK_LAUNCH("cond_tables_kernel");
  return GC_OK;
}

int gc_fold_affine_into_linear(void* stream, const void* w, int32_t dtype, int64_t ldw, const float* bias,
                               const float* scale_offset, void* w_out, int64_t ldw_out, float* bias_out, int32_t n,
                               int32_t k) {
  GC_REQUIRE(w && scale_offset && w_out && bias_out, "gc_fold_affine_into_linear: null buffer");
  GC_REQUIRE(dtype_ok(dtype) && n > 0 && k > 0, "gc_fold_affine_into_linear: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(fold_affine_kernel, dim3(n), dim3(128), 0, st, w, dtype, ldw, bias, scale_offset, w_out,
                              ldw_out, bias_out, k), "fold_affine_kernel");
  GC_CHECK_LAUNCH("fold_affine_kernel");
  return GC_OK;
}

int gc_dpm_update(void* stream, const float* f, int64_t ldf, const float* x_cur, const float* x_base, int64_t ldx,
                  const float* sched, float* x_out, void* xin_out, int32_t xin_dtype, int64_t ld_xin, int64_t rows,
                  int32_t cols) {
  GC_REQUIRE(f && x_cur && x_base && sched && x_out, "gc_dpm_update: null buffer");
  GC_REQUIRE(cols > 0 && ldf >= cols && ldx >= cols, "gc_dpm_update: bad sizes");
  if (xin_out) GC_REQUIRE(dtype_ok(xin_dtype) && ld_xin >= cols, "gc_dpm_update: bad xin");
  if (rows <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(dpm_update_kernel, dim3(grid_for(rows * cols, 256 * 4, 8)), dim3(256), 0, st, f, ldf, x_cur,
                              x_base, ldx, sched, x_out, xin_out, xin_dtype, ld_xin, rows, cols), "dpm_update_kernel");
  GC_CHECK_LAUNCH("dpm_update_kernel");
  return GC_OK;
}

int gc_cast_pad(void* stream, const void* src, int32_t src_dtype, int64_t ld_src, int32_t cols_src, void* dst,
                int32_t dst_dtype, int64_t ld_dst, int32_t cols_dst, const float* scale_dev, int64_t rows) {
  GC_REQUIRE(src && dst, "gc_cast_pad: null buffer");
  GC_REQUIRE(dtype_ok(src_dtype) && dtype_ok(dst_dtype), "gc_cast_pad: bad dtype");
  GC_REQUIRE(cols_src >= 0 && cols_dst > 0 && ld_src >= cols_src && ld_dst >= cols_dst, "gc_cast_pad: bad sizes");
  if (rows <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(cast_pad_kernel, dim3(grid_for(rows * cols_dst, 256 * 4, 8)), dim3(256), 0, st, src, src_dtype,
                              ld_src, cols_src, dst, dst_dtype, ld_dst, cols_dst, scale_dev, rows), "cast_pad_kernel");
  GC_CHECK_LAUNCH("cast_pad_kernel");
  return GC_OK;
}

int gc_normalize_cast(void* stream, const float* src, int64_t ld_src, int32_t cols, const float* loc, const float* scale,
                      const float* fill_pre, const float* fill_post, void* dst, int32_t dst_dtype, int64_t ld_dst, int64_t rows) {
  GC_REQUIRE(src && dst, "gc_normalize_cast: null buffer");
  GC_REQUIRE(dtype_ok(dst_dtype) && cols > 0 && ld_src >= cols && ld_dst >= cols, "gc_normalize_cast: bad arguments");
  if (rows <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(normalize_cast_kernel, dim3(grid_for(rows * cols, 256 * 4, 8)), dim3(256), 0, st, src, ld_src, (int)cols,
                              loc, scale, fill_pre, fill_post, dst, (int)dst_dtype, ld_dst, rows), "normalize_cast_kernel");
  return GC_OK;
}

int gc_unnormalize_residual(void* stream, const float* pred, int64_t ld_pred, int32_t cols, const float* scale, const float* loc,
                            const float* window, int64_t ld_window, const int32_t* res_col, const float* res_fill,
                            const int32_t* nan_cols, int32_t nan_per_col, float* out, int64_t ldo, int64_t rows) {
  GC_REQUIRE(pred && out, "gc_unnormalize_residual: null buffer");
  GC_REQUIRE(cols > 0 && ld_pred >= cols && ldo >= cols, "gc_unnormalize_residual: bad sizes");
  GC_REQUIRE((res_col == nullptr && nan_cols == nullptr) || window != nullptr, "gc_unnormalize_residual: window missing");
  GC_REQUIRE(nan_cols == nullptr || nan_per_col > 0, "gc_unnormalize_residual: nan_per_col");
  if (rows <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(unnormalize_residual_kernel, dim3(grid_for(rows * cols, 256 * 4, 8)), dim3(256), 0, st, pred, ld_pred,
                              (int)cols, scale, loc, window, ld_window, res_col, res_fill, nan_cols, (int)nan_per_col, out, ldo, rows),
                "unnormalize_residual_kernel");
  return GC_OK;
}

int gc_select_columns(void* stream, const float* src0, int64_t ld0, const float* src1, int64_t ld1, const float* src2,
                      int64_t ld2, const int32_t* table, float* out, int64_t ldo, int64_t rows, int32_t cols_out) {
  GC_REQUIRE(src0 && table && out, "gc_select_columns: null buffer");
  GC_REQUIRE(cols_out > 0 && ldo >= cols_out, "gc_select_columns: bad sizes");
  GC_REQUIRE(out != src0 && out != src1 && out != src2, "gc_select_columns: out aliases a source");
  if (rows <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(select_columns_kernel, dim3(grid_for(rows * cols_out, 256 * 4, 8)), dim3(256), 0, st, src0, ld0,
                              src1, ld1, src2, ld2, table, out, ldo, rows, cols_out), "select_columns_kernel");
  GC_CHECK_LAUNCH("select_columns_kernel");
  return GC_OK;
}

int gc_edge_hidden(void* stream, const void* base, int64_t ld_base, int64_t period, const void* g0, const int32_t* idx0,
                   int64_t ld0, const void* g1, const int32_t* idx1, int64_t ld1, int32_t act, void* out, int64_t ldo,
                   int64_t rows, int32_t cols) {
  GC_REQUIRE(base && g0 && idx0 && out, "gc_edge_hidden: null buffer");
  GC_REQUIRE((g1 == nullptr) == (idx1 == nullptr), "gc_edge_hidden: g1 and idx1 go together");
  GC_REQUIRE(cols == 128 || cols == 256 || cols == 512, "gc_edge_hidden: cols=%d (supported: 128, 256, 512)", cols);
  GC_REQUIRE(period > 0 && rows >= 0 && rows % period == 0, "gc_edge_hidden: rows must be a multiple of period");
  GC_REQUIRE(ld_base % 8 == 0 && ld0 % 8 == 0 && ldo % 8 == 0 && (g1 == nullptr || ld1 % 8 == 0) && aligned16(base) &&
                 aligned16(g0) && aligned16(out) && (g1 == nullptr || aligned16(g1)),
             "gc_edge_hidden: alignment");
  GC_REQUIRE(act == GC_ACT_NONE || act == GC_ACT_SWISH || act == GC_ACT_GELU_TANH, "gc_edge_hidden: act=%d", act);
  if (rows == 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = grid_for(period, 8, 8);
#define GC_LAUNCH_EH(NV)                                                                                            \
  GC_CHECK_CUDA(launch_kernel(edge_hidden_kernel<NV>, dim3(grid), dim3(256), 0, st,                                  \
                              reinterpret_cast<const __nv_bfloat16*>(base), ld_base, period,                         \
                              reinterpret_cast<const __nv_bfloat16*>(g0), idx0, ld0,                                 \
                              reinterpret_cast<const __nv_bfloat16*>(g1), idx1, ld1, (int)act,                       \
                              reinterpret_cast<__nv_bfloat16*>(out), ldo, rows), "edge_hidden_kernel")
  if (cols == 128) GC_LAUNCH_EH(4);
  else if (cols == 256) GC_LAUNCH_EH(8);
  else GC_LAUNCH_EH(16);
#undef GC_LAUNCH_EH
  GC_CHECK_LAUNCH("edge_hidden_kernel");
  return GC_OK;
}

int gc_fair_crps(void* stream, const float* members, int64_t ld_members, int32_t num_members, const float* truth,
                 const float* weights, int32_t channels, float* out, int64_t n) {
  GC_REQUIRE(members && truth && out, "gc_fair_crps: null buffer");
  GC_REQUIRE(num_members >= 2 && num_members <= CRPS_MAX_MEMBERS, "gc_fair_crps: num_members=%d (2 .. %d)", num_members,
             CRPS_MAX_MEMBERS);
  GC_REQUIRE(channels >= 1 && ld_members >= n, "gc_fair_crps: bad sizes");
  if (n <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(num_members) * 256 * sizeof(float);
  GC_CHECK_CUDA(cudaFuncSetAttribute(fair_crps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                "cudaFuncSetAttribute(fair_crps_kernel)");
  GC_CHECK_CUDA(launch_kernel(fair_crps_kernel, dim3(grid_for(n, 256, 2)), dim3(256), smem, st, members, ld_members,
                              (int)num_members, truth, weights, (int)channels, out, n), "fair_crps_kernel");
  GC_CHECK_LAUNCH("fair_crps_kernel");
  return GC_OK;
}

int gc_column_sums(void* stream, const float* x, int64_t rows, int32_t cols, float* sums) {
  GC_REQUIRE(x && sums, "gc_column_sums: null buffer");
  GC_REQUIRE(cols >= 1 && rows >= 0, "gc_column_sums: bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(column_sums_kernel, dim3(cols), dim3(256), 0, st, x, rows, (int)cols, sums), "column_sums_kernel");
  GC_CHECK_LAUNCH("column_sums_kernel");
  return GC_OK;
}

int gc_ensemble_accumulate(void* stream, const float* x, float* sum, float* sumsq, int64_t n) {
  GC_REQUIRE(x && sum && sumsq, "gc_ensemble_accumulate: null buffer");
  if (n <= 0) return GC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GC_CHECK_CUDA(launch_kernel(ensemble_accumulate_kernel, dim3(grid_for(n, 256 * 4, 8)), dim3(256), 0, st, x, sum, sumsq, n),
                "ensemble_accumulate_kernel");
  GC_CHECK_LAUNCH("ensemble_accumulate_kernel");
  return GC_OK;
}

}  // extern "C"
