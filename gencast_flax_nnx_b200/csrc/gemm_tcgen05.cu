// bf16 tensor-core linear layer for sm_100a: TMA-staged operand tiles, tcgen05.mma
// with the fp32 accumulator in tensor memory, fused epilogue read back with tcgen05.ld.
//
// One CTA computes a 128 x 128 output tile.  Warp roles:
//   warp 0     TMA producer (one elected lane): A[128 x 64] and W[128 x 64] bf16 tiles,
//              128-byte swizzle, STAGES-deep ring guarded by full/empty mbarriers
//   warp 1     allocates 128 TMEM columns, issues tcgen05.mma (M=128, N=128, K=16) x 4
//              per stage, releases stages with tcgen05.commit
//   warps 2-5  epilogue: each warp owns the 32 TMEM lanes (= output rows) of its
//              quarter, reads 16 columns at a time and applies the fused epilogue
//              of common.cuh (bias / addend / row gathers / activation / residual).
// Two CTAs fit per SM (3 x 32 KB stages each, 128 of 512 TMEM columns each), so one
// CTA's epilogue overlaps the other's main loop.
//
// The K dimension may be split in up to three (A_s, W_s) segments: this is how the
// [e | sender | receiver] and [node | aggregate] concatenations of the reference
// (common/typed_graph_net.py:301-305, :315-326) are consumed without being built.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace gc {

long long* g_gemm_trace = nullptr;   // debug: clock stamps of CTA 0 of the persistent kernel (null in normal use)

namespace {

#define GC_GTRACE(role, index)                                                                          \
  do {                                                                                                  \
    if (trace != nullptr && blockIdx.x == 0 && (index) < 512) trace[(role) * 512 + (index)] = clock64(); \
  } while (0)

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 192;

// BN = 128 is the throughput shape; BN = 64 doubles the CTA count for the small
// mesh-side problems (a few thousand rows) that would otherwise leave most SMs idle.
template <int BN, int STAGES_>
struct GemmCfg {
  static constexpr int STAGES = STAGES_;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;   // barrier block: up to 2 * 8 + 2 words of 8 bytes
  static constexpr int SMEM_BYTES = BIAS_OFFSET + BN * 4 + 1024 /*alignment slack*/;
  static constexpr uint32_t TMEM_COLS = BN;
};

struct GemmMaps {
  CUtensorMap a[GC_MAX_SEGMENTS];
  CUtensorMap w[GC_MAX_SEGMENTS];
  CUtensorMap out;                 // only valid when shape.store_mode != STORE_DIRECT
};

// How the persistent kernels write their tiles.
//   STORE_DIRECT  each thread stores its own row segments (general fused epilogue: addend, gathers, ...)
//   STORE_TMA     bias / activation only: rows are staged in shared memory (128-byte swizzle) and
//                 written with cp.async.bulk.tensor, i.e. as full lines instead of 32 partial lines
//                 per instruction
//   STORE_TMA_ADD fp32 output that is also the residual (x = x + A W^T + b): same staging, written
//                 with cp.reduce.async.bulk.tensor .add, so the residual is never read by the SM
constexpr int STORE_DIRECT = 0;
constexpr int STORE_TMA = 1;
constexpr int STORE_TMA_ADD = 2;

struct GemmShape {
  int kblocks[GC_MAX_SEGMENTS];
  int num_segments;
  int n_tiles;
  int store_mode;
  int gather_staged;               // row gathers go through the shared-memory transposition (staged store modes only)
  int early_w;                     // W tiles of the first stages may be fetched before griddepcontrol.wait
};

// Epilogue of one warp for one accumulator buffer: 32 rows (its TMEM lane quarter) x HALF columns.
// `release` is called once, as soon as the last tcgen05.ld of the buffer has landed in registers.
// stage_smem: this warp's two 4 KB staging buffers (1024-byte aligned) for the TMA store modes.
//
// Row gathers in the staged modes (gather_smem != 0): the rows of a gather table that the 32 output
// rows of this warp need are fetched with coalesced 16-byte loads (4 lanes per 64-byte row segment,
// 8 rows per instruction), transposed through a 2 KB swizzled shared-memory buffer into the
// row-per-thread layout of the accumulator, and added in fp32.  A thread-per-row global load touches
// 32 lines per instruction and made the L1 tag stage the bottleneck of the edge GEMMs (2.8 x the time
// of the same GEMM without gathers); this way it is 8.  Loads run one 32-column chunk ahead.
// NBUF: staging buffers of 4 KB per warp (2 = a TMA store may still be reading one while the next is filled).
template <int HALF, int NBUF = 2, typename WaitAcc, typename Release>
__device__ __forceinline__ void epilogue_warp_tile(const EpilogueParams& ep, const CUtensorMap* out_map, int store_mode,
                                                   float alpha, uint32_t taddr, const float* bias_ptr, int64_t row0,
                                                   int col_base, int lane, uint32_t stage_smem, int& stage_use,
                                                   uint32_t gather_smem, WaitAcc wait_acc, Release release,
                                                   long long* etrace = nullptr, int* etrace_ev = nullptr,
                                                   const int ncols = HALF) {
  using namespace sm100;
#define GC_ESTAMP()                                                                                     \
  do {                                                                                                  \
    if (etrace != nullptr && *etrace_ev < 4 * 512) { etrace[4 * 512 + *etrace_ev] = clock64(); ++*etrace_ev; } \
  } while (0)
  const int64_t row = row0 + lane;
  const bool gathers = gather_smem != 0u && store_mode != STORE_DIRECT;
  const bool has_g1 = gathers && ep.gsrc1 != nullptr;
  int gi0 = 0, gi1 = 0;
  if (gathers && row < ep.m) {
    gi0 = __ldg(ep.gidx0 + row);
    if (has_g1) gi1 = __ldg(ep.gidx1 + row);
  }
  // lane -> (row 8 i + lane / 4, 16-byte unit lane % 4) of a [32 rows x 32 bf16] block
  const int g_sub = lane >> 2, g_unit = lane & 3;
  auto gather_issue = [&](const void* src, int64_t ld, int gi, int c, uint4 (&g)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t idx = __shfl_sync(0xffffffffu, gi, 8 * i + g_sub);
      g[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + idx * ld + col_base + c) + g_unit);
    }
  };
  auto gather_add = [&](const uint4 (&g)[4], float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t rl = static_cast<uint32_t>(8 * i + g_sub);
      const uint32_t addr = gather_smem + rl * 64u + ((static_cast<uint32_t>(g_unit) ^ ((rl >> 1) & 3u)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(g[i].x), "r"(g[i].y), "r"(g[i].z), "r"(g[i].w)
                   : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t addr = gather_smem + static_cast<uint32_t>(lane) * 64u +
                            ((static_cast<uint32_t>(u) ^ ((static_cast<uint32_t>(lane) >> 1) & 3u)) << 4);
      uint4 raw;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(addr));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        v[8 * u + 2 * j] += f.x;
        v[8 * u + 2 * j + 1] += f.y;
      }
    }
    __syncwarp();
  };
  uint4 ga[4], gb[4];
  if (gathers) {
    gather_issue(ep.gsrc0, ep.ldg0, gi0, 0, ga);
    if (has_g1) gather_issue(ep.gsrc1, ep.ldg1, gi1, 0, gb);
  }
  wait_acc();
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr, r);
  // Not unrolled: every register array is indexed by compile-time constants inside one chunk, and four
  // copies of this body (~35 KB of SASS each way through the store modes) thrash the instruction cache
  // of an SM whose ten warps sit in different copies ('no_inst' stalls in the source-level profile).
#pragma unroll 1
  for (int c = 0; c < ncols; c += 32) {
    float v[32];
    GC_ESTAMP();                 // chunk start
    tc_wait_ld();
    GC_ESTAMP();                 // accumulator chunk in registers
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (c + 32 < ncols) {
      tmem_ld_32x32b_x32(taddr + c + 32, r);
    } else {
      tc_fence_before();
      __syncwarp();
      release();
    }
    if (store_mode == STORE_DIRECT) {
      if (row < ep.m) epilogue_row_segment<32, true>(ep, alpha, row, col_base + c, v, bias_ptr ? bias_ptr + c : nullptr);
      continue;
    }
    // ---- staged path: bias / activation in registers
    if (ep.alpha_dev != nullptr) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= alpha;
    }
    if (bias_ptr != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = reinterpret_cast<const float4*>(bias_ptr + c)[i];
        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    }
    if (gathers) {
      // this chunk's rows are in ga / gb; the next chunk's loads are issued before they are consumed
      uint4 na[4] = {}, nb[4] = {};
      const bool more = c + 32 < ncols;
      if (more) gather_issue(ep.gsrc0, ep.ldg0, gi0, c + 32, na);
      gather_add(ga, v);
      if (has_g1) {
        if (more) gather_issue(ep.gsrc1, ep.ldg1, gi1, c + 32, nb);
        gather_add(gb, v);
      }
      if (more) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { ga[i] = na[i]; gb[i] = nb[i]; }
      }
    }
    if (ep.act == GC_ACT_SWISH) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = swish_fast(v[i]);
    } else if (ep.act == GC_ACT_GELU_TANH) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_fast(v[i]);
    }
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    if (ep.out_dtype == GC_BF16) {
      // 32 bf16 = 64 B = units (c/32 & 1) * 4 .. + 3 of this row in a 64-column chunk
      const int half_chunk = (c >> 5) & 1;
      if (half_chunk == 0) {
        // first half of a new chunk: the buffer's previous store must have finished reading it
        if (lane == 0) bulk_wait_group_read<NBUF - 1>();
        __syncwarp();
      }
      GC_ESTAMP();               // staging buffer free
      const uint32_t buf = stage_smem + static_cast<uint32_t>(stage_use & (NBUF - 1)) * 4096u + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * u + 2 * j], v[8 * u + 2 * j + 1]);
          pk[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        const uint32_t unit = static_cast<uint32_t>(half_chunk * 4 + u) ^ sw;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + unit * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                     "r"(pk[3]) : "memory");
      }
      if (half_chunk == 1 || c + 32 >= ncols) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(out_map, stage_smem + static_cast<uint32_t>(stage_use & (NBUF - 1)) * 4096u, col_base + (c & ~63), static_cast<int>(row0));
          bulk_commit_group();
        }
        ++stage_use;
      }
      GC_ESTAMP();               // chunk stored / TMA store issued
    } else {
      // 32 fp32 = 128 B = one full row of a 32-column chunk
      if (lane == 0) bulk_wait_group_read<NBUF - 1>();
      __syncwarp();
      const uint32_t buf = stage_smem + static_cast<uint32_t>(stage_use & (NBUF - 1)) * 4096u + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t unit = static_cast<uint32_t>(u) ^ sw;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + unit * 16), "r"(__float_as_uint(v[4 * u])),
                     "r"(__float_as_uint(v[4 * u + 1])), "r"(__float_as_uint(v[4 * u + 2])), "r"(__float_as_uint(v[4 * u + 3]))
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t src = stage_smem + static_cast<uint32_t>(stage_use & (NBUF - 1)) * 4096u;
        if (store_mode == STORE_TMA_ADD) tma_reduce_add_2d(out_map, src, col_base + c, static_cast<int>(row0));
        else tma_store_2d(out_map, src, col_base + c, static_cast<int>(row0));
        bulk_commit_group();
      }
      ++stage_use;
    }
  }
}

template <int BN, int NSTAGES>
__global__ void __launch_bounds__(NUM_THREADS, NSTAGES > 4 ? 1 : 2)
gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const GemmShape shape, const EpilogueParams ep) {
  using namespace sm100;
  pdl_launch_dependents();     // let the next kernel's launch and prologue overlap this one
  using C = GemmCfg<BN, NSTAGES>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bars = smem_base + C::BAR_OFFSET;
  float* bias_s = reinterpret_cast<float*>(smem_gen + C::BIAS_OFFSET);
  // barrier block: full[STAGES] | empty[STAGES] | tmem_full | tmem_ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_smem = bars + 8u * (2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x % shape.n_tiles;
  const int m_blk = blockIdx.x / shape.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < shape.num_segments; ++s) {
      prefetch_tensormap(&maps.a[s]);
      prefetch_tensormap(&maps.w[s]);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
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

  // Producer and MMA warps run warp-uniform code and elect one lane for the TMA / tcgen05 instructions: inside an
  // `if (lane == 0)` region the compiler treats the operands as divergent and wraps every UTCHMMA / UTMALDG in an
  // ELECT + R2UR.BROADCAST x5 + BRA.U.ANY loop (~15 dependent instructions per MMA: the 590-650 clk k-block period
  // measured in round 1 against 512 clk of tensor work); with uniform control flow the descriptors live in uniform
  // registers and the four MMAs of a k-block issue back to back.
  if (warp == 0) {
    // Weights first: with static W (shape.early_w) the W tiles of the first ring pass are requested before
    // waiting for the predecessor grid, so their HBM latency overlaps its tail.
    int pre = 0;
    if (shape.early_w) {
      for (int s = 0; s < shape.num_segments && pre < STAGES; ++s)
        for (int kb = 0; kb < shape.kblocks[s] && pre < STAGES; ++kb, ++pre) {
          if (elect_one()) {
            mbar_arrive_expect_tx(full_bar(pre), C::STAGE_BYTES);
            tma_load_2d(smem_b + pre * C::B_STAGE_BYTES, &maps.w[s], full_bar(pre), kb * BK, n_blk * BN);
          }
          __syncwarp();
        }
    }
    pdl_wait();            // first touch of operands a predecessor may have produced
    int stage = 0, issued = 0;
    uint32_t phase = 0;
    for (int s = 0; s < shape.num_segments; ++s) {
      for (int kb = 0; kb < shape.kblocks[s]; ++kb, ++issued) {
        if (issued >= pre) mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          if (issued >= pre) {
            mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
            tma_load_2d(smem_b + stage * C::B_STAGE_BYTES, &maps.w[s], full_bar(stage), kb * BK, n_blk * BN);
          }
          tma_load_2d(smem_a + stage * A_STAGE_BYTES, &maps.a[s], full_bar(stage), kb * BK, m_blk * BM);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int s = 0; s < shape.num_segments; ++s) {
      for (int kb = 0; kb < shape.kblocks[s]; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t da = desc_kmajor_sw128(smem_a + stage * A_STAGE_BYTES);
        const uint64_t db = desc_kmajor_sw128(smem_b + stage * C::B_STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle span: +2 in the (>>4) address field
            umma_f16(tmem_base, da + 2u * k, db + 2u * k, idesc, (accumulate | static_cast<uint32_t>(k)) != 0u ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // Epilogue: TMEM lane quarter is fixed by warp id modulo 4.  32 columns per step, the next
    // step's tcgen05.ld is in flight while the current one is processed.
    const int q = warp & 3;
    const int64_t row = static_cast<int64_t>(m_blk) * BM + q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float* bias_ptr = ep.bias != nullptr ? bias_s : nullptr;
    pdl_wait();
    {
      // bias tile -> shared memory while the main loop runs (read back as broadcasts)
      const int t = threadIdx.x - 64;
      if (t < BN) bias_s[t] = ep.bias != nullptr ? __ldg(ep.bias + n_blk * BN + t) : 0.0f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const float alpha = ep.alpha_dev != nullptr ? __ldg(ep.alpha_dev) : 1.0f;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr, r);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      if (c + 32 < BN) tmem_ld_32x32b_x32(taddr + c + 32, r);
      if (row < ep.m) epilogue_row_segment<32, true>(ep, alpha, row, n_blk * BN + c, v, bias_ptr ? bias_ptr + c : nullptr);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant for problems with at least ~2 tiles per SM: one CTA per SM walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  The accumulator is double buffered in TMEM, so the
// tensor core starts tile i+1 while eight epilogue warps (two per TMEM lane quarter, half the
// columns each) drain tile i; the TMA producer runs ahead through the operand ring across tile
// boundaries.  128 x 128 tiles are bound by the L2 -> SM operand feed (64 FLOP per byte staged,
// ~6.3 KB/clk chip-wide => ~1/3 of the tensor peak), so the tile is 128 x 256 whenever N allows
// it (87 FLOP/B): TMEM then holds exactly two 256-column accumulators.
// ---------------------------------------------------------------------------------------------
template <int PBN>
struct PersistCfg {
  static constexpr int STAGES = PBN == 256 ? 3 : 4;
  static constexpr int B_STAGE_BYTES = PBN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;        // 8 warps x 2 x 4 KB store staging
  static constexpr int GATHER_BYTES = PBN == 256 ? 0 : 8 * 2048;     // 8 warps x 2 KB gather transposition (no room at 256)
  static constexpr int GATHER_OFFSET = STAGING_OFFSET + 8 * 8192;
  static constexpr int BAR_OFFSET = GATHER_OFFSET + GATHER_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
  static constexpr int SMEM_BYTES = BIAS_OFFSET + 2 * PBN * 4 + 1024;
};
constexpr int P_THREADS = 320;

template <int PBN>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_bf16_tcgen05_persistent_kernel(const __grid_constant__ GemmMaps maps, const GemmShape shape, const EpilogueParams ep,
                                    const int num_tiles, long long* trace) {
  using namespace sm100;
  pdl_launch_dependents();     // let the next kernel's launch and prologue overlap this one
  using C = PersistCfg<PBN>;
  constexpr int PSTAGES = C::STAGES;
  constexpr int HALF = PBN / 2;                 // columns per epilogue warp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + PSTAGES * A_STAGE_BYTES;
  const uint32_t bars = smem_base + C::BAR_OFFSET;
  float* bias_s = reinterpret_cast<float*>(smem_gen + C::BIAS_OFFSET);      // [2][PBN]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (PSTAGES + s); };
  auto acc_full = [&](int b) { return bars + 8u * (2 * PSTAGES + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (2 * PSTAGES + 2 + b); };
  const uint32_t tmem_ptr_smem = bars + 8u * (2 * PSTAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < shape.num_segments; ++s) {
      prefetch_tensormap(&maps.a[s]);
      prefetch_tensormap(&maps.w[s]);
    }
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 2 * PBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    auto load_w = [&](int st, int s, int kb, int n_blk) {
      // W box is 128 rows: two boxes for a 256-wide tile
#pragma unroll
      for (int h = 0; h < PBN / 128; ++h)
        tma_load_2d(smem_b + st * C::B_STAGE_BYTES + h * (128 * BK * 2), &maps.w[s], full_bar(st), kb * BK,
                    n_blk * PBN + h * 128);
    };
    // weights of the first ring pass before the wait for the predecessor grid (see the one-tile kernel)
    int pre = 0;
    if (shape.early_w && static_cast<int>(blockIdx.x) < num_tiles) {
      const int n_blk0 = static_cast<int>(blockIdx.x) % shape.n_tiles;
      for (int s = 0; s < shape.num_segments && pre < PSTAGES; ++s)
        for (int kb = 0; kb < shape.kblocks[s] && pre < PSTAGES; ++kb, ++pre) {
          if (elect_one()) {
            mbar_arrive_expect_tx(full_bar(pre), C::STAGE_BYTES);
            load_w(pre, s, kb, n_blk0);
          }
          __syncwarp();
        }
    }
    pdl_wait();
    int stage = 0, issued = 0;
    uint32_t phase = 0;
    int ev = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile % shape.n_tiles, m_blk = tile / shape.n_tiles;
      for (int s = 0; s < shape.num_segments; ++s) {
        for (int kb = 0; kb < shape.kblocks[s]; ++kb, ++issued) {
          if (issued >= pre) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            GC_GTRACE(0, ev); ++ev;
          }
          if (elect_one()) {
            if (issued >= pre) {
              mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
              load_w(stage, s, kb, n_blk);
            }
            tma_load_2d(smem_a + stage * A_STAGE_BYTES, &maps.a[s], full_bar(stage), kb * BK, m_blk * BM);
          }
          __syncwarp();
          if (++stage == PSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16_f32(BM, PBN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt & 1;
      GC_GTRACE(1, 4 * lt);
      mbar_wait(acc_empty(b), ((lt >> 1) & 1) ^ 1u);      // epilogue has drained this accumulator buffer
      GC_GTRACE(1, 4 * lt + 1);
      tc_fence_after();
      uint32_t accumulate = 0;
      for (int s = 0; s < shape.num_segments; ++s) {
        for (int kb = 0; kb < shape.kblocks[s]; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t da = desc_kmajor_sw128(smem_a + stage * A_STAGE_BYTES);
          const uint64_t db = desc_kmajor_sw128(smem_b + stage * C::B_STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_f16(tmem_base + b * PBN, da + 2u * k, db + 2u * k, idesc, (accumulate | static_cast<uint32_t>(k)) != 0u ? 1u : 0u);
            umma_commit(empty_bar(stage));
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == PSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (elect_one()) umma_commit(acc_full(b));
      __syncwarp();
      GC_GTRACE(1, 4 * lt + 2);
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;             // 0..255 within the epilogue group
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    pdl_wait();
    const float alpha = ep.alpha_dev != nullptr ? __ldg(ep.alpha_dev) : 1.0f;
    const uint32_t stage_smem = smem_base + C::STAGING_OFFSET + static_cast<uint32_t>(warp - 2) * 8192u;
    const uint32_t gather_smem = (C::GATHER_BYTES != 0 && shape.gather_staged)
                                     ? smem_base + C::GATHER_OFFSET + static_cast<uint32_t>(warp - 2) * 2048u : 0u;
    int stage_use = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int n_blk = tile % shape.n_tiles, m_blk = tile / shape.n_tiles;
      const int b = lt & 1;
      // this tile's bias slice -> shared memory while the tensor core is still busy with it
      if (et < PBN) bias_s[b * PBN + et] = ep.bias != nullptr ? __ldg(ep.bias + n_blk * PBN + et) : 0.0f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float* bias_ptr = ep.bias != nullptr ? bias_s + b * PBN + half * HALF : nullptr;
      const int col_base = n_blk * PBN + half * HALF;
      const uint32_t taddr = tmem_base + b * PBN + half * HALF + lane_addr;
      if (threadIdx.x == 64) GC_GTRACE(2, 4 * lt);
      const uint32_t acc_bar = acc_empty(b), full_bar_b = acc_full(b);
      const uint32_t full_parity = (lt >> 1) & 1;
      epilogue_warp_tile<HALF>(ep, &maps.out, shape.store_mode, alpha, taddr, bias_ptr,
                               static_cast<int64_t>(m_blk) * BM + q * 32, col_base, lane, stage_smem, stage_use, gather_smem,
                               [&]() { mbar_wait(full_bar_b, full_parity); tc_fence_after(); },
                               [&]() { if (lane == 0) mbar_arrive(acc_bar); });
      if (threadIdx.x == 64) GC_GTRACE(2, 4 * lt + 2);
    }
    if (lane == 0) bulk_wait_group_all();        // staged stores have landed before the grid completes
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * PBN);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05.mma.cta_group::2) for the large problems: a cluster of two CTAs owns a
// 256 x 256 output tile.  Each CTA stages its 128 rows of A and its 128-row half of the W tile,
// the leader issues M = 256 MMAs that read both halves, each CTA's TMEM receives its 128 x 256
// slice.  Per SM this halves the W bytes pulled from L2 per FLOP (32 KB per k-block for 128 x 256
// outputs instead of 48 KB), which is what bounds the single-CTA kernels.  Same persistent loop,
// double-buffered accumulator and eight epilogue warps per CTA as above.
// ---------------------------------------------------------------------------------------------
constexpr int QBN = 256;                        // tile width (both CTAs)
constexpr int Q_B_STAGE_BYTES = 128 * BK * 2;   // this CTA's half of the W tile
constexpr int Q_STAGE_BYTES = A_STAGE_BYTES + Q_B_STAGE_BYTES;
constexpr int SMEM_OPT_IN_MAX = 232448;         // 227 KB per CTA
// Shared-memory plan: Q_STAGES operand stages of 32 KB | 8 warps x NBUF x 4 KB store staging | (GATHER) 8 warps x 2 KB
// gather transposition | barriers | bias.  The operand ring is what bounds the kernel: a TMA load takes ~2 300 clk
// from issue to arrival under load while the MMAs of one stage take 512, so the tensor pipe runs at
// min(1, STAGES * 512 / 2 300) of its rate (measured: operands arrive in bursts of STAGES).
template <int STAGES, int NBUF, bool GATHER>
struct PairCfg {
  static constexpr int Q_STAGES = STAGES;
  static constexpr int STAGING_OFFSET = STAGES * Q_STAGE_BYTES;
  static constexpr int GATHER_OFFSET = STAGING_OFFSET + 8 * NBUF * 4096;
  static constexpr int BAR_OFFSET = GATHER_OFFSET + (GATHER ? 8 * 2048 : 0);
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
  static constexpr int USED_BYTES = BIAS_OFFSET + 2 * QBN * 4;
  // slack for rounding the dynamic shared memory base up to 1024 B (the kernel traps if it does not suffice)
  static constexpr int SLACK = USED_BYTES + 1024 <= SMEM_OPT_IN_MAX ? 1024 : SMEM_OPT_IN_MAX - USED_BYTES;
  static constexpr int SMEM_BYTES = USED_BYTES + SLACK;
  static_assert(SLACK >= 0, "pair kernel configuration does not fit in shared memory");
};

template <int STAGES, int NBUF, bool GATHER>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ GemmMaps maps, const GemmShape shape, const EpilogueParams ep,
                              const int full_units, const int num_units, long long* trace) {
  using namespace sm100;
  pdl_launch_dependents();
  constexpr int HALF = QBN / 2;                 // columns per epilogue warp
  using C = PairCfg<STAGES, NBUF, GATHER>;
  constexpr int Q_STAGES = C::Q_STAGES;
  constexpr int Q_STAGING_OFFSET = C::STAGING_OFFSET, Q_GATHER_OFFSET = C::GATHER_OFFSET;
  constexpr int Q_BAR_OFFSET = C::BAR_OFFSET, Q_BIAS_OFFSET = C::BIAS_OFFSET;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (smem_base - smem_u32(smem_raw) > static_cast<uint32_t>(C::SLACK)) __trap();
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + Q_STAGES * A_STAGE_BYTES;
  const uint32_t bars = smem_base + Q_BAR_OFFSET;
  float* bias_s = reinterpret_cast<float*>(smem_gen + Q_BIAS_OFFSET);      // [2][QBN]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (Q_STAGES + s); };
  auto acc_full = [&](int b) { return bars + 8u * (2 * Q_STAGES + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (2 * Q_STAGES + 2 + b); };
  const uint32_t tmem_ptr_smem = bars + 8u * (2 * Q_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // Work units: the first `full_units` are the 256 x 256 tiles of the complete rounds (tile = unit).  The tiles of a
  // last, partial round that would leave more than half of the pairs idle are issued as two 256 x 128 half-width units
  // each (same A rows, half of the W rows), so that round takes half as long: 324 tiles on 74 pairs (the N = 512
  // mesh-side GEMMs) are 4.5 rounds instead of 5.  Which columns a unit covers is a pure function of the problem size,
  // so results stay bitwise reproducible.
  auto decode = [&](int u, int& m_pair, int& col0, int& width) {
    int tile = u;
    width = QBN;
    int sub = 0;
    if (u >= full_units) {
      const int h = u - full_units;
      tile = full_units + (h >> 1);
      sub = h & 1;
      width = QBN / 2;
    }
    m_pair = tile / shape.n_tiles;
    col0 = (tile % shape.n_tiles) * QBN + sub * (QBN / 2);
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < shape.num_segments; ++s) {
      prefetch_tensormap(&maps.a[s]);
      prefetch_tensormap(&maps.w[s]);
    }
    for (int s = 0; s < Q_STAGES; ++s) {
      mbar_init(full_bar(s), 1);                // leader's is the one that counts: 1 arrive + both CTAs' bytes
      mbar_init(empty_bar(s), 1);               // one multicast commit from the leader
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);                // multicast commit from the leader
      mbar_init(acc_empty(b), 16);              // 8 epilogue warps of each CTA (leader's copy is used)
    }
    fence_mbar_init();
  }
  __syncwarp();
  cluster_sync_all();                           // both CTAs' barriers exist before anything remote happens
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr_smem, 2 * QBN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // weights of the first ring pass before the wait for the predecessor grid (see the one-tile kernel)
    int pre = 0;
    if (shape.early_w && pair_id < num_units) {
      int m0, c0, w0;
      decode(pair_id, m0, c0, w0);
      const int n_row0 = c0 + static_cast<int>(rank) * (w0 / 2);
      for (int s = 0; s < shape.num_segments && pre < Q_STAGES; ++s)
        for (int kb = 0; kb < shape.kblocks[s] && pre < Q_STAGES; ++kb, ++pre) {
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(full_bar(pre), 2 * Q_STAGE_BYTES);
            tma_load_2d_pair(smem_b + pre * Q_B_STAGE_BYTES, &maps.w[s], full_bar(pre), kb * BK, n_row0);
          }
          __syncwarp();
        }
    }
    pdl_wait();
    int stage = 0, issued = 0;
    uint32_t phase = 0;
    int ev = 0;
    for (int unit = pair_id; unit < num_units; unit += num_pairs) {
      int m_pair, col0, width;
      decode(unit, m_pair, col0, width);
      const int m_row = m_pair * 256 + static_cast<int>(rank) * 128;
      // each CTA stages its half of the unit's W rows; the box is always 128 rows (a half-width unit uses the first 64)
      const int n_row = col0 + static_cast<int>(rank) * (width / 2);
      for (int s = 0; s < shape.num_segments; ++s) {
        for (int kb = 0; kb < shape.kblocks[s]; ++kb, ++issued) {
          if (issued >= pre) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            GC_GTRACE(0, ev); ++ev;
          }
          if (elect_one()) {
            if (issued >= pre) {
              // the leader arms its barrier for the bytes of both CTAs
              if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Q_STAGE_BYTES);
              tma_load_2d_pair(smem_b + stage * Q_B_STAGE_BYTES, &maps.w[s], full_bar(stage), kb * BK, n_row);
            }
            tma_load_2d_pair(smem_a + stage * A_STAGE_BYTES, &maps.a[s], full_bar(stage), kb * BK, m_row);
          }
          __syncwarp();
          if (++stage == Q_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc_full = idesc_bf16_f32(256, QBN, 0, 0);
      constexpr uint32_t idesc_half = idesc_bf16_f32(256, QBN / 2, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      int mev = 0;
      for (int unit = pair_id; unit < num_units; unit += num_pairs, ++lt) {
        const uint32_t idesc = unit < full_units ? idesc_full : idesc_half;
        const int b = lt & 1;
        GC_GTRACE(1, 4 * lt);
        mbar_wait(acc_empty(b), ((lt >> 1) & 1) ^ 1u);      // both CTAs' epilogues have drained this buffer
        GC_GTRACE(1, 4 * lt + 1);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int s = 0; s < shape.num_segments; ++s) {
          for (int kb = 0; kb < shape.kblocks[s]; ++kb) {
            mbar_wait(full_bar(stage), phase);
            GC_GTRACE(3, mev); ++mev;
            tc_fence_after();
            const uint64_t da = desc_kmajor_sw128(smem_a + stage * A_STAGE_BYTES);
            const uint64_t db = desc_kmajor_sw128(smem_b + stage * Q_B_STAGE_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_f16_pair(tmem_base + b * QBN, da + 2u * k, db + 2u * k, idesc, (accumulate | static_cast<uint32_t>(k)) != 0u ? 1u : 0u);
              umma_commit_pair(empty_bar(stage));               // frees the stage in both CTAs
            }
            __syncwarp();
            accumulate = 1;
            if (++stage == Q_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        if (elect_one()) umma_commit_pair(acc_full(b));
        __syncwarp();
        GC_GTRACE(1, 4 * lt + 2);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;             // 0..255 within the epilogue group
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    pdl_wait();
    const float alpha = ep.alpha_dev != nullptr ? __ldg(ep.alpha_dev) : 1.0f;
    const uint32_t stage_smem = smem_base + Q_STAGING_OFFSET + static_cast<uint32_t>(warp - 2) * (NBUF * 4096u);
    const uint32_t gather_smem = (GATHER && shape.gather_staged) ? smem_base + Q_GATHER_OFFSET + static_cast<uint32_t>(warp - 2) * 2048u : 0u;
    int stage_use = 0;
    int etrace_ev = 0;
    int lt = 0;
    for (int unit = pair_id; unit < num_units; unit += num_pairs, ++lt) {
      int m_pair, col0, width;
      decode(unit, m_pair, col0, width);
      const int b = lt & 1;
      if (et < width) bias_s[b * QBN + et] = ep.bias != nullptr ? __ldg(ep.bias + col0 + et) : 0.0f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int wcols = width / 2;                 // columns per epilogue warp: 128, or 64 in a half-width unit
      const float* bias_ptr = ep.bias != nullptr ? bias_s + b * QBN + half * wcols : nullptr;
      const int col_base = col0 + half * wcols;
      const uint32_t taddr = tmem_base + b * QBN + half * wcols + lane_addr;
      const uint32_t acc_bar = acc_empty(b), full_bar_b = acc_full(b);
      const uint32_t full_parity = (lt >> 1) & 1;
      const int64_t row0 = static_cast<int64_t>(m_pair) * 256 + rank * 128 + q * 32;
      auto wait_acc = [&]() {
        if (threadIdx.x == 64) GC_GTRACE(2, 4 * lt);
        mbar_wait(full_bar_b, full_parity);
        if (threadIdx.x == 64) GC_GTRACE(2, 4 * lt + 1);
        tc_fence_after();
      };
      auto release = [&]() { if (lane == 0) mbar_arrive_cluster(acc_bar, 0); };
      long long* etr = (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) ? trace : nullptr;
      // one instantiation for both widths: in a half-width unit each warp walks 64 columns instead of 128
      epilogue_warp_tile<HALF, NBUF>(ep, &maps.out, shape.store_mode, alpha, taddr, bias_ptr, row0, col_base, lane, stage_smem, stage_use,
                               gather_smem, wait_acc, release, etr, &etrace_ev, wcols);
      if (threadIdx.x == 64) GC_GTRACE(2, 4 * lt + 2);
    }
    if (lane == 0) bulk_wait_group_all();
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();                           // the peer may still be signalling / reading this CTA
  if (warp == 1) tmem_dealloc_pair(tmem_base, 2 * QBN);
}

// ---------------------------------------------------------------------------------------------
// A-resident CTA-pair variant for single-segment problems with K <= 512 (the QKV / out-projection /
// FFW-in GEMMs of the transformer and the K = L node / edge MLP layers): an experiment, opt-in.  The
// 256 x 256 pair tile pulls 32 KB per CTA per 64-wide k-block from L2, i.e. 64 B/clk/SM at the tensor
// peak, 9.5 KB/clk chip-wide.  Here each pair walks a contiguous range of the (m, n) tile order
// with n fastest, keeps its 128 rows x K of A in shared memory (8 x 16 KB) for all the n-tiles of that
// row block and streams only its half of the W tile (16 KB per k-block): 32 B/clk/SM.  A k-block slots
// are released one by one during the last n-tile of a row block, so the next block's A streams in under
// the running MMAs.  W ring of 3 stages, one 4 KB store-staging buffer per epilogue warp.
// ---------------------------------------------------------------------------------------------
constexpr int R_KB_MAX = 8;
constexpr int R_STAGES = 3;
constexpr int R_A_BYTES = R_KB_MAX * A_STAGE_BYTES;                  // 128 KB
constexpr int R_B_STAGE_BYTES = 128 * BK * 2;
constexpr int R_STAGING_OFFSET = R_A_BYTES + R_STAGES * R_B_STAGE_BYTES;
constexpr int R_BAR_OFFSET = R_STAGING_OFFSET + 8 * 4096;
constexpr int R_BIAS_OFFSET = R_BAR_OFFSET + 256;
constexpr int R_SMEM_BYTES = R_BIAS_OFFSET + 2 * QBN * 4 + 1024;

__global__ void __launch_bounds__(P_THREADS, 1)
gemm_bf16_tcgen05_pair_resident_kernel(const __grid_constant__ GemmMaps maps, const GemmShape shape, const EpilogueParams ep,
                                       const int num_tiles) {
  using namespace sm100;
  pdl_launch_dependents();
  constexpr int HALF = QBN / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + R_A_BYTES;
  const uint32_t bars = smem_base + R_BAR_OFFSET;
  float* bias_s = reinterpret_cast<float*>(smem_gen + R_BIAS_OFFSET);      // [2][QBN]
  auto a_full = [&](int kb) { return bars + 8u * kb; };
  auto a_empty = [&](int kb) { return bars + 8u * (R_KB_MAX + kb); };
  auto b_full = [&](int s) { return bars + 8u * (2 * R_KB_MAX + s); };
  auto b_empty = [&](int s) { return bars + 8u * (2 * R_KB_MAX + R_STAGES + s); };
  auto acc_full = [&](int b) { return bars + 8u * (2 * R_KB_MAX + 2 * R_STAGES + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (2 * R_KB_MAX + 2 * R_STAGES + 2 + b); };
  const uint32_t tmem_ptr_smem = bars + 8u * (2 * R_KB_MAX + 2 * R_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int KB = shape.kblocks[0];
  const int t_begin = static_cast<int>(static_cast<int64_t>(pair_id) * num_tiles / num_pairs);
  const int t_end = static_cast<int>(static_cast<int64_t>(pair_id + 1) * num_tiles / num_pairs);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.a[0]);
    prefetch_tensormap(&maps.w[0]);
    for (int kb = 0; kb < R_KB_MAX; ++kb) { mbar_init(a_full(kb), 1); mbar_init(a_empty(kb), 1); }
    for (int s = 0; s < R_STAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 16); }
    fence_mbar_init();
  }
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr_smem, 2 * QBN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      int stage = 0, group = -1, prev_m = -1;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int n_blk = tile % shape.n_tiles, m_pair = tile / shape.n_tiles;
        const bool new_group = m_pair != prev_m;
        if (new_group) { ++group; prev_m = m_pair; }
        const int m_row = m_pair * 256 + static_cast<int>(rank) * 128;
        const int n_row = n_blk * QBN + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < KB; ++kb) {
          if (new_group) {
            mbar_wait(a_empty(kb), (static_cast<uint32_t>(group) & 1u) ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(a_full(kb), 2 * A_STAGE_BYTES);
            tma_load_2d_pair(smem_a + kb * A_STAGE_BYTES, &maps.a[0], a_full(kb), kb * BK, m_row);
          }
          mbar_wait(b_empty(stage), phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(b_full(stage), 2 * R_B_STAGE_BYTES);
          tma_load_2d_pair(smem_b + stage * R_B_STAGE_BYTES, &maps.w[0], b_full(stage), kb * BK, n_row);
          if (++stage == R_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(256, QBN, 0, 0);
      int stage = 0, group = -1, prev_m = -1, lt = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
        const int m_pair = tile / shape.n_tiles;
        const bool new_group = m_pair != prev_m;
        if (new_group) { ++group; prev_m = m_pair; }
        const bool last_of_group = tile + 1 == t_end || (tile + 1) / shape.n_tiles != m_pair;
        const int b = lt & 1;
        mbar_wait(acc_empty(b), ((lt >> 1) & 1) ^ 1u);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int kb = 0; kb < KB; ++kb) {
          if (new_group) mbar_wait(a_full(kb), static_cast<uint32_t>(group) & 1u);
          mbar_wait(b_full(stage), phase);
          tc_fence_after();
          const uint64_t da = desc_kmajor_sw128(smem_a + kb * A_STAGE_BYTES);
          const uint64_t db = desc_kmajor_sw128(smem_b + stage * R_B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_f16_pair(tmem_base + b * QBN, da + 2u * k, db + 2u * k, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit_pair(b_empty(stage));
          if (last_of_group) umma_commit_pair(a_empty(kb));     // this A k-block is free in both CTAs
          if (++stage == R_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(acc_full(b));
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    pdl_wait();
    const float alpha = ep.alpha_dev != nullptr ? __ldg(ep.alpha_dev) : 1.0f;
    const uint32_t stage_smem = smem_base + R_STAGING_OFFSET + static_cast<uint32_t>(warp - 2) * 4096u;
    int stage_use = 0;
    int lt = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      const int n_blk = tile % shape.n_tiles, m_pair = tile / shape.n_tiles;
      const int b = lt & 1;
      if (et < QBN) bias_s[b * QBN + et] = ep.bias != nullptr ? __ldg(ep.bias + n_blk * QBN + et) : 0.0f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float* bias_ptr = ep.bias != nullptr ? bias_s + b * QBN + half * HALF : nullptr;
      const int col_base = n_blk * QBN + half * HALF;
      const uint32_t taddr = tmem_base + b * QBN + half * HALF + lane_addr;
      const uint32_t acc_bar = acc_empty(b), full_bar_b = acc_full(b);
      const uint32_t full_parity = (lt >> 1) & 1;
      epilogue_warp_tile<HALF, 1>(ep, &maps.out, shape.store_mode, alpha, taddr, bias_ptr,
                                  static_cast<int64_t>(m_pair) * 256 + rank * 128 + q * 32, col_base, lane, stage_smem, stage_use,
                                  0u, [&]() { mbar_wait(full_bar_b, full_parity); tc_fence_after(); },
                                  [&]() { if (lane == 0) mbar_arrive_cluster(acc_bar, 0); });
    }
    if (lane == 0) bulk_wait_group_all();
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 2 * QBN);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return GC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return GC_ERR_CUDA;
  }
  return GC_OK;
}

int make_tmap_out_2d(CUtensorMap* out, void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return GC_ERR_CUDA;
  }
  const uint64_t esize = dtype == GC_BF16 ? 2 : 4;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esize), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype == GC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(out) failed with CUresult %d (rows=%llu cols=%llu ld=%llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return GC_ERR_CUDA;
  }
  return GC_OK;
}

bool early_w_enabled() {
  static const bool on = []() {
    const char* v = getenv("GENCAST_EARLY_W");
    return !(v != nullptr && v[0] == '0');
  }();
  return on;
}

template <int BN, int NSTAGES>
int launch_cfg(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep) {
  using C = GemmCfg<BN, NSTAGES>;
  GemmMaps maps;
  GemmShape shape;
  shape.num_segments = a.num_segments;
  shape.early_w = ((a.flags & GC_GEMM_STATIC_WEIGHTS) && early_w_enabled()) ? 1 : 0;
  shape.store_mode = STORE_DIRECT;
  shape.gather_staged = 0;
  shape.n_tiles = a.n / BN;
  for (int s = 0; s < GC_MAX_SEGMENTS; ++s) shape.kblocks[s] = 0;
  for (int s = 0; s < a.num_segments; ++s) {
    shape.kblocks[s] = a.k[s] / BK;
    int rc = make_tmap_bf16_2d(&maps.a[s], a.a[s], (uint64_t)a.m, (uint64_t)a.k[s], (uint64_t)a.lda[s], BK, BM);
    if (rc != GC_OK) return rc;
    rc = make_tmap_bf16_2d(&maps.w[s], a.w[s], (uint64_t)a.n, (uint64_t)a.k[s], (uint64_t)a.ldw[s], BK, BN);
    if (rc != GC_OK) return rc;
  }
  for (int s = a.num_segments; s < GC_MAX_SEGMENTS; ++s) {
    maps.a[s] = maps.a[0];
    maps.w[s] = maps.w[0];
  }
  cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, NSTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel)");
  const int64_t m_tiles = (a.m + BM - 1) / BM;
  const int64_t grid = m_tiles * shape.n_tiles;
  if (grid > 0x7fffffffLL) {
    set_error("gc_gemm: too many tiles (%lld)", (long long)grid);
    return GC_ERR_INVALID_ARGUMENT;
  }
  GC_CHECK_CUDA(launch_kernel(gemm_bf16_tcgen05_kernel<BN, NSTAGES>, dim3((unsigned)grid), dim3(NUM_THREADS),
                              (size_t)C::SMEM_BYTES, stream, maps, shape, ep), "gemm_bf16_tcgen05_kernel");
  return GC_OK;
}

int fill_out_map(GemmMaps& maps, GemmShape& shape, const gc_gemm_args& a, bool can_stage_gathers);

template <int PBN>
int launch_persistent(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep) {
  using C = PersistCfg<PBN>;
  GemmMaps maps;
  GemmShape shape;
  shape.num_segments = a.num_segments;
  shape.early_w = ((a.flags & GC_GEMM_STATIC_WEIGHTS) && early_w_enabled()) ? 1 : 0;
  shape.n_tiles = a.n / PBN;
  for (int s = 0; s < GC_MAX_SEGMENTS; ++s) shape.kblocks[s] = 0;
  for (int s = 0; s < a.num_segments; ++s) {
    shape.kblocks[s] = a.k[s] / BK;
    int rc = make_tmap_bf16_2d(&maps.a[s], a.a[s], (uint64_t)a.m, (uint64_t)a.k[s], (uint64_t)a.lda[s], BK, BM);
    if (rc != GC_OK) return rc;
    rc = make_tmap_bf16_2d(&maps.w[s], a.w[s], (uint64_t)a.n, (uint64_t)a.k[s], (uint64_t)a.ldw[s], BK, 128);
    if (rc != GC_OK) return rc;
  }
  for (int s = a.num_segments; s < GC_MAX_SEGMENTS; ++s) {
    maps.a[s] = maps.a[0];
    maps.w[s] = maps.w[0];
  }
  {
    const int rc = fill_out_map(maps, shape, a, C::GATHER_BYTES != 0);
    if (rc != GC_OK) return rc;
  }
  GC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_persistent_kernel<PBN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES), "cudaFuncSetAttribute(gemm_bf16_tcgen05_persistent_kernel)");
  const int64_t num_tiles = ((a.m + BM - 1) / BM) * shape.n_tiles;
  if (num_tiles > 0x7fffffffLL) {
    set_error("gc_gemm: too many tiles (%lld)", (long long)num_tiles);
    return GC_ERR_INVALID_ARGUMENT;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(num_tiles < sms ? num_tiles : sms);
  GC_CHECK_CUDA(launch_kernel(gemm_bf16_tcgen05_persistent_kernel<PBN>, dim3(grid), dim3(P_THREADS), (size_t)C::SMEM_BYTES,
                              stream, maps, shape, ep, (int)num_tiles, g_gemm_trace), "gemm_bf16_tcgen05_persistent_kernel");
  return GC_OK;
}

bool pair_kernel_enabled() {
  static const bool on = []() {
    const char* v = getenv("GENCAST_GEMM_PAIR");
    return !(v != nullptr && v[0] == '0');
  }();
  return on;
}

bool tail_split_enabled() {
  static const bool on = []() {
    const char* v = getenv("GENCAST_GEMM_TAIL_SPLIT");
    return !(v != nullptr && v[0] == '0');
  }();
  return on;
}

int pair_stages() {
  static const int n = []() {
    const char* v = getenv("GENCAST_GEMM_STAGES");
    return (v != nullptr && v[0] >= '4' && v[0] <= '6') ? v[0] - '0' : 6;
  }();
  return n;
}

int launch_pair(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep) {
  GemmMaps maps;
  GemmShape shape;
  shape.num_segments = a.num_segments;
  shape.early_w = ((a.flags & GC_GEMM_STATIC_WEIGHTS) && early_w_enabled()) ? 1 : 0;
  shape.n_tiles = a.n / QBN;
  for (int s = 0; s < GC_MAX_SEGMENTS; ++s) shape.kblocks[s] = 0;
  for (int s = 0; s < a.num_segments; ++s) {
    shape.kblocks[s] = a.k[s] / BK;
    int rc = make_tmap_bf16_2d(&maps.a[s], a.a[s], (uint64_t)a.m, (uint64_t)a.k[s], (uint64_t)a.lda[s], BK, BM);
    if (rc != GC_OK) return rc;
    rc = make_tmap_bf16_2d(&maps.w[s], a.w[s], (uint64_t)a.n, (uint64_t)a.k[s], (uint64_t)a.ldw[s], BK, 128);
    if (rc != GC_OK) return rc;
  }
  for (int s = a.num_segments; s < GC_MAX_SEGMENTS; ++s) {
    maps.a[s] = maps.a[0];
    maps.w[s] = maps.w[0];
  }
  {
    const int rc = fill_out_map(maps, shape, a, true);
    if (rc != GC_OK) return rc;
  }
  const int64_t num_tiles = ((a.m + 255) / 256) * shape.n_tiles;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t pairs = sms / 2;
  // tiles of a last round that would leave at least half of the pairs idle become two half-width units each
  int64_t full_units = num_tiles, num_units = num_tiles;
  if (tail_split_enabled() && num_tiles > pairs) {
    const int64_t rest = num_tiles % pairs;
    if (rest > 0 && 2 * rest <= pairs) {
      full_units = num_tiles - rest;
      num_units = full_units + 2 * rest;
    }
  }
  if (num_units < pairs) pairs = num_units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(P_THREADS);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // ring depth: 6 stages (one staging buffer per warp, no gather transposition buffers) unless the epilogue gathers
  // rows through shared memory, then 5; GENCAST_GEMM_STAGES=4 selects the round-1 layout (4 stages, 2 staging buffers)
  const int stages = pair_stages();
  auto launch = [&](auto kernel, int smem_bytes) -> int {
    GC_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes),
                  "cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_kernel)");
    cfg.dynamicSmemBytes = smem_bytes;
    GC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, shape, ep, (int)full_units, (int)num_units, g_gemm_trace),
                  "gemm_bf16_tcgen05_pair_kernel");
    return GC_OK;
  };
  if (stages <= 4) return launch(gemm_bf16_tcgen05_pair_kernel<4, 2, true>, PairCfg<4, 2, true>::SMEM_BYTES);
  if (stages == 5 || shape.gather_staged) return launch(gemm_bf16_tcgen05_pair_kernel<5, 1, true>, PairCfg<5, 1, true>::SMEM_BYTES);
  return launch(gemm_bf16_tcgen05_pair_kernel<6, 1, false>, PairCfg<6, 1, false>::SMEM_BYTES);
}

// Opt-in (GENCAST_GEMM_ARES=1): measured on B200 at M = 40 968 it halves the L2 -> SM operand bytes but is
// time-neutral (QKV 65.0 vs 60.8 us, FFW-in 89.9 vs 92.2 us, M = 260 640 N = K = 512: 138.5 vs 140.6 us), i.e. the
// operand feed is not what holds the pair kernel at ~87 % MMA issue inside a tile; see profiles/README.md.
bool resident_kernel_enabled() {
  static const bool on = []() {
    const char* v = getenv("GENCAST_GEMM_ARES");
    return v != nullptr && v[0] == '1';
  }();
  return on;
}

int launch_pair_resident(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep) {
  GemmMaps maps;
  GemmShape shape;
  shape.num_segments = 1;
  shape.early_w = 0;
  shape.n_tiles = a.n / QBN;
  for (int s = 0; s < GC_MAX_SEGMENTS; ++s) shape.kblocks[s] = 0;
  shape.kblocks[0] = a.k[0] / BK;
  int rc = make_tmap_bf16_2d(&maps.a[0], a.a[0], (uint64_t)a.m, (uint64_t)a.k[0], (uint64_t)a.lda[0], BK, BM);
  if (rc != GC_OK) return rc;
  rc = make_tmap_bf16_2d(&maps.w[0], a.w[0], (uint64_t)a.n, (uint64_t)a.k[0], (uint64_t)a.ldw[0], BK, 128);
  if (rc != GC_OK) return rc;
  for (int s = 1; s < GC_MAX_SEGMENTS; ++s) {
    maps.a[s] = maps.a[0];
    maps.w[s] = maps.w[0];
  }
  rc = fill_out_map(maps, shape, a, false);
  if (rc != GC_OK) return rc;
  GC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     R_SMEM_BYTES), "cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_resident_kernel)");
  const int64_t num_tiles = ((a.m + 255) / 256) * shape.n_tiles;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t pairs = sms / 2;
  if (num_tiles < pairs) pairs = num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = R_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  GC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_pair_resident_kernel, maps, shape, ep, (int)num_tiles),
                "gemm_bf16_tcgen05_pair_resident_kernel");
  return GC_OK;
}

// STORE_TMA / STORE_TMA_ADD when the fused epilogue is bias + activation (+ in-place fp32 residual) only.
bool gathers_stageable(const gc_gemm_args& a) {
  // bf16 tables, first slot used, 16-byte aligned rows (checked by gc_gemm), 256-wide column blocks stay inside a row
  return a.gather_src[0] != nullptr && a.gather_dtype == GC_BF16;
}

int choose_store_mode(const gc_gemm_args& a, bool can_stage_gathers) {
  const bool has_gather = a.gather_src[0] != nullptr || a.gather_src[1] != nullptr;
  if (a.addend != nullptr) return STORE_DIRECT;
  if (has_gather && !(can_stage_gathers && gathers_stageable(a))) return STORE_DIRECT;
  if (a.residual == nullptr) return STORE_TMA;
  if (a.residual == a.out && a.res_dtype == GC_F32 && a.out_dtype == GC_F32 && a.ld_res == a.ldo) return STORE_TMA_ADD;
  return STORE_DIRECT;
}

int fill_out_map(GemmMaps& maps, GemmShape& shape, const gc_gemm_args& a, bool can_stage_gathers) {
  shape.store_mode = choose_store_mode(a, can_stage_gathers);
  shape.gather_staged = (shape.store_mode != STORE_DIRECT && a.gather_src[0] != nullptr) ? 1 : 0;
  if (shape.store_mode == STORE_DIRECT) {
    maps.out = maps.a[0];
    return GC_OK;
  }
  return make_tmap_out_2d(&maps.out, a.out, a.out_dtype, (uint64_t)a.m, (uint64_t)a.n, (uint64_t)a.ldo, 32);
}

int launch_gemm_tcgen05(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep) {

  // Fewer than ~2 CTAs per SM with 128-wide tiles: halve the tile width to spread the work.
  const int64_t tiles128 = ((a.m + BM - 1) / BM) * (a.n / 128);
  if (tiles128 >= 2 * 148) {
    // 256-wide tiles when N allows it and there are still >= 2 tiles per SM
    if (a.n % 256 == 0 && tiles128 >= 4 * 148) {
      if (pair_kernel_enabled()) {
        const bool has_gather = a.gather_src[0] != nullptr || a.gather_src[1] != nullptr;
        if (resident_kernel_enabled() && a.num_segments == 1 && a.k[0] <= R_KB_MAX * BK && a.n / QBN >= 2 && !has_gather &&
            a.addend == nullptr)
          return launch_pair_resident(stream, a, ep);
        return launch_pair(stream, a, ep);
      }
      return launch_persistent<256>(stream, a, ep);
    }
    return launch_persistent<128>(stream, a, ep);
  }
  // One CTA per SM at most: a deep ring (8 x 24 KB in flight) keeps the per-SM L2 link busy
  // through the long serial K loops of the skinny mesh-side GEMMs.
  if (2 * tiles128 <= 148) return launch_cfg<64, 8>(stream, a, ep);
  return launch_cfg<64, 4>(stream, a, ep);
}

}  // namespace gc

// Debug hook (not part of the public header): clock-stamp buffer of at least 8 * 512 int64.
extern "C" __attribute__((visibility("default"))) void gc_debug_set_gemm_trace(void* ptr) {
  gc::g_gemm_trace = reinterpret_cast<long long*>(ptr);
}
