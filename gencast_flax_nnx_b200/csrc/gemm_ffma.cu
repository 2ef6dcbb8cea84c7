// fp32 linear layer on the FFMA pipe (the fp32 parity mode of the denoiser; the
// throughput path is the bf16 tcgen05 kernel).  128 x 128 output tile per CTA,
// 256 threads, 8 x 8 outputs per thread, K staged 16 at a time through shared
// memory (transposed so the inner product reads are conflict free), register
// prefetch of the next K slab.  Same fused epilogue as the tensor-core kernel.
#include "common.cuh"

namespace gc {

namespace {

constexpr int FBM = 128;
constexpr int FBN = 128;
constexpr int FBK = 16;
constexpr int FPAD = 4;

struct FfmaArgs {
  const float* a[GC_MAX_SEGMENTS];
  const float* w[GC_MAX_SEGMENTS];
  int64_t lda[GC_MAX_SEGMENTS];
  int64_t ldw[GC_MAX_SEGMENTS];
  int k[GC_MAX_SEGMENTS];
  int num_segments;
  int n_tiles;
};

__global__ void __launch_bounds__(256, 2) gemm_f32_ffma_kernel(const FfmaArgs g, const EpilogueParams ep) {
  __shared__ __align__(16) float As[2][FBK][FBM + FPAD];
  __shared__ __align__(16) float Bs[2][FBK][FBN + FPAD];

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int n_blk = blockIdx.x % g.n_tiles;
  const int m_blk = blockIdx.x / g.n_tiles;
  const int64_t row0 = static_cast<int64_t>(m_blk) * FBM;
  const int col0 = n_blk * FBN;

  // loader mapping: 128 rows x 16 k = 512 float4; thread loads rows lr and lr + 64, k quad lk
  const int lr = tid >> 2;
  const int lk = (tid & 3) * 4;
  // compute mapping: rows {ty*4..+3, 64+ty*4..+3}, cols {tx*4..+3, 64+tx*4..+3}
  const int ty = tid >> 4;
  const int tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  int buf = 0;
  for (int s = 0; s < g.num_segments; ++s) {
    const float* A = g.a[s];
    const float* W = g.w[s];
    const int64_t lda = g.lda[s], ldw = g.ldw[s];
    const int nk = g.k[s] / FBK;

    float4 ra[2], rb[2];
    auto fetch = [&](int kb) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t r = row0 + lr + 64 * h;
        ra[h] = r < ep.m ? __ldg(reinterpret_cast<const float4*>(A + r * lda + kb * FBK + lk)) : make_float4(0, 0, 0, 0);
        const int c = col0 + lr + 64 * h;
        rb[h] = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(c) * ldw + kb * FBK + lk));
      }
    };
    auto stash = [&](int b) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lr + 64 * h;
        As[b][lk + 0][r] = ra[h].x; As[b][lk + 1][r] = ra[h].y; As[b][lk + 2][r] = ra[h].z; As[b][lk + 3][r] = ra[h].w;
        Bs[b][lk + 0][r] = rb[h].x; Bs[b][lk + 1][r] = rb[h].y; Bs[b][lk + 2][r] = rb[h].z; Bs[b][lk + 3][r] = rb[h].w;
      }
    };

    fetch(0);
    __syncthreads();   // previous segment's readers are done with both buffers
    stash(buf);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
      if (kb + 1 < nk) fetch(kb + 1);
#pragma unroll
      for (int kk = 0; kk < FBK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (kb + 1 < nk) {
        stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  const float alpha = ep.alpha_dev != nullptr ? __ldg(ep.alpha_dev) : 1.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= ep.m) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      epilogue_row_segment<4, false>(ep, alpha, row, col0 + h * 64 + tx * 4, v,
                                     ep.bias != nullptr ? ep.bias + col0 + h * 64 + tx * 4 : nullptr);
    }
  }
}

}  // namespace

int launch_gemm_ffma(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep) {
  FfmaArgs g;
  g.num_segments = a.num_segments;
  g.n_tiles = a.n / FBN;
  for (int s = 0; s < GC_MAX_SEGMENTS; ++s) {
    const bool on = s < a.num_segments;
    g.a[s] = on ? reinterpret_cast<const float*>(a.a[s]) : nullptr;
    g.w[s] = on ? reinterpret_cast<const float*>(a.w[s]) : nullptr;
    g.lda[s] = on ? a.lda[s] : 0;
    g.ldw[s] = on ? a.ldw[s] : 0;
    g.k[s] = on ? a.k[s] : 0;
  }
  const int64_t m_tiles = (a.m + FBM - 1) / FBM;
  const int64_t grid = m_tiles * g.n_tiles;
  if (grid > 0x7fffffffLL) {
    set_error("gc_gemm: too many tiles (%lld)", (long long)grid);
    return GC_ERR_INVALID_ARGUMENT;
  }
  GC_CHECK_CUDA(launch_kernel(gemm_f32_ffma_kernel, dim3((unsigned)grid), dim3(256), 0, stream, g, ep),
                "gemm_f32_ffma_kernel");
  return GC_OK;
}

}  // namespace gc
