// Exact k-hop neighbourhood attention over a CSR pattern, CUDA-core version.
//
// One warp per (node, head).  Scores: each lane takes one neighbour of a chunk of
// 32 and forms the full q.k dot product from a 128-bit vectorised read of that
// neighbour's key row; online softmax across the warp in fp32; values: every lane
// owns head_dim/32 output channels and the chunk's probabilities are broadcast with
// shuffles, so the value rows are read fully coalesced.  K and V of the whole mesh
// (<= 84 MB) live in the 126 MB L2, so this kernel is L2-gather bound; it is the
// fp32-parity path and the fallback for head dims the tensor-core kernel does not
// cover.
#include "common.cuh"

namespace gc {

namespace {

template <int D>
__global__ void __launch_bounds__(128) khop_attention_csr_kernel(const void* __restrict__ qkv, int dtype, int64_t ld,
                                                                 const int32_t* __restrict__ nbr_ptr,
                                                                 const int32_t* __restrict__ nbr_idx,
                                                                 void* __restrict__ out, int64_t ldo, int64_t nodes,
                                                                 int heads) {
  constexpr int DPL = D / 32;
  __shared__ __align__(16) float q_s[4][D];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t hd = static_cast<int64_t>(heads) * D;
  const float scale = rsqrtf(static_cast<float>(D));
  const int64_t items = nodes * heads;
  for (int64_t item0 = static_cast<int64_t>(blockIdx.x) * 4; item0 < items; item0 += static_cast<int64_t>(gridDim.x) * 4) {
    const int64_t item = item0 + warp;
    if (item < items) {
      const int64_t node = item / heads;
      const int head = static_cast<int>(item - node * heads);
      const int64_t qoff = node * ld + static_cast<int64_t>(head) * D;
      const int64_t koff = hd + static_cast<int64_t>(head) * D;
      const int64_t voff = 2 * hd + static_cast<int64_t>(head) * D;
      {
        float t[DPL];
        if constexpr (DPL == 1) {
          t[0] = dtype == GC_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(qkv)[qoff + lane])
                                  : reinterpret_cast<const float*>(qkv)[qoff + lane];
        } else if constexpr (DPL == 2) {
          if (dtype == GC_BF16) {
            float2 f = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(qkv) + qoff)[lane]);
            t[0] = f.x; t[1] = f.y;
          } else {
            float2 f = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(qkv) + qoff)[lane];
            t[0] = f.x; t[1] = f.y;
          }
        } else {
          load_as_float<DPL>(qkv, dtype, qoff + lane * DPL, t);
        }
#pragma unroll
        for (int i = 0; i < DPL; ++i) q_s[warp][lane * DPL + i] = t[i];
      }
      __syncwarp();
      const int beg = __ldg(nbr_ptr + node), end = __ldg(nbr_ptr + node + 1);
      float m = -INFINITY, l = 0.0f;
      float acc[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc[i] = 0.0f;
      for (int c0 = beg; c0 < end; c0 += 32) {
        const int cnt = min(32, end - c0);
        const bool valid = lane < cnt;
        const int j = valid ? __ldg(nbr_idx + c0 + lane) : 0;
        float s = 0.0f;
        const int64_t krow = static_cast<int64_t>(j) * ld + koff;
#pragma unroll
        for (int c = 0; c < D; c += 8) {
          float kv[8];
          load_as_float<8>(qkv, dtype, krow + c, kv);
          const float4 qa = *reinterpret_cast<const float4*>(&q_s[warp][c]);
          const float4 qb = *reinterpret_cast<const float4*>(&q_s[warp][c + 4]);
          s = fmaf(qa.x, kv[0], s); s = fmaf(qa.y, kv[1], s); s = fmaf(qa.z, kv[2], s); s = fmaf(qa.w, kv[3], s);
          s = fmaf(qb.x, kv[4], s); s = fmaf(qb.y, kv[5], s); s = fmaf(qb.z, kv[6], s); s = fmaf(qb.w, kv[7], s);
        }
        s = valid ? s * scale : -INFINITY;
        const float m_new = fmaxf(m, warp_max(s));
        const float p = valid ? __expf(s - m_new) : 0.0f;
        const float corr = __expf(m - m_new);   // m = -inf on the first chunk -> 0
        l = l * corr + warp_sum(p);
#pragma unroll
        for (int i = 0; i < DPL; ++i) acc[i] *= corr;
        m = m_new;
        for (int t = 0; t < cnt; ++t) {
          const float pj = __shfl_sync(0xffffffffu, p, t);
          const int jj = __shfl_sync(0xffffffffu, j, t);
          const int64_t vrow = static_cast<int64_t>(jj) * ld + voff + lane * DPL;
          if constexpr (DPL == 1) {
            const float v = dtype == GC_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(qkv)[vrow])
                                             : reinterpret_cast<const float*>(qkv)[vrow];
            acc[0] = fmaf(pj, v, acc[0]);
          } else if constexpr (DPL == 2) {
            float2 f;
            if (dtype == GC_BF16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(qkv) + vrow));
            else f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(qkv) + vrow);
            acc[0] = fmaf(pj, f.x, acc[0]); acc[1] = fmaf(pj, f.y, acc[1]);
          } else {
            float v[DPL];
            load_as_float<DPL>(qkv, dtype, vrow, v);
#pragma unroll
            for (int i = 0; i < DPL; ++i) acc[i] = fmaf(pj, v[i], acc[i]);
          }
        }
      }
      const float inv = l > 0.0f ? 1.0f / l : 0.0f;     // rows without neighbours (padding) produce zeros
      const int64_t ooff = node * ldo + static_cast<int64_t>(head) * D + lane * DPL;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        const float o = acc[i] * inv;
        if (dtype == GC_BF16) reinterpret_cast<__nv_bfloat16*>(out)[ooff + i] = __float2bfloat16_rn(o);
        else reinterpret_cast<float*>(out)[ooff + i] = o;
      }
    }
    __syncwarp();
  }
}

}  // namespace

int launch_khop_attention_csr(cudaStream_t st, const void* qkv, int dtype, int64_t ld_qkv, const int32_t* nbr_ptr,
                              const int32_t* nbr_idx, void* out, int64_t ldo, int64_t nodes, int heads, int head_dim) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t need = (nodes * heads + 3) / 4;
  const int64_t cap = static_cast<int64_t>(sms) * 16;
  const unsigned grid = static_cast<unsigned>(need < cap ? need : cap);
  if (head_dim == 32) GC_CHECK_CUDA(launch_kernel(khop_attention_csr_kernel<32>, dim3(grid), dim3(128), 0, st, qkv, dtype, ld_qkv, nbr_ptr, nbr_idx, out, ldo, nodes, heads), "khop_attention_csr_kernel");
  else if (head_dim == 64) GC_CHECK_CUDA(launch_kernel(khop_attention_csr_kernel<64>, dim3(grid), dim3(128), 0, st, qkv, dtype, ld_qkv, nbr_ptr, nbr_idx, out, ldo, nodes, heads), "khop_attention_csr_kernel");
  else if (head_dim == 128) GC_CHECK_CUDA(launch_kernel(khop_attention_csr_kernel<128>, dim3(grid), dim3(128), 0, st, qkv, dtype, ld_qkv, nbr_ptr, nbr_idx, out, ldo, nodes, heads), "khop_attention_csr_kernel");
  else {
    set_error("gc_khop_attention: head_dim=%d (supported: 32, 64, 128)", head_dim);
    return GC_ERR_UNSUPPORTED;
  }
  GC_CHECK_LAUNCH("khop_attention_csr_kernel");
  return GC_OK;
}

}  // namespace gc

extern "C" int gc_khop_attention(void* stream, const void* qkv, int32_t dtype, int64_t ld_qkv, const int32_t* nbr_ptr,
                                 const int32_t* nbr_idx, int32_t max_degree, void* out, int64_t ldo, int64_t nodes,
                                 int32_t heads, int32_t head_dim) {
  using namespace gc;
  (void)max_degree;
  GC_REQUIRE(qkv && nbr_ptr && nbr_idx && out, "gc_khop_attention: null buffer");
  GC_REQUIRE(dtype == GC_F32 || dtype == GC_BF16, "gc_khop_attention: dtype=%d", dtype);
  GC_REQUIRE(heads >= 1 && ld_qkv >= 3LL * heads * head_dim && ldo >= 1LL * heads * head_dim, "gc_khop_attention: bad sizes");
  GC_REQUIRE(aligned16(qkv) && ld_qkv % 8 == 0 && ldo % 8 == 0, "gc_khop_attention: alignment");
  if (nodes <= 0) return GC_OK;
  return launch_khop_attention_csr(reinterpret_cast<cudaStream_t>(stream), qkv, dtype, ld_qkv, nbr_ptr, nbr_idx, out, ldo,
                                   nodes, heads, head_dim);
}
