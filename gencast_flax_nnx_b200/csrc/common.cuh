// Shared device/host helpers for the gencast_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gencast_b200.h"

namespace gc {

// ---------------------------------------------------------------------------
// Host-side error plumbing (thread-local message, errno-style return codes).
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define GC_REQUIRE(cond, ...)                      \
  do {                                             \
    if (!(cond)) {                                 \
      ::gc::set_error(__VA_ARGS__);                \
      return GC_ERR_INVALID_ARGUMENT;              \
    }                                              \
  } while (0)

#define GC_CHECK_LAUNCH(what)                                      \
  do {                                                             \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return ::gc::cuda_fail(_e, what);       \
  } while (0)

// ---------------------------------------------------------------------------
// Programmatic dependent launch: every kernel is launched with the
// programmatic-stream-serialization attribute, signals its dependents at entry and
// waits for its predecessor (griddepcontrol.wait) right before its first access to
// memory a previous kernel may have written or may still be reading.  The next
// kernel's launch latency, block scheduling and prologue (barrier init, TMEM
// allocation, descriptor prefetch, weight-side loads) overlap the current
// kernel's tail.  GENCAST_PDL=0 disables the attribute (the device-side
// instructions are then no-ops).
// ---------------------------------------------------------------------------
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define GC_CHECK_CUDA(expr, what)                                  \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::gc::cuda_fail(_e, what);       \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t dtype_size(int dt) { return dt == GC_BF16 ? 2 : 4; }

// ---------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float swish_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}

// One MUFU op; |error| ~ 2^-11, far below bf16 resolution: used by the bf16 paths only.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float swish_fast(float x) {
  // x * sigmoid(x), sigmoid(x) = 0.5 + 0.5 tanh(x / 2)
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(u), h);
}

template <bool FAST>
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == GC_ACT_SWISH) return FAST ? swish_fast(x) : swish_f(x);
  if (act == GC_ACT_GELU_TANH) return FAST ? gelu_tanh_fast(x) : gelu_tanh_f(x);
  return x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Load NV (multiple of 4 for f32, 8 for bf16 fast path) consecutive elements as float.
template <int NV>
__device__ __forceinline__ void load_as_float(const void* base, int dtype, int64_t elem_off, float (&t)[NV]) {
  if (dtype == GC_BF16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + elem_off;
    if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int i = 0; i < NV / 8; ++i) {
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(p) + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h[j]);
          t[i * 8 + 2 * j] = f.x;
          t[i * 8 + 2 * j + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        uint2 raw = __ldg(reinterpret_cast<const uint2*>(p) + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
        float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
        t[i * 4] = f0.x; t[i * 4 + 1] = f0.y; t[i * 4 + 2] = f1.x; t[i * 4 + 3] = f1.y;
      }
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      float4 f = __ldg(p + i);
      t[i * 4] = f.x; t[i * 4 + 1] = f.y; t[i * 4 + 2] = f.z; t[i * 4 + 3] = f.w;
    }
  }
}

template <int NV>
__device__ __forceinline__ void store_from_float(void* base, int dtype, int64_t elem_off, const float (&t)[NV]) {
  if (dtype == GC_BF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + elem_off;
    if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int i = 0; i < NV / 8; ++i) {
        uint4 raw;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(t[i * 8 + 2 * j], t[i * 8 + 2 * j + 1]);
        reinterpret_cast<uint4*>(p)[i] = raw;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        uint2 raw;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
        h[0] = __floats2bfloat162_rn(t[i * 4], t[i * 4 + 1]);
        h[1] = __floats2bfloat162_rn(t[i * 4 + 2], t[i * 4 + 3]);
        reinterpret_cast<uint2*>(p)[i] = raw;
      }
    }
  } else {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) p[i] = make_float4(t[i * 4], t[i * 4 + 1], t[i * 4 + 2], t[i * 4 + 3]);
  }
}

// ---------------------------------------------------------------------------
// The fused linear-layer epilogue, shared by the tcgen05 and the FFMA GEMMs.
// ---------------------------------------------------------------------------
struct EpilogueParams {
  const float* bias;
  const float* alpha_dev;
  const void* addend; int64_t ld_addend; int addend_dtype;
  const void* gsrc0; const int32_t* gidx0; int64_t ldg0;
  const void* gsrc1; const int32_t* gidx1; int64_t ldg1;
  int gather_dtype;
  int act;
  const void* residual; int64_t ld_res; int res_dtype;
  void* out; int64_t ldo; int out_dtype;
  int64_t m; int n;
};

// v[NV]: accumulators of row `row`, columns [col0, col0+NV).  alpha is preloaded.
// bias_ptr points at the NV bias values of these columns (global or shared memory) or is null.
// FAST selects the MUFU-based activations (bf16 paths).
template <int NV, bool FAST>
__device__ __forceinline__ void epilogue_row_segment(const EpilogueParams& p, float alpha, int64_t row, int col0,
                                                     float (&v)[NV], const float* bias_ptr) {
  float t[NV];
  if (p.alpha_dev != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= alpha;
  }
  if (bias_ptr != nullptr) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 b = reinterpret_cast<const float4*>(bias_ptr)[i];
      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (p.addend != nullptr) {
    load_as_float<NV>(p.addend, p.addend_dtype, row * p.ld_addend + col0, t);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += t[i];
  }
  if (p.gsrc0 != nullptr) {
    const int64_t r = __ldg(p.gidx0 + row);
    load_as_float<NV>(p.gsrc0, p.gather_dtype, r * p.ldg0 + col0, t);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += t[i];
  }
  if (p.gsrc1 != nullptr) {
    const int64_t r = __ldg(p.gidx1 + row);
    load_as_float<NV>(p.gsrc1, p.gather_dtype, r * p.ldg1 + col0, t);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += t[i];
  }
  if (p.act == GC_ACT_SWISH) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = apply_act<FAST>(v[i], GC_ACT_SWISH);
  } else if (p.act == GC_ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = apply_act<FAST>(v[i], GC_ACT_GELU_TANH);
  }
  if (p.residual != nullptr) {
    load_as_float<NV>(p.residual, p.res_dtype, row * p.ld_res + col0, t);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += t[i];
  }
  store_from_float<NV>(p.out, p.out_dtype, row * p.ldo + col0, v);
}

// Launchers implemented per translation unit.
int launch_gemm_tcgen05(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep);
int launch_gemm_ffma(cudaStream_t stream, const gc_gemm_args& a, const EpilogueParams& ep);

}  // namespace gc
