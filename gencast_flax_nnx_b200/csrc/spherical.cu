// Inverse real spherical-harmonic transform for the sampler's isotropic white noise
// (gencast/samplers_utils.py:250-346: coefficients drawn per (total wavenumber l, zonal wavenumber m), fields
// synthesised on the lat/lon grid by dinosaur's RealSphericalHarmonics.to_nodal):
//     field(lat, lon_j) = sum_m [ A_m(lat) cos(m phi_j) + B_m(lat) sin(m phi_j) ],   phi_j = 2 pi j / n_lon
//     A_m(lat) = sum_{l >= m} c_cos[m, l] Pbar[m, l, lat],   B_m likewise from c_sin
// Two kernels: the Legendre synthesis (a small batched contraction over l) and the longitude synthesis, which
// evaluates the trigonometric sum directly with an exact n_lon-periodic twiddle table (m j mod n_lon) and writes
// the sampler's state layout [member, lat * n_lon + lon, channel] -- no transposes, no FFT plan, no workspace.
// Set-up work of a 12 h step (once per 40 network evaluations): ~10 GFLOP at 1 deg x 4 members x 82 channels.
#include "common.cuh"

namespace gc {
namespace {

// spec[((member * n_lat + lat) * C + ch) * L + m] = (A_m, B_m) of field f = member * C + ch
__global__ void __launch_bounds__(256) sh_legendre_kernel(const float* __restrict__ coef, const float* __restrict__ table,
                                                          float2* __restrict__ spec, int L, int F, int C, int n_lat) {
  pdl_launch_dependents();
  pdl_wait();
  const int lat = blockIdx.x * 32 + (threadIdx.x & 31);
  const int f = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int m = blockIdx.z;
  if (lat >= n_lat || f >= F) return;
  const float* cc = coef + (static_cast<int64_t>(m) * F + f) * L;                 // [2][L(m)][F][L(l)]
  const float* cs = cc + static_cast<int64_t>(L) * F * L;
  const float* t = table + static_cast<int64_t>(m) * L * n_lat + lat;             // [L(m)][L(l)][n_lat]
  float a = 0.0f, b = 0.0f;
  for (int l = m; l < L; ++l) {
    const float p = __ldg(t + static_cast<int64_t>(l) * n_lat);
    a = fmaf(__ldg(cc + l), p, a);
    b = fmaf(__ldg(cs + l), p, b);
  }
  const int member = f / C, ch = f - member * C;
  spec[((static_cast<int64_t>(member) * n_lat + lat) * C + ch) * L + m] = make_float2(a, m == 0 ? 0.0f : b);
}

// One block per (member, lat): the row's C x L spectrum in shared memory, one thread per longitude.
constexpr int SH_CHUNK = 8;
__global__ void __launch_bounds__(512) sh_longitude_kernel(const float2* __restrict__ spec, float* __restrict__ out, int L, int C,
                                                           int n_lat, int n_lon) {
  extern __shared__ float2 sh[];                 // [C][L] spectrum | [n_lon] (cos, sin) twiddles
  float2* tw = sh + static_cast<int64_t>(C) * L;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;                // member * n_lat + lat
  const float2* src = spec + row * C * L;
  for (int i = threadIdx.x; i < C * L; i += blockDim.x) sh[i] = src[i];
  for (int i = threadIdx.x; i < n_lon; i += blockDim.x) {
    float s, c;
    sincospif(2.0f * static_cast<float>(i) / static_cast<float>(n_lon), &s, &c);
    tw[i] = make_float2(c, s);
  }
  __syncthreads();
  const int member = static_cast<int>(row / n_lat), lat = static_cast<int>(row - static_cast<int64_t>(member) * n_lat);
  for (int j = threadIdx.x; j < n_lon; j += blockDim.x) {
    float* dst = out + ((static_cast<int64_t>(member) * n_lat + lat) * n_lon + j) * C;
    for (int c0 = 0; c0 < C; c0 += SH_CHUNK) {
      float acc[SH_CHUNK];
#pragma unroll
      for (int k = 0; k < SH_CHUNK; ++k) acc[k] = 0.0f;
      int idx = 0;                               // (m * j) mod n_lon
      for (int m = 0; m < L; ++m) {
        const float2 w = tw[idx];
#pragma unroll
        for (int k = 0; k < SH_CHUNK; ++k) {
          if (c0 + k < C) {
            const float2 ab = sh[(c0 + k) * L + m];
            acc[k] = fmaf(ab.x, w.x, fmaf(ab.y, w.y, acc[k]));
          }
        }
        idx += j;
        if (idx >= n_lon) idx -= n_lon;
      }
#pragma unroll
      for (int k = 0; k < SH_CHUNK; ++k)
        if (c0 + k < C) dst[c0 + k] = acc[k];
    }
  }
}

}  // namespace
}  // namespace gc

extern "C" int gc_sh_synthesis(void* stream, const float* coef, const float* table, float* spec, float* out,
                               int32_t wavenumbers, int32_t members, int32_t channels, int32_t n_lat, int32_t n_lon) {
  using namespace gc;
  GC_REQUIRE(coef && table && spec && out, "gc_sh_synthesis: null buffer");
  GC_REQUIRE(wavenumbers >= 1 && members >= 1 && channels >= 1 && n_lat >= 1 && n_lon >= 2, "gc_sh_synthesis: bad sizes");
  GC_REQUIRE(2 * wavenumbers <= n_lon, "gc_sh_synthesis: %d zonal wavenumbers do not fit %d longitudes", wavenumbers, n_lon);
  const size_t smem = (static_cast<size_t>(channels) * wavenumbers + n_lon) * sizeof(float2);
  GC_REQUIRE(smem <= 227 * 1024, "gc_sh_synthesis: channels * wavenumbers = %d * %d does not fit shared memory", channels, wavenumbers);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int F = members * channels;
  GC_CHECK_CUDA(launch_kernel(sh_legendre_kernel, dim3((n_lat + 31) / 32, (F + 7) / 8, wavenumbers), dim3(256), 0, st, coef, table,
                              reinterpret_cast<float2*>(spec), (int)wavenumbers, F, (int)channels, (int)n_lat),
                "sh_legendre_kernel");
  GC_CHECK_CUDA(cudaFuncSetAttribute(sh_longitude_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                "cudaFuncSetAttribute(sh_longitude_kernel)");
  const int threads = n_lon >= 512 ? 512 : ((n_lon + 31) / 32) * 32;
  GC_CHECK_CUDA(launch_kernel(sh_longitude_kernel, dim3(members * n_lat), dim3(threads), smem, st,
                              reinterpret_cast<const float2*>(spec), out, (int)wavenumbers, (int)channels, (int)n_lat, (int)n_lon),
                "sh_longitude_kernel");
  return GC_OK;
}
