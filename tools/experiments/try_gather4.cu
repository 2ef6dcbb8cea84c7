// Experiment: semantics of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (which box shape the tensor map needs,
// where the four rows land in shared memory, how the 128B swizzle applies).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/try_gather4 tools/experiments/try_gather4.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap map, uint16_t* out, int* status, int c0, int r0, int r1, int r2, int r3,
                  int expect_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xdead;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(expect_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(&map), "r"(b), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
  }
  // bounded wait
  long long t0 = clock64();
  int ok = 0;
  while (clock64() - t0 < 20000000LL) {
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nmbarrier.test_wait.parity.shared::cta.b64 P, [%1], 0;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p) : "r"(b) : "memory");
    if (p) { ok = 1; break; }
  }
  if (threadIdx.x == 0) *status = ok;
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  const int R = 1000, C = 256;
  std::vector<uint16_t> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 64 + (c & 63) + ((c >> 6) << 14));
  uint16_t *d, *o; int* st;
  cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&o, 8192); cudaMalloc(&st, 4);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int boxrows : {1, 4}) for (int sw : {0, 1}) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)boxrows}; cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box rows %d swizzle %d: encode rc=%d\n", boxrows, sw, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    const int rows[4] = {5, 900, 17, 333};
    cudaMemset(st, 0, 4);
    k<<<1, 128, 8192 + 1024>>>(m, o, st, 64, rows[0], rows[1], rows[2], rows[3], 4 * 128);
    cudaError_t e = cudaDeviceSynchronize();
    int ok = 0; cudaMemcpy(&ok, st, 4, cudaMemcpyDeviceToHost);
    std::vector<uint16_t> r(4096); cudaMemcpy(r.data(), o, 8192, cudaMemcpyDeviceToHost);
    printf("  launch: %s, barrier completed: %d\n", cudaGetErrorString(e), ok);
    if (e != cudaSuccess) return 1;
    // where did row i, col 64 + j land?  expect value rows[i] * 64 + j + (1 << 14)
    for (int i = 0; i < 4; ++i) {
      int found_plain = 0, found_sw = 0;
      for (int j = 0; j < 64; ++j) {
        const uint16_t want = (uint16_t)(rows[i] * 64 + j + (1 << 14));
        if (r[i * 64 + j] == want) ++found_plain;
        const int unit = j >> 3, sunit = unit ^ (i & 7);
        if (r[i * 64 + sunit * 8 + (j & 7)] == want) ++found_sw;
      }
      printf("  row %d (%d): %d/64 at plain position, %d/64 at 128B-swizzled position\n", i, rows[i], found_plain, found_sw);
    }
    int touched = 0; for (int i = 0; i < 4096; ++i) touched += r[i] != 0xdead;
    printf("  smem elements written: %d (expected 256)\n", touched);
  }
  return 0;
}
