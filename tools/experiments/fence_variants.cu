// Experiment: what the spellings of the generic -> async proxy fence compile to on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -c tools/experiments/fence_variants.cu -o /tmp/f.o && cuobjdump -sass /tmp/f.o | grep -E "Function|MEMBAR|FENCE"
// Result (CUDA 12.9): all three are MEMBAR.ALL.{CTA,GPU} + FENCE.VIEW.ASYNC.S; the MEMBAR waits for every memory
// operation the thread has in flight, prefetched global loads included, so a thread that prefetches must not be the
// one that fences (edge_fused.cu, attention_gather.cu: the consumer warp fences after its barrier wait).
#include <cstdint>
__global__ void fence_shared_cta(float* o, const float* in) {
  extern __shared__ float s[];
  const float v = in[threadIdx.x];
  s[threadIdx.x] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  o[threadIdx.x] = s[threadIdx.x ^ 1] + v;
}
__global__ void fence_all(float* o, const float* in) {
  extern __shared__ float s[];
  const float v = in[threadIdx.x];
  s[threadIdx.x] = 1.f;
  asm volatile("fence.proxy.async;" ::: "memory");
  o[threadIdx.x] = s[threadIdx.x ^ 1] + v;
}
__global__ void fence_release_restrict(float* o, const float* in) {
  extern __shared__ float s[];
  const float v = in[threadIdx.x];
  s[threadIdx.x] = 1.f;
  asm volatile("fence.proxy.async::generic.release.sync_restrict::shared::cta.cluster;" ::: "memory");
  o[threadIdx.x] = s[threadIdx.x ^ 1] + v;
}
