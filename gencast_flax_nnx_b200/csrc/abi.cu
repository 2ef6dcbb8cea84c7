// C-ABI plumbing: error reporting, introspection and the gc_gemm front end.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = []() {
    const char* v = getenv("GENCAST_PDL");
    return !(v != nullptr && v[0] == '0');
  }();
  return on;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return GC_ERR_CUDA;
}

}  // namespace gc

extern "C" {

const char* gc_last_error(void) { return gc::g_err; }

int gc_abi_version(void) { return 1; }

int gc_sizeof_gemm_args(void) { return static_cast<int>(sizeof(gc_gemm_args)); }

int gc_device_supports_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int gc_gemm(void* stream, const gc_gemm_args* a) {
  using namespace gc;
  GC_REQUIRE(a != nullptr, "gc_gemm: null args");
  GC_REQUIRE(a->num_segments >= 1 && a->num_segments <= GC_MAX_SEGMENTS, "gc_gemm: num_segments=%d", a->num_segments);
  GC_REQUIRE(a->m >= 0 && a->n > 0, "gc_gemm: bad m=%lld n=%d", (long long)a->m, a->n);
  GC_REQUIRE(a->n % 128 == 0, "gc_gemm: n=%d must be a multiple of 128", a->n);
  GC_REQUIRE(a->dtype == GC_F32 || a->dtype == GC_BF16, "gc_gemm: dtype=%d", a->dtype);
  const int64_t ld_align = a->dtype == GC_BF16 ? 8 : 4;
  for (int s = 0; s < a->num_segments; ++s) {
    GC_REQUIRE(a->a[s] && a->w[s], "gc_gemm: null operand in segment %d", s);
    GC_REQUIRE(a->k[s] > 0 && a->k[s] % 64 == 0, "gc_gemm: k[%d]=%d must be a positive multiple of 64", s, a->k[s]);
    GC_REQUIRE(a->lda[s] >= a->k[s] && a->ldw[s] >= a->k[s], "gc_gemm: leading dimension < k in segment %d", s);
    GC_REQUIRE(a->lda[s] % ld_align == 0 && a->ldw[s] % ld_align == 0, "gc_gemm: leading dimensions must be 16-byte multiples");
    GC_REQUIRE(aligned16(a->a[s]) && aligned16(a->w[s]), "gc_gemm: operands must be 16-byte aligned");
  }
  GC_REQUIRE(a->out != nullptr && aligned16(a->out) && a->ldo >= a->n && a->ldo % 8 == 0, "gc_gemm: bad output");
  GC_REQUIRE(a->out_dtype == GC_F32 || a->out_dtype == GC_BF16, "gc_gemm: out_dtype=%d", a->out_dtype);
  GC_REQUIRE(a->act >= GC_ACT_NONE && a->act <= GC_ACT_GELU_TANH, "gc_gemm: act=%d", a->act);
  if (a->addend) GC_REQUIRE(aligned16(a->addend) && a->ld_addend % 8 == 0, "gc_gemm: addend alignment");
  if (a->residual) GC_REQUIRE(aligned16(a->residual) && a->ld_res % 8 == 0, "gc_gemm: residual alignment");
  for (int j = 0; j < 2; ++j) {
    if (a->gather_src[j]) {
      GC_REQUIRE(a->gather_idx[j] != nullptr, "gc_gemm: gather_src[%d] without gather_idx", j);
      GC_REQUIRE(aligned16(a->gather_src[j]) && a->ld_gather[j] % 8 == 0, "gc_gemm: gather alignment");
    }
  }
  if (a->m == 0) return GC_OK;

  EpilogueParams ep;
  ep.bias = a->bias;
  ep.alpha_dev = a->alpha_dev;
  ep.addend = a->addend; ep.ld_addend = a->ld_addend; ep.addend_dtype = a->addend_dtype;
  ep.gsrc0 = a->gather_src[0]; ep.gidx0 = a->gather_idx[0]; ep.ldg0 = a->ld_gather[0];
  ep.gsrc1 = a->gather_src[1]; ep.gidx1 = a->gather_idx[1]; ep.ldg1 = a->ld_gather[1];
  ep.gather_dtype = a->gather_dtype;
  ep.act = a->act;
  ep.residual = a->residual; ep.ld_res = a->ld_res; ep.res_dtype = a->res_dtype;
  ep.out = a->out; ep.ldo = a->ldo; ep.out_dtype = a->out_dtype;
  ep.m = a->m; ep.n = a->n;

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == GC_BF16) return launch_gemm_tcgen05(st, *a, ep);
  return launch_gemm_ffma(st, *a, ep);
}

}  // extern "C"
