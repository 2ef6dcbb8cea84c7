// XLA FFI (jax.ffi custom-call) handlers over the plain C ABI of include/gencast_b200.h.
//
// Not part of the default build: the XLA FFI headers ship inside jaxlib
// (jax.ffi.include_dir()), which is not installable in the build image.  Where jaxlib is
// present, compile this file against that include directory and link it with
// libgencast_b200.so; the handlers only translate FFI buffers into the
// pointer/size arguments of the launchers and forward the caller's stream.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"
#include "../../include/gencast_b200.h"

namespace ffi = xla::ffi;

static int dtype_code(ffi::DataType t) { return t == ffi::DataType::BF16 ? GC_BF16 : GC_F32; }

static ffi::Error status(int rc) {
  if (rc == GC_OK) return ffi::Error::Success();
  return ffi::Error(rc == GC_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    gc_last_error());
}

// y = act(a @ w^T + bias) (+ residual): nnx.Linear inside MLP / transformer projections.
static ffi::Error LinearImpl(cudaStream_t stream, ffi::AnyBuffer a, ffi::AnyBuffer w, ffi::Buffer<ffi::F32> bias,
                             ffi::Result<ffi::AnyBuffer> out, int32_t act) {
  gc_gemm_args g = {};
  g.a[0] = a.untyped_data(); g.w[0] = w.untyped_data();
  g.lda[0] = a.dimensions()[1]; g.ldw[0] = w.dimensions()[1]; g.k[0] = (int32_t)a.dimensions()[1];
  g.num_segments = 1; g.m = a.dimensions()[0]; g.n = (int32_t)w.dimensions()[0];
  g.dtype = dtype_code(a.element_type());
  g.bias = bias.typed_data(); g.act = act;
  g.out = out->untyped_data(); g.ldo = g.n; g.out_dtype = dtype_code(out->element_type());
  return status(gc_gemm(stream, &g));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gc_linear_ffi, LinearImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::AnyBuffer>().Attr<int32_t>("act"),
                              {ffi::Traits::kCmdBufferCompatible});

// LayerNorm + conditional affine (+ residual): common/mlp.py:59-65,121-145.
static ffi::Error LnCondImpl(cudaStream_t stream, ffi::AnyBuffer x, ffi::Buffer<ffi::F32> scale_offset,
                             ffi::Result<ffi::AnyBuffer> out) {
  const int64_t rows = x.dimensions()[0];
  const int32_t cols = (int32_t)x.dimensions()[1];
  return status(gc_ln_cond(stream, x.untyped_data(), dtype_code(x.element_type()), cols, scale_offset.typed_data(), 1,
                           nullptr, 0, 0, out->untyped_data(), dtype_code(out->element_type()), cols, rows, cols));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gc_ln_cond_ffi, LnCondImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::AnyBuffer>(),
                              {ffi::Traits::kCmdBufferCompatible});

// jraph.segment_sum of LN+cond'ed edge rows, receiver-sorted CSR.
static ffi::Error SegmentSumImpl(cudaStream_t stream, ffi::AnyBuffer y, ffi::Buffer<ffi::F32> scale_offset,
                                 ffi::Buffer<ffi::S32> row_ptr, ffi::Buffer<ffi::S32> edge_perm,
                                 ffi::Result<ffi::AnyBuffer> out) {
  const int32_t cols = (int32_t)y.dimensions()[1];
  return status(gc_ln_cond_segment_sum(stream, y.untyped_data(), dtype_code(y.element_type()), cols,
                                       scale_offset.typed_data(), 1, row_ptr.typed_data(), edge_perm.typed_data(),
                                       out->untyped_data(), dtype_code(out->element_type()), cols,
                                       out->dimensions()[0], cols));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gc_segment_sum_ffi, SegmentSumImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Ret<ffi::AnyBuffer>(),
                              {ffi::Traits::kCmdBufferCompatible});

// TriblockdiagMHA without projections, on the block-sparse tile list.
static ffi::Error AttentionImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> qkv, ffi::Buffer<ffi::S32> tile_ptr,
                                ffi::Buffer<ffi::S32> tile_kv, ffi::Buffer<ffi::U32> tile_mask,
                                ffi::Result<ffi::Buffer<ffi::BF16>> out, int32_t heads, int32_t head_dim) {
  return status(gc_khop_attention_tiles(stream, qkv.typed_data(), qkv.dimensions()[1], tile_ptr.typed_data(),
                                        tile_kv.typed_data(), tile_mask.typed_data(), out->typed_data(),
                                        out->dimensions()[1], qkv.dimensions()[0], heads, head_dim));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gc_khop_attention_ffi, AttentionImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>().Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("heads").Attr<int32_t>("head_dim"),
                              {ffi::Traits::kCmdBufferCompatible});
#endif  // __has_include("xla/ffi/api/ffi.h")
