// Minimal stand-in for the part of the XLA FFI C++ API (xla/ffi/api/ffi.h, shipped inside jaxlib) that
// gencast_flax_nnx_b200/csrc/xla_ffi_shim.cc uses.  jaxlib cannot be installed in the build image, so the shim cannot be
// compiled against the real header here; this stub lets the compiler at least type-check every handler body - in particular
// every call into the C ABI of include/gencast_b200.h - so that the shim cannot silently drift from the library
// (tests/test_abi.py runs `g++ -fsyntax-only` with this directory on the include path).  Not used at run time.
#pragma once
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>
#include <vector>

namespace xla {
namespace ffi {

enum class DataType { F32, BF16, S32, U32 };
constexpr DataType F32 = DataType::F32;
constexpr DataType BF16 = DataType::BF16;
constexpr DataType S32 = DataType::S32;
constexpr DataType U32 = DataType::U32;

enum class ErrorCode { kInvalidArgument, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

struct Span {
  int64_t operator[](size_t) const { return 0; }
  size_t size() const { return 0; }
};

template <DataType> struct NativeOf { using type = float; };
template <> struct NativeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeOf<DataType::U32> { using type = uint32_t; };
template <> struct NativeOf<DataType::BF16> { using type = uint16_t; };

class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  Span dimensions() const { return {}; }
  DataType element_type() const { return DataType::F32; }
  size_t element_count() const { return 0; }
};

template <DataType T>
class Buffer {
 public:
  using Native = typename NativeOf<T>::type;
  Native* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  Span dimensions() const { return {}; }
  DataType element_type() const { return T; }
  size_t element_count() const { return 0; }
};

template <typename B>
class Result {
 public:
  B* operator->() { return &b_; }
  B& operator*() { return b_; }
 private:
  B b_;
};

template <typename S> struct PlatformStream {};

enum class Traits { kCmdBufferCompatible };

struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
};

struct Ffi {
  static Binding Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// The real macro defines an exported XLA_FFI_Error* symbol(XLA_FFI_CallFrame*); here it only has to consume its arguments
// and reference the implementation, so that unused-function / signature mistakes still surface.
#define XLA_FFI_STUB_CAT2(a, b) a##b
#define XLA_FFI_STUB_CAT(a, b) XLA_FFI_STUB_CAT2(a, b)
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding, ...)                    \
  extern "C" const void* name() {                                                  \
    (void)(binding);                                                               \
    return reinterpret_cast<const void*>(&impl);                                   \
  }
