// TEST SCAFFOLDING -- a minimal stand-in for the part of XLA's FFI C++ API (xla/ffi/api/ffi.h, shipped inside jaxlib, which
// is not installable in this image) that bpl_next_b200/csrc/xla_ffi_shim.cc uses: enough to type-check the handlers and to
// call their implementations with device buffers from a test (tests/test_xla_ffi_shim.py).  It is NOT XLA: the binding
// builder only records nothing and the registration macro just exposes the implementation function.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { U8, U16, S32, F32 };
template <DataType T> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<U16> { using type = uint16_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<F32> { using type = float; };

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool failure() const { return code_ != ErrorCode::kOk; }
  const std::string& message() const { return msg_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T>
class Span {
 public:
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T& back() const { return p_[n_ - 1]; }

 private:
  const T* p_;
  size_t n_;
};

template <DataType T>
class Buffer {
 public:
  using Native = typename NativeOf<T>::type;
  Buffer(void* data, std::vector<int64_t> dims) : data_(static_cast<Native*>(data)), dims_(std::move(dims)) {}
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
  Native* typed_data() const { return data_; }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : dims_) n *= static_cast<size_t>(d);
    return n;
  }
  size_t size_bytes() const { return element_count() * sizeof(Native); }

 private:
  Native* data_;
  std::vector<int64_t> dims_;
};
template <DataType T>
class ResultBuffer {  // XLA's Result<Buffer<T>>: pointer-like
 public:
  explicit ResultBuffer(Buffer<T> b) : b_(std::move(b)) {}
  Buffer<T>* operator->() { return &b_; }

 private:
  Buffer<T> b_;
};

template <typename T> struct PlatformStream {};

struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an XLA_FFI_Handler symbol; here: evaluate the binding expression (type-check) and keep the impl
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  static auto name##_binding_check = (binding);            \
  auto* const name = &impl
