// Minimal stand-in for XLA's xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE, not the product and
// not XLA.  It models just enough of the typed-FFI surface (Buffer / ResultBuffer / Span /
// Error / PlatformStream / the Ffi::Bind() builder / XLA_FFI_DEFINE_HANDLER_SYMBOL) for
// jax_sgmc_b200/csrc/ffi_shim.cc to compile with g++ on a box without jax, so that
//   * every handler's parameter list is type-checked against its binding (Ctx, Arg, Attr,
//     Ret in declaration order, as XLA decodes a call frame), and
//   * every forwarded call is type-checked against the prototypes of include/sgmc_b200.h.
// The handlers it defines are inert (they never decode a call frame).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum DataType { U8, S32, S64, U32, F32 };
template <DataType> struct NativeType;
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<S64> { using type = int64_t; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<F32> { using type = float; };

template <typename T>
class Span {
 public:
  Span() = default;
  Span(T* b, size_t n) : b_(b), n_(n) {}
  T* begin() const { return b_; }
  T* end() const { return b_ + n_; }
  size_t size() const { return n_; }
  T& operator[](size_t i) const { return b_[i]; }
 private:
  T* b_ = nullptr;
  size_t n_ = 0;
};

enum class ErrorCode { kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : ok_(false), message_(std::move(message)) { (void)code; }
  static Error Success() { return Error(); }
  bool success() const { return ok_; }
 private:
  bool ok_ = true;
  std::string message_;
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeType<dt>::type;
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<const int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return count_; }
  size_t size_bytes() const { return count_ * sizeof(T); }
 private:
  T* data_ = nullptr;
  Span<const int64_t> dims_;
  size_t count_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }
 private:
  T value_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

template <typename T> struct PlatformStream {};

namespace internal {
template <typename T> struct CtxParam;
template <typename T> struct CtxParam<PlatformStream<T>> { using type = T; };
}  // namespace internal

// Bind().Ctx<..>().Arg<..>().Attr<..>("name").Ret<..>() records, in order, the parameter
// types the handler implementation must take.
template <typename... Params>
struct Binding {
  template <typename T> Binding<Params..., typename internal::CtxParam<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Params..., T> Arg() const { return {}; }
  template <typename T> Binding<Params..., T> Attr(const char*) const { return {}; }
  template <typename T> Binding<Params..., Result<T>> Ret() const { return {}; }
  template <typename Fn> struct Matches : std::false_type {};
  template <typename... Args>
  struct Matches<Error (*)(Args...)> : std::is_same<std::tuple<Args...>, std::tuple<Params...>> {};
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                   \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>::value,                   \
                "handler " #symbol ": parameters of " #impl " do not match its binding");      \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame* frame) {                                 \
    (void)frame;                                                                               \
    (void)&impl;                                                                               \
    return nullptr;                                                                            \
  }
