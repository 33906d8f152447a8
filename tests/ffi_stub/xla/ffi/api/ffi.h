// TEST INFRASTRUCTURE - a minimal restatement of the public XLA typed-FFI C++ API (xla/ffi/api/ffi.h as documented in the JAX
// FFI tutorial), just enough to TYPE-CHECK gpfy_b200/csrc/xla_ffi_shim.cc in an image without jaxlib: buffer accessors,
// Span, Error / ErrorCode, PlatformStream, and a binding builder that records the decoded argument types so that
// XLA_FFI_DEFINE_HANDLER_SYMBOL can static_assert that the handler is callable with exactly those types.
// It executes nothing and is never linked into the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum DataType { U8, S32, S64, F32, F64 };
template <DataType> struct NativeTypeOf;
template <> struct NativeTypeOf<U8> { using type = uint8_t; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };
template <> struct NativeTypeOf<S64> { using type = int64_t; };
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<F64> { using type = double; };

template <class T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
  const T* begin() const { return ptr; }
  const T* end() const { return ptr + n; }
};

template <DataType dt>
struct Buffer {
  using T = typename NativeTypeOf<dt>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  Span<const int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <class T>
struct Result {
  T value;
  T* operator->() { return &value; }
  T& operator*() { return value; }
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kOk, kInvalidArgument, kUnimplemented, kResourceExhausted, kInternal };
struct Error {
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(ErrorCode::kOk, ""); }
};

template <class T> struct PlatformStream {};

namespace stub {
template <class T> struct Decoded { using type = T; };
template <class T> struct Decoded<PlatformStream<T>> { using type = T; };
template <class... Ts> struct Binding {
  template <class T> Binding<Ts..., typename Decoded<T>::type> Ctx() const { return {}; }
  template <class T> Binding<Ts..., T> Arg() const { return {}; }
  template <class T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <class T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <class Fn> static constexpr bool Matches() { return std::is_invocable_r_v<Error, Fn, Ts...>; }
};
}  // namespace stub

struct Ffi {
  static stub::Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding)                                                              \
  static_assert(decltype(binding)::template Matches<decltype(&fn)>(), "handler signature does not match its binding"); \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*) { return nullptr; }
