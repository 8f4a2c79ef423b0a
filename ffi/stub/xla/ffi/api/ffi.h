// STAND-IN for jaxlib's xla/ffi/api/ffi.h -- NOT the real header.  jax / jaxlib are not installable in the build
// image (no network), so ffi/qtx_ffi.cc is type-checked against this restatement of the small part of the public
// XLA FFI C++ API it uses (names and semantics as documented in the JAX FFI tutorial): Buffer<dtype>, Result<...>,
// PlatformStream, Error, Ffi::Bind().Ctx().Arg().Ret().Attr() and XLA_FFI_DEFINE_HANDLER_SYMBOL.  The macro here only
// checks that the handler is invocable with the decoded argument list; building against a real jaxlib
// (make -C ffi XLA_FFI_INCLUDE=$(python -c "from jax import ffi; print(ffi.include_dir())")) replaces it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum class DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16, C64, C128 };
inline constexpr DataType S8 = DataType::S8, S32 = DataType::S32, S64 = DataType::S64, U8 = DataType::U8,
                          U16 = DataType::U16, F32 = DataType::F32, F64 = DataType::F64, C128 = DataType::C128;

namespace internal {
template <DataType> struct NativeType;
template <> struct NativeType<DataType::S8> { using type = int8_t; };
template <> struct NativeType<DataType::S32> { using type = int32_t; };
template <> struct NativeType<DataType::S64> { using type = int64_t; };
template <> struct NativeType<DataType::U8> { using type = uint8_t; };
template <> struct NativeType<DataType::U16> { using type = uint16_t; };
template <> struct NativeType<DataType::F32> { using type = float; };
template <> struct NativeType<DataType::F64> { using type = double; };
}  // namespace internal

template <typename T>
class Span {
 public:
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T& back() const { return d_[n_ - 1]; }

 private:
  const T* d_;
  size_t n_;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename internal::NativeType<dtype>::type;
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_, rank_); }
  size_t element_count() const {
    size_t n = 1;
    for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i];
    return n;
  }
  size_t size_bytes() const { return element_count() * sizeof(T); }

 private:
  T* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename T> struct Decoded { using type = T; };
template <typename T> struct Decoded<PlatformStream<T>> { using type = T; };
}  // namespace internal

template <typename... Ts>
struct Binding {
  template <typename T> Binding<Ts..., typename internal::Decoded<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename F>
  static constexpr bool Accepts() { return std::is_invocable_r_v<Error, F, Ts...>; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                                                     \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>(), #impl " does not match its binding"); \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame*) { return nullptr; }
