// MOCK of the slice of xla/ffi/api/ffi.h that csrc/lob_ffi.cc uses (TEST INFRASTRUCTURE).  The real header ships with
// jaxlib (jax.ffi.include_dir()), which is not installable in this image; this stand-in has the same names and call
// shapes (xla::ffi::Span, AnyBuffer, Result<AnyBuffer>, RemainingArgs / RemainingRets with get<T>(i) -> ErrorOr<T>,
// Error, Ffi::Bind().Ctx<>().Attr<>().RemainingArgs().RemainingRets(), XLA_FFI_DEFINE_HANDLER_SYMBOL), so that
// tests/test_ffi_tables.py can at least compile the handler source and call it with fake buffers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace xla::ffi {

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  const T* data() const { return d_; }
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
 private:
  const T* d_ = nullptr;
  size_t n_ = 0;
};

class Error {
 public:
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { Error e; e.fail_ = true; e.msg_ = std::move(m); return e; }
  bool failure() const { return fail_; }
  bool success() const { return !fail_; }
  const std::string& message() const { return msg_; }
 private:
  bool fail_ = false;
  std::string msg_;
};

class AnyBuffer {
 public:
  explicit AnyBuffer(void* p = nullptr) : p_(p) {}
  void* untyped_data() const { return p_; }
 private:
  void* p_;
};
template <typename T>
class Result {
 public:
  explicit Result(T v) : v_(v) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }
 private:
  T v_;
};
template <typename T>
using ErrorOr = std::optional<T>;

class RemainingArgs {
 public:
  std::vector<void*> ptrs;
  size_t size() const { return ptrs.size(); }
  template <typename T> ErrorOr<T> get(size_t i) const { return i < ptrs.size() ? ErrorOr<T>(T(ptrs[i])) : std::nullopt; }
};
class RemainingRets {
 public:
  std::vector<void*> ptrs;
  size_t size() const { return ptrs.size(); }
  template <typename T> ErrorOr<Result<T>> get(size_t i) const {
    return i < ptrs.size() ? ErrorOr<Result<T>>(Result<T>(T(ptrs[i]))) : std::nullopt;
  }
};

template <typename T> struct PlatformStream {};

// the binding DSL: only its shape matters here
struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
  Binding& RemainingArgs() { return *this; }
  Binding& RemainingRets() { return *this; }
};
struct Ffi { static Binding Bind() { return Binding(); } };

}  // namespace xla::ffi

// the real macro defines an exported XLA_FFI_Handler; the mock exports the implementation's address under that name
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  extern "C" void* name() { (void)(binding); return reinterpret_cast<void*>(&impl); }
