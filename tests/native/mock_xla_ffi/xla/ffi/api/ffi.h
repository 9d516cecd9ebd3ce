// Minimal stand-in for jaxlib's `xla/ffi/api/ffi.h` (TEST INFRASTRUCTURE, not product): just enough of the public C++ FFI surface -
// Buffer / ResultBuffer / Span / Error / the binder chain / XLA_FFI_DEFINE_HANDLER_SYMBOL - for integration/xla_ffi/eincm_xla_ffi.cc to
// compile unchanged and for tests/native/ffi_harness.cpp to drive its handler bodies with device pointers.  jax / jaxlib are not
// installed in this image; with the real header the same source builds the handler XLA calls.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { S16, S32, S64, F32, F64 };

template <DataType dt> struct NativeType;
template <> struct NativeType<S16> { using type = int16_t; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<S64> { using type = int64_t; };
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<F64> { using type = double; };

template <typename T>
class Span {
 public:
    Span() = default;
    Span(const T* data, size_t size) : data_(data), size_(size) {}
    const T* begin() const { return data_; }
    const T* end() const { return data_ + size_; }
    size_t size() const { return size_; }
    const T& operator[](size_t i) const { return data_[i]; }
 private:
    const T* data_ = nullptr;
    size_t size_ = 0;
};

template <DataType dt>
class Buffer {
 public:
    using T = typename NativeType<dt>::type;
    Buffer() = default;
    Buffer(T* data, std::vector<int64_t> dims) : data_(data), dims_(std::move(dims)) {}
    T* typed_data() const { return data_; }
    void* untyped_data() const { return data_; }
    Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
    size_t element_count() const { size_t n = 1; for (int64_t d : dims_) n *= (size_t)d; return n; }
 private:
    T* data_ = nullptr;
    std::vector<int64_t> dims_;
};

template <typename T>
class Result {
 public:
    Result() = default;
    explicit Result(T v) : v_(std::move(v)) {}
    T* operator->() { return &v_; }
    T& operator*() { return v_; }
 private:
    T v_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

class Error {
 public:
    static Error Success() { return Error(); }
    static Error InvalidArgument(std::string m) { return Error(1, std::move(m)); }
    static Error Internal(std::string m) { return Error(2, std::move(m)); }
    bool success() const { return code_ == 0; }
    const std::string& message() const { return msg_; }
    int code() const { return code_; }
 private:
    Error() = default;
    Error(int c, std::string m) : code_(c), msg_(std::move(m)) {}
    int code_ = 0;
    std::string msg_;
};

template <typename T> struct PlatformStream {};
struct DeviceOrdinal {};

// binder chain: every step type-checks and returns the binder again
struct Binder {
    template <typename T> Binder& Ctx() { return *this; }
    template <typename T> Binder& Arg() { return *this; }
    template <typename T> Binder& Ret() { return *this; }
    template <typename T> Binder& Attr(const char*) { return *this; }
};
struct Ffi { static Binder Bind() { return Binder(); } };

}  // namespace ffi
}  // namespace xla

// the real macro defines the exported XLA_FFI_Handler symbol; the mock only keeps the operands alive for the type checker
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binder) \
    extern "C" int name##_mock_bound() { (void)(binder); return (&impl) != nullptr; }
