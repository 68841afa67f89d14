// MOCK of the slice of jaxlib's xla/ffi/api/ffi.h that madrona-learn_b200/csrc/ffi_xla.cc uses.
//
// Test infrastructure only (tests/test_abi.py::test_ffi_shim_type_checks): jax / jaxlib are not installable in
// this image, so the real header does not exist here.  This stand-in lets g++ TYPE-CHECK the shim: every handler
// body (its calls into include/mlb200.h are checked against the C prototypes) and every binding (the binder
// accumulates the C++ parameter types of .Ctx / .Arg / .Ret / .Attr in order and .To() static_asserts that the
// implementation function is invocable with exactly those).  It mirrors names and call shapes of the public
// XLA FFI C++ API (xla::ffi::Ffi::Bind, Buffer<dtype>, AnyBuffer, Result<>, Error, Span, PlatformStream,
// Traits, XLA_FFI_DEFINE_HANDLER_SYMBOL); it does not implement the call frame and cannot run anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { U8, S32, U32, F32, F64, BF16 };
template <DataType D> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<BF16> { using type = uint16_t; };

template <typename T>
struct Span {
    const T* ptr = nullptr;
    size_t n = 0;
    const T* begin() const { return ptr; }
    const T* data() const { return ptr; }
    size_t size() const { return n; }
    const T& operator[](size_t i) const { return ptr[i]; }
};

class Error {
  public:
    static Error Success() { return Error(); }
    static Error Internal(std::string m) { Error e; e.msg_ = std::move(m); e.ok_ = false; return e; }
    static Error InvalidArgument(std::string m) { return Internal(std::move(m)); }
    bool success() const { return ok_; }
  private:
    bool ok_ = true;
    std::string msg_;
};

class AnyBuffer {
  public:
    void* untyped_data() const { return data_; }
    Span<const int64_t> dimensions() const { return dims_; }
    size_t size_bytes() const { return bytes_; }
    size_t element_count() const { return count_; }
  private:
    void* data_ = nullptr;
    Span<const int64_t> dims_;
    size_t bytes_ = 0, count_ = 0;
};

template <DataType D>
class Buffer {
  public:
    using T = typename NativeOf<D>::type;
    T* typed_data() const { return data_; }
    void* untyped_data() const { return data_; }
    Span<const int64_t> dimensions() const { return dims_; }
    size_t size_bytes() const { return count_ * sizeof(T); }
    size_t element_count() const { return count_; }
  private:
    T* data_ = nullptr;
    Span<const int64_t> dims_;
    size_t count_ = 0;
};

template <typename B>
class Result {
  public:
    B* operator->() { return &b_; }
    B& operator*() { return b_; }
  private:
    B b_;
};
template <DataType D> using ResultBuffer = Result<Buffer<D>>;

template <typename S> struct PlatformStream { using type = S; };

enum class Traits : uint32_t { kCmdBufferCompatible = 1 };

namespace detail {
template <typename T> struct CtxParam { using type = T; };
template <typename S> struct CtxParam<PlatformStream<S>> { using type = S; };
}  // namespace detail

template <typename... Ps>
class Binding {
  public:
    template <typename T> Binding<Ps..., typename detail::CtxParam<T>::type> Ctx() const { return {}; }
    template <typename T> Binding<Ps..., T> Arg() const { return {}; }
    template <typename T> Binding<Ps..., Result<T>> Ret() const { return {}; }
    template <typename T> Binding<Ps..., T> Attr(const char*) const { return {}; }
    template <typename Fn>
    int To(Fn&&) const {
        static_assert(std::is_invocable_r<Error, Fn, Ps...>::value,
                      "XLA-FFI binding and handler implementation disagree on the parameter list");
        return 0;
    }
};

struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an XLA_FFI_Handler symbol; the mock only instantiates the binding's type check
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding, ...) \
    extern "C" int sym##_mock_check() { return (binding).To(impl); }
