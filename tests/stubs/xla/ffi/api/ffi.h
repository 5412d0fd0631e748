// TEST INFRASTRUCTURE: a signature-level stand-in for jaxlib's xla/ffi/api/ffi.h (absent from this image), so that
// physics-based-rl_b200/csrc/ffi_xla.cc is parsed and type-checked in CI: every handler's C++ signature must be invocable with
// exactly the parameter list its binding declares (context, attributes, operands, results, in binding order), which is the
// check XLA's real `Ffi::Bind()...To(fn)` performs.  Nothing here executes; the real header replaces it on a box with JAX.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

#define PBRL_XLA_FFI_STUB 1

namespace xla {
namespace ffi {

enum class DataType { F32, S32 };
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType S32 = DataType::S32;
template <DataType> struct NativeType;
template <> struct NativeType<DataType::F32> { using type = float; };
template <> struct NativeType<DataType::S32> { using type = int32_t; };

template <typename T>
class Span {
 public:
    const T& operator[](size_t i) const { return data_[i]; }
    size_t size() const { return size_; }
 private:
    const T* data_ = nullptr;
    size_t size_ = 0;
};

template <DataType dtype>
class Buffer {
 public:
    using Native = typename NativeType<dtype>::type;
    Native* typed_data() const { return nullptr; }
    Span<int64_t> dimensions() const { return {}; }
    size_t element_count() const { return 0; }
    size_t size_bytes() const { return 0; }
};

template <typename B>
class Result {
 public:
    B* operator->() { return &b_; }
    B& operator*() { return b_; }
 private:
    B b_;
};

enum class ErrorCode { kInternal, kInvalidArgument };
class Error {
 public:
    Error() = default;
    Error(ErrorCode, std::string) {}
    static Error Success() { return Error(); }
};

template <typename T> struct PlatformStream {};
template <typename C> struct CtxParam;
template <typename T> struct CtxParam<PlatformStream<T>> { using type = T; };

template <typename... Ts>
struct Binding {
    template <typename C> Binding<Ts..., typename CtxParam<C>::type> Ctx() const { return {}; }
    template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
    template <typename A> Binding<Ts..., A> Arg() const { return {}; }
    template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
    template <typename Fn> static constexpr bool kMatches = std::is_invocable_r_v<Error, Fn, Ts...>;
    static constexpr size_t kArity = sizeof...(Ts);
};
struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                          \
    static_assert(std::remove_reference_t<decltype(binding)>::template kMatches<decltype(&impl)>,                     \
                  #symbol ": the handler's parameters do not match its binding (context, attributes, operands, results)"); \
    extern "C" void* symbol(void*) { return nullptr; }
