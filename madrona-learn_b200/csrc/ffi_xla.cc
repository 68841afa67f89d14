// XLA-FFI shim: wraps libmlb200 entry points as jax.ffi custom-call targets.
//
// NOT part of the default build: it needs jaxlib's header tree (xla/ffi/api/ffi.h), which is
// not present in this image (jax is not installable here), so this TU is compiled only by
// `make ffi JAX_INCLUDE=<dir>` on a machine that has jax.  The handlers are thin: they pull
// the CUDA stream and the device buffers out of the FFI call frame and forward to the C-ABI
// in include/mlb200.h, which is what the tests in this repository exercise through ctypes.
// See INTEGRATION.md section 1 for the Python side (jax.ffi.register_ffi_target / ffi_call).
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include "../../include/mlb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc, const char* what) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error::Internal(std::string(what) + " failed with code " + std::to_string(rc));
}

// advantages, returns = gae(rewards [T,N], values [T,N], dones u8 [T,N], bootstrap [N])
static ffi::Error GaeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> rewards,
                          ffi::Buffer<ffi::F32> values, ffi::Buffer<ffi::U8> dones,
                          ffi::Buffer<ffi::F32> bootstrap, ffi::ResultBuffer<ffi::F32> adv,
                          ffi::ResultBuffer<ffi::F32> ret, float gamma, float gamma_lambda) {
    const auto dims = rewards.dimensions();
    const int T = static_cast<int>(dims[0]);
    long long N = 1;
    for (size_t i = 1; i < dims.size(); ++i) N *= dims[i];
    return Status(mlb_gae_f32(stream, rewards.typed_data(), values.typed_data(), dones.typed_data(),
                              bootstrap.typed_data(), adv->typed_data(), ret->typed_data(), T, N,
                              gamma, gamma_lambda, nullptr, nullptr, nullptr, 0),
                  "mlb_gae_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_gae_f32_ffi, GaeImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::U8>>()
        .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Attr<float>("gamma").Attr<float>("gamma_lambda"),
    {ffi::Traits::kCmdBufferCompatible});

// out = (x - mean_rstd[0]) * mean_rstd[1]
static ffi::Error ZscoreApplyImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x,
                                  ffi::Buffer<ffi::F32> mean_rstd, ffi::ResultBuffer<ffi::F32> out) {
    return Status(mlb_zscore_apply_f32(stream, x.typed_data(), out->typed_data(),
                                       static_cast<long long>(x.element_count()),
                                       mean_rstd.typed_data()),
                  "mlb_zscore_apply_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_zscore_apply_f32_ffi, ZscoreApplyImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>(),
    {ffi::Traits::kCmdBufferCompatible});

// mb[s, m, :] = store[idx[m] / B, s, idx[m] % B, :]   (store [C, T', B, row])
static ffi::Error GatherImpl(cudaStream_t stream, ffi::AnyBuffer store, ffi::Buffer<ffi::S32> idx,
                             ffi::Result<ffi::AnyBuffer> out) {
    const auto d = store.dimensions();
    const int C = static_cast<int>(d[0]), Tp = static_cast<int>(d[1]);
    const long long B = d[2];
    long long row = static_cast<long long>(store.size_bytes()) / (static_cast<long long>(C) * Tp * B);
    return Status(mlb_mb_gather(stream, store.untyped_data(), idx.typed_data(), out->untyped_data(),
                                C, Tp, B, static_cast<long long>(idx.element_count()), row),
                  "mlb_mb_gather");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_mb_gather_ffi, GatherImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::AnyBuffer>().Arg<ffi::Buffer<ffi::S32>>().Ret<ffi::AnyBuffer>(),
    {ffi::Traits::kCmdBufferCompatible});
#endif  // __has_include("xla/ffi/api/ffi.h")
