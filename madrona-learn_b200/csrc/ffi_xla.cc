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

#include <initializer_list>
#include <string>

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

// ---------------------------------------------------------------------------------------------------------
// The learner path proper.  Conventions: every output (and every scratch buffer the kernel needs) is an FFI
// result the caller declares with jax.ShapeDtypeStruct; buffers the C-ABI updates IN PLACE (parameters, Adam
// moments, PRNG keys, the gradient arena) are passed as operand AND result with
// ffi_call(..., input_output_aliases={i: j}) -- the handler refuses un-aliased pairs; small host-side tables
// (bucket counts, per-component scales) are static attributes (Span).
// ---------------------------------------------------------------------------------------------------------
static long long RowsOf(ffi::Span<const int64_t> d) {          // product of all but the last dimension
    long long n = 1;
    for (size_t i = 0; i + 1 < d.size(); ++i) n *= d[i];
    return n;
}
static ffi::Error Aliased(const void* in, const void* out, const char* what) {
    if (in == out) return ffi::Error::Success();
    return ffi::Error::InvalidArgument(std::string(what) + ": operand and result must be aliased (input_output_aliases)");
}

// returns = compute_returns(rewards [T,N], dones u8 [T,N], bootstrap [N])      (ml/algo_common.py:45-82)
static ffi::Error ReturnsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> rewards, ffi::Buffer<ffi::U8> dones,
                              ffi::Buffer<ffi::F32> bootstrap, ffi::ResultBuffer<ffi::F32> ret, float gamma) {
    const auto dims = rewards.dimensions();
    const int T = static_cast<int>(dims[0]);
    const long long N = static_cast<long long>(rewards.element_count()) / (T > 0 ? T : 1);
    return Status(mlb_returns_f32(stream, rewards.typed_data(), dones.typed_data(), bootstrap.typed_data(),
                                  ret->typed_data(), T, N, gamma), "mlb_returns_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_returns_f32_ffi, ReturnsImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Attr<float>("gamma"),
    {ffi::Traits::kCmdBufferCompatible});

// key', perm [E, J] = E x (split, permutation(arange(J)))                      (ml/ppo.py:445-458)
static ffi::Error PermutationsImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> key, ffi::ResultBuffer<ffi::U32> key_out,
                                   ffi::ResultBuffer<ffi::S32> perm, ffi::ResultBuffer<ffi::U8> ws,
                                   int32_t partitionable) {
    if (auto e = Aliased(key.typed_data(), key_out->typed_data(), "update key"); !e.success()) return e;
    const auto d = perm->dimensions();
    return Status(mlb_ppo_permutations(stream, key_out->typed_data(), perm->typed_data(), static_cast<int>(d[0]),
                                       static_cast<long long>(d[1]), partitionable, ws->typed_data(),
                                       ws->size_bytes()), "mlb_ppo_permutations");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_ppo_permutations_ffi, PermutationsImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::U32>>().Ret<ffi::Buffer<ffi::U32>>().Ret<ffi::Buffer<ffi::S32>>()
        .Ret<ffi::Buffer<ffi::U8>>().Attr<int32_t>("partitionable"),
    {ffi::Traits::kCmdBufferCompatible});

// rollout store of one step: reward / done slabs + discounted env-return trace  (ml/rollouts.py:946-978)
static ffi::Error PostStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> rewards, ffi::Buffer<ffi::U8> dones,
                               ffi::Buffer<ffi::F32> env_returns, ffi::ResultBuffer<ffi::F32> reward_slab,
                               ffi::ResultBuffer<ffi::U8> done_slab, ffi::ResultBuffer<ffi::F32> env_returns_out,
                               ffi::ResultBuffer<ffi::F32> trace, float gamma) {
    if (auto e = Aliased(env_returns.typed_data(), env_returns_out->typed_data(), "env_returns"); !e.success()) return e;
    return Status(mlb_post_step_store_f32(stream, rewards.typed_data(), dones.typed_data(), reward_slab->typed_data(),
                                          done_slab->typed_data(), env_returns_out->typed_data(), trace->typed_data(),
                                          static_cast<long long>(rewards.element_count()), gamma),
                  "mlb_post_step_store_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_post_step_store_f32_ffi, PostStepImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Attr<float>("gamma"),
    {ffi::Traits::kCmdBufferCompatible});

// actions, log_probs, values = DiscreteActionDistributions.sample / best on a head [rows, ld]   (ml/dists.py:26-58)
static ffi::Error SampleDiscreteImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> head, ffi::Buffer<ffi::U32> policy_key,
                                     ffi::ResultBuffer<ffi::S32> actions, ffi::ResultBuffer<ffi::F32> log_probs,
                                     ffi::ResultBuffer<ffi::F32> values, ffi::Span<const int32_t> buckets,
                                     int32_t partitionable, int32_t deterministic) {
    const auto d = head.dimensions();
    const int ld = static_cast<int>(d[d.size() - 1]);
    return Status(mlb_sample_discrete_f32(stream, head.typed_data(), ld, policy_key.typed_data(), buckets.data(),
                                          static_cast<int>(buckets.size()), RowsOf(d), partitionable, deterministic,
                                          actions->typed_data(), log_probs->typed_data(), values->typed_data(),
                                          nullptr, 1), "mlb_sample_discrete_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_sample_discrete_f32_ffi, SampleDiscreteImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::U32>>()
        .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Attr<ffi::Span<const int32_t>>("buckets").Attr<int32_t>("partitionable").Attr<int32_t>("deterministic"),
    {ffi::Traits::kCmdBufferCompatible});

// C = op(A) op(B) on tcgen05 kind::tf32 -- an f32 dot_general at XLA's default precision      (ml/models.py:110-154)
static ffi::Error GemmTf32Impl(cudaStream_t stream, ffi::Buffer<ffi::F32> A, ffi::Buffer<ffi::F32> B,
                               ffi::ResultBuffer<ffi::F32> C, int32_t trans_a, int32_t trans_b) {
    const auto da = A.dimensions(), db = B.dimensions(), dc = C->dimensions();
    const int M = static_cast<int>(dc[0]), N = static_cast<int>(dc[1]);
    const int K = static_cast<int>(trans_a ? da[0] : da[1]);
    return Status(mlb_gemm_tf32_tc(stream, A.typed_data(), B.typed_data(), C->typed_data(), nullptr, M, N, K,
                                   static_cast<int>(da[1]), static_cast<int>(db[1]), N, trans_a, trans_b, 0, 1),
                  "mlb_gemm_tf32_tc");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_gemm_tf32_tc_ffi, GemmTf32Impl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Attr<int32_t>("trans_a").Attr<int32_t>("trans_b"),
    {ffi::Traits::kCmdBufferCompatible});

// z, y, stats = one Dense -> LayerNorm -> ReLU layer, compute_dtype=float32                    (ml/models.py:107-117)
static ffi::Error DenseLnReluTf32Impl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> w,
                                      ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> bias,
                                      ffi::ResultBuffer<ffi::F32> z, ffi::ResultBuffer<ffi::F32> y,
                                      ffi::ResultBuffer<ffi::F32> stats) {
    const auto dx = x.dimensions(), dw = w.dimensions();
    const int K = static_cast<int>(dw[0]), H = static_cast<int>(dw[1]);
    return Status(mlb_dense_ln_relu_fwd_tf32(stream, x.typed_data(), w.typed_data(), scale.typed_data(),
                                             bias.typed_data(), z->typed_data(), y->typed_data(), stats->typed_data(),
                                             RowsOf(dx), K, H, K), "mlb_dense_ln_relu_fwd_tf32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_dense_ln_relu_fwd_tf32_ffi, DenseLnReluTf32Impl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>(),
    {ffi::Traits::kCmdBufferCompatible});

// dz, dscale, dbias = vjp of LayerNorm + ReLU (the bwd of the custom_vjp around the layer)
static ffi::Error LnReluBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> dy, ffi::Buffer<ffi::F32> z,
                                ffi::Buffer<ffi::F32> stats, ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> bias,
                                ffi::ResultBuffer<ffi::F32> dz, ffi::ResultBuffer<ffi::F32> dscale,
                                ffi::ResultBuffer<ffi::F32> dbias) {
    const auto d = z.dimensions();
    const int H = static_cast<int>(d[d.size() - 1]);
    if (cudaMemsetAsync(dscale->typed_data(), 0, dscale->size_bytes(), stream) != cudaSuccess ||
        cudaMemsetAsync(dbias->typed_data(), 0, dbias->size_bytes(), stream) != cudaSuccess)
        return ffi::Error::Internal("cudaMemsetAsync failed");
    return Status(mlb_ln_relu_bwd_f32(stream, dy.typed_data(), z.typed_data(), stats.typed_data(), scale.typed_data(),
                                      bias.typed_data(), dz->typed_data(), dscale->typed_data(), dbias->typed_data(),
                                      RowsOf(d), H), "mlb_ln_relu_bwd_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_ln_relu_bwd_f32_ffi, LnReluBwdImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>(),
    {ffi::Traits::kCmdBufferCompatible});

// y, xhat, rstd = one Dense -> LayerNorm -> ReLU layer, compute_dtype=bfloat16 (X [M,K], Wt = W^T [H,K], bf16)
static ffi::Error DenseLnReluTcImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> X, ffi::Buffer<ffi::BF16> Wt,
                                    ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> bias,
                                    ffi::ResultBuffer<ffi::BF16> Y, ffi::ResultBuffer<ffi::BF16> XH,
                                    ffi::ResultBuffer<ffi::F32> rstd) {
    const auto dx = X.dimensions(), dw = Wt.dimensions();
    const int K = static_cast<int>(dw[1]), H = static_cast<int>(dw[0]);
    return Status(mlb_dense_ln_relu_fwd_tc(stream, X.typed_data(), Wt.typed_data(), scale.typed_data(),
                                           bias.typed_data(), Y->typed_data(), XH->typed_data(), rstd->typed_data(),
                                           static_cast<int>(RowsOf(dx)), K, H, K, K), "mlb_dense_ln_relu_fwd_tc");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_dense_ln_relu_fwd_tc_ffi, DenseLnReluTcImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::BF16>>().Ret<ffi::Buffer<ffi::BF16>>()
        .Ret<ffi::Buffer<ffi::F32>>(),
    {ffi::Traits::kCmdBufferCompatible});

// dz_prev, dscale, dbias = LN'/ReLU'(dz W^T) of the previous layer (bf16 path; W [H_prev, K] is this layer's kernel)
static ffi::Error DenseDxLnBwdTcImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> DZ, ffi::Buffer<ffi::BF16> W,
                                     ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> bias,
                                     ffi::Buffer<ffi::BF16> XH, ffi::Buffer<ffi::F32> rstd,
                                     ffi::ResultBuffer<ffi::BF16> DZ_out, ffi::ResultBuffer<ffi::F32> dscale,
                                     ffi::ResultBuffer<ffi::F32> dbias) {
    const auto dz = DZ.dimensions(), dw = W.dimensions();
    const int K = static_cast<int>(dw[1]), HN = static_cast<int>(dw[0]);
    if (cudaMemsetAsync(dscale->typed_data(), 0, dscale->size_bytes(), stream) != cudaSuccess ||
        cudaMemsetAsync(dbias->typed_data(), 0, dbias->size_bytes(), stream) != cudaSuccess)
        return ffi::Error::Internal("cudaMemsetAsync failed");
    return Status(mlb_dense_dx_lnbwd_tc(stream, DZ.typed_data(), W.typed_data(), scale.typed_data(), bias.typed_data(),
                                        XH.typed_data(), rstd.typed_data(), DZ_out->typed_data(),
                                        dscale->typed_data(), dbias->typed_data(), static_cast<int>(RowsOf(dz)), K, HN,
                                        K, K), "mlb_dense_dx_lnbwd_tc");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_dense_dx_lnbwd_tc_ffi, DenseDxLnBwdTcImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::BF16>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>(),
    {ffi::Traits::kCmdBufferCompatible});

// The fused PPO loss + head gradients + the five per-minibatch metrics (ml/ppo.py:129-262, 351-362).
// head [rows, ld] f32; d_head f32 [rows, ld]; stats = raw mlb_ppo_stats bytes; ws from mlb_ppo_loss_workspace.
static ffi::Error PpoLossImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> head, ffi::Buffer<ffi::S32> actions,
                              ffi::Buffer<ffi::F32> old_log_probs, ffi::Buffer<ffi::F32> advantages,
                              ffi::Buffer<ffi::F32> returns, ffi::Buffer<ffi::F32> adv_mean_rstd,
                              ffi::ResultBuffer<ffi::F32> d_head, ffi::ResultBuffer<ffi::F32> d_bias,
                              ffi::ResultBuffer<ffi::U8> stats, ffi::ResultBuffer<ffi::U8> ws,
                              ffi::Span<const int32_t> buckets, ffi::Span<const float> obj_scale,
                              ffi::Span<const float> ent_scale, int64_t minibatch, float clip_coef,
                              float value_loss_coef, int32_t flags) {
    const auto d = head.dimensions();
    const int ld = static_cast<int>(d[d.size() - 1]);
    if (stats->size_bytes() < sizeof(mlb_ppo_stats)) return ffi::Error::InvalidArgument("stats buffer too small");
    if (cudaMemsetAsync(d_bias->typed_data(), 0, d_bias->size_bytes(), stream) != cudaSuccess)
        return ffi::Error::Internal("cudaMemsetAsync failed");
    return Status(mlb_ppo_loss_f32(stream, head.typed_data(), ld, actions.typed_data(), old_log_probs.typed_data(),
                                   advantages.typed_data(), returns.typed_data(), nullptr, nullptr,
                                   adv_mean_rstd.typed_data(), nullptr, buckets.data(), obj_scale.data(),
                                   ent_scale.data(), static_cast<int>(buckets.size()), RowsOf(d),
                                   static_cast<long long>(minibatch), clip_coef, value_loss_coef, flags,
                                   d_head->typed_data(), d_bias->typed_data(),
                                   reinterpret_cast<mlb_ppo_stats*>(stats->typed_data()), ws->typed_data(),
                                   ws->size_bytes(), nullptr, 1), "mlb_ppo_loss_f32");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_ppo_loss_f32_ffi, PpoLossImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>()
        .Ret<ffi::Buffer<ffi::U8>>()
        .Attr<ffi::Span<const int32_t>>("buckets").Attr<ffi::Span<const float>>("obj_scale")
        .Attr<ffi::Span<const float>>("ent_scale").Attr<int64_t>("minibatch").Attr<float>("clip_coef")
        .Attr<float>("value_loss_coef").Attr<int32_t>("flags"),
    {ffi::Traits::kCmdBufferCompatible});

// clip_by_global_norm -> adam -> kernel re-projection / LayerNorm renorm over the flat arena (ml/ppo.py:283-338),
// in place: params / m / v / step / grads are aliased operand-result pairs (grads is cleared behind the update).
static ffi::Error OptimizerImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> grads,
                                ffi::Buffer<ffi::F32> m, ffi::Buffer<ffi::F32> v, ffi::Buffer<ffi::S32> step,
                                ffi::Buffer<ffi::U8> segments, ffi::Buffer<ffi::U32> sync_state,
                                ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> grads_out,
                                ffi::ResultBuffer<ffi::F32> m_out, ffi::ResultBuffer<ffi::F32> v_out,
                                ffi::ResultBuffer<ffi::S32> step_out, ffi::ResultBuffer<ffi::U32> sync_out,
                                ffi::ResultBuffer<ffi::F64> grad_sumsq, ffi::ResultBuffer<ffi::U8> ws, float lr,
                                float b1, float b2, float eps, float max_grad_norm) {
    for (auto e : {Aliased(params.typed_data(), params_out->typed_data(), "params"),
                   Aliased(grads.typed_data(), grads_out->typed_data(), "grads"),
                   Aliased(m.typed_data(), m_out->typed_data(), "adam m"),
                   Aliased(v.typed_data(), v_out->typed_data(), "adam v"),
                   Aliased(step.typed_data(), step_out->typed_data(), "adam step"),
                   Aliased(sync_state.typed_data(), sync_out->typed_data(), "barrier state")})
        if (!e.success()) return e;
    const int nseg = static_cast<int>(segments.size_bytes() / sizeof(mlb_segment));
    return Status(mlb_optimizer_step_fused(stream, params_out->typed_data(), grads_out->typed_data(),
                                           m_out->typed_data(), v_out->typed_data(),
                                           static_cast<long long>(params.element_count()),
                                           reinterpret_cast<const mlb_segment*>(segments.typed_data()), nseg, nullptr,
                                           step_out->typed_data(), grad_sumsq->typed_data(), 0, lr, b1, b2, eps,
                                           max_grad_norm, 1.0f, sync_out->typed_data(), ws->typed_data(),
                                           ws->size_bytes(), grads_out->typed_data()), "mlb_optimizer_step_fused");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mlb_optimizer_step_fused_ffi, OptimizerImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::U8>>()
        .Arg<ffi::Buffer<ffi::U32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U32>>()
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>()
        .Attr<float>("lr").Attr<float>("b1").Attr<float>("b2").Attr<float>("eps").Attr<float>("max_grad_norm"),
    {ffi::Traits::kCmdBufferCompatible});
#endif  // __has_include("xla/ffi/api/ffi.h")
