// K2/K3: moments, z-score, Metric reductions, per-minibatch statistics, EMA normaliser.
//
// Replaces zscore_data (ml/algo_common.py:133-140), Metric.init_from_data
// (ml/metrics.py:31-48) and EMANormalizer (ml/moving_avg.py:48-198).  Reductions accumulate
// in double at thread level (B200 runs FP64 adds at half FP32 rate; these kernels are
// HBM-bound), warp-shuffle + shared-memory tree per block, per-block partials in the
// caller's workspace, one finishing block -- no atomics, deterministic.
#include "common.cuh"

namespace {

struct MomPartial { double s, ss; float mn, mx; };

constexpr int RED_BLOCK = 256;
constexpr int RED_MAX_BLOCKS = MLB_NUM_SMS * 8;

unsigned red_grid(long long n) {
    long long b = (n + (long long)RED_BLOCK * 16 - 1) / ((long long)RED_BLOCK * 16);
    if (b < 1) b = 1;
    if (b > RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
    return (unsigned)b;
}

__global__ void __launch_bounds__(RED_BLOCK)
moments_partial_kernel(const float* __restrict__ x, long long n, MomPartial* __restrict__ part) {
    double s = 0.0, ss = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const long long n4 = vec_ok ? n / 4 : 0;
    for (long long i = tid; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        // fp32 within the 4-vector, double across vectors
        const float ls = (v.x + v.y) + (v.z + v.w);
        s += (double)ls;
        ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
        mn = fminf(mn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
        mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    for (long long i = n4 * 4 + tid; i < n; i += stride) {
        const float v = x[i];
        s += (double)v; ss += (double)v * v;
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    __shared__ double smd[32];
    __shared__ float smf[32];
    s = block_sum_d(s, smd);
    ss = block_sum_d(ss, smd);
    mn = block_min_f(mn, smf);
    mx = block_max_f(mx, smf);
    if (threadIdx.x == 0) { part[blockIdx.x].s = s; part[blockIdx.x].ss = ss;
                            part[blockIdx.x].mn = mn; part[blockIdx.x].mx = mx; }
}

// mode 0: out f32[4] = {mean, rstd, var, n};  mode 1: mlb_metric
__global__ void __launch_bounds__(RED_BLOCK)
moments_final_kernel(const MomPartial* __restrict__ part, int nparts, double count,
                     float var_floor, float* __restrict__ out4, mlb_metric* __restrict__ metric) {
    double s = 0.0, ss = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    for (int b = threadIdx.x; b < nparts; b += blockDim.x) {
        s += part[b].s; ss += part[b].ss;
        mn = fminf(mn, part[b].mn); mx = fmaxf(mx, part[b].mx);
    }
    __shared__ double smd[32];
    __shared__ float smf[32];
    s = block_sum_d(s, smd);
    ss = block_sum_d(ss, smd);
    mn = block_min_f(mn, smf);
    mx = block_max_f(mx, smf);
    if (threadIdx.x == 0) {
        const double mean = s / count;
        double m2 = ss - s * mean;
        if (m2 < 0.0) m2 = 0.0;
        if (out4) {
            const float var = (float)(m2 / count);
            out4[0] = (float)mean;
            out4[1] = rsqrtf(fmaxf(var, var_floor));
            out4[2] = var;
            out4[3] = (float)count;
        }
        if (metric) {
            metric->mean = (float)mean; metric->m2 = (float)m2;
            metric->min = mn; metric->max = mx; metric->count = (int32_t)count;
        }
    }
}

__global__ void __launch_bounds__(256)
zscore_apply_kernel(const float* __restrict__ x, float* __restrict__ out, long long n,
                    const float* __restrict__ mean_rstd) {
    const float mean = mean_rstd[0], rstd = mean_rstd[1];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const long long n4 = vec_ok ? n / 4 : 0;
    for (long long i = tid; i < n4; i += stride) {
        float4 v = ld_stream_f4(x + 4 * i);
        v.x = __fmul_rn(__fadd_rn(v.x, -mean), rstd);
        v.y = __fmul_rn(__fadd_rn(v.y, -mean), rstd);
        v.z = __fmul_rn(__fadd_rn(v.z, -mean), rstd);
        v.w = __fmul_rn(__fadd_rn(v.w, -mean), rstd);
        st_stream_f4(out + 4 * i, v);
    }
    for (long long i = n4 * 4 + tid; i < n; i += stride)
        out[i] = __fmul_rn(__fadd_rn(x[i], -mean), rstd);
}

// One thread per column; emits {sum, sumsq} per (chunk, column) trajectory.
__global__ void __launch_bounds__(256)
traj_moments_kernel(const float* __restrict__ x, int T, long long N, int C,
                    double* __restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int Tp = T / C;
    for (int c = 0; c < C; ++c) {
        double s = 0.0, ss = 0.0;
        const float* p = x + (long long)c * Tp * N + n;
#pragma unroll 8
        for (int t = 0; t < Tp; ++t) {
            const float v = __ldg(p + (long long)t * N);
            s += (double)v; ss += (double)v * (double)v;
        }
        double2 r = make_double2(s, ss);
        reinterpret_cast<double2*>(out)[(long long)c * N + n] = r;
    }
}

// grid = E * (J/M) blocks; block k reduces the M trajectories of minibatch k.
__global__ void __launch_bounds__(256)
mb_moments_kernel(const double* __restrict__ tm, const int32_t* __restrict__ perm,
                  long long J, long long M, int Tp, float var_floor, float* __restrict__ out,
                  double* __restrict__ raw) {
    const long long nmb = J / M;
    const long long e = blockIdx.x / nmb, k = blockIdx.x % nmb;
    const int32_t* idx = perm + e * J + k * M;
    double s = 0.0, ss = 0.0;
    for (long long m = threadIdx.x; m < M; m += blockDim.x) {
        const double2 v = reinterpret_cast<const double2*>(tm)[idx[m]];
        s += v.x; ss += v.y;
    }
    __shared__ double smd[32];
    s = block_sum_d(s, smd);
    ss = block_sum_d(ss, smd);
    if (threadIdx.x == 0) {
        if (raw) { raw[2ll * blockIdx.x] = s; raw[2ll * blockIdx.x + 1] = ss; }
        if (out) {
            const double count = (double)M * (double)Tp;
            const double mean = s / count;
            double m2 = ss - s * mean;
            if (m2 < 0.0) m2 = 0.0;
            const float var = (float)(m2 / count);
            float* o = out + 4ll * blockIdx.x;
            o[0] = (float)mean; o[1] = rsqrtf(fmaxf(var, var_floor)); o[2] = var; o[3] = (float)count;
        }
    }
}

// raw [K][2] = {sum, sumsq} (already summed over ranks) -> out [K][4]
__global__ void moments_finalize_kernel(const double* __restrict__ raw, int K, double count,
                                        float var_floor, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const double s = raw[2 * k], ss = raw[2 * k + 1];
    const double mean = s / count;
    double m2 = ss - s * mean;
    if (m2 < 0.0) m2 = 0.0;
    const float var = (float)(m2 / count);
    float* o = out + 4ll * k;
    o[0] = (float)mean; o[1] = rsqrtf(fmaxf(var, var_floor)); o[2] = var; o[3] = (float)count;
}

// EMANormalizer.update_estimates (ml/moving_avg.py:131-181), one thread per feature.
__device__ __forceinline__ void ema_update_one(float& mu, float& inv_sigma, float& sigma,
                                               float& mu_b, float& ssq_b, int N,
                                               float x_mean, float x_var, float decay, float eps) {
    const float mean_delta = __fadd_rn(x_mean, -mu);
    const float oma = decay;
    const float alpha = __fadd_rn(1.f, -oma);
    const float newN = (float)(N + 1);
    const float new_mu_b = __fadd_rn(__fmul_rn(oma, mu_b), __fmul_rn(alpha, x_mean));
    const float cross = __fmul_rn(__fmul_rn(__fdiv_rn((float)N, newN), __fmul_rn(oma, alpha)),
                                  __fmul_rn(mean_delta, mean_delta));
    const float new_ssq_b = __fadd_rn(__fadd_rn(__fmul_rn(oma, ssq_b), __fmul_rn(alpha, x_var)), cross);
    // -1 / expm1(N * log(decay)); the correction factor is evaluated in double and rounded
    // (expm1f/logf are not correctly rounded on device; the reference's XLA versions are not
    // either -- tolerance rel 1e-5 in the tests).
    const float arg = __fmul_rn(newN, (float)log((double)oma));
    const float bias_corr = (float)(-1.0 / expm1((double)arg));
    const float new_mu = __fmul_rn(new_mu_b, bias_corr);
    const float new_ssq = __fmul_rn(new_ssq_b, bias_corr);
    const float new_inv = (float)(1.0 / sqrt((double)fmaxf(new_ssq, eps)));
    mu = new_mu; inv_sigma = new_inv; sigma = __fdiv_rn(1.f, new_inv);
    mu_b = new_mu_b; ssq_b = new_ssq_b;
}

__global__ void ema_update_kernel(float* __restrict__ state, int dim,
                                  const float* __restrict__ bm, const float* __restrict__ bv,
                                  float decay, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int* Np = reinterpret_cast<int*>(state + 5 * dim);
    const int N = *Np;
    if (i < dim) {
        float mu = state[i], inv = state[dim + i], sg = state[2 * dim + i],
              mub = state[3 * dim + i], sqb = state[4 * dim + i];
        ema_update_one(mu, inv, sg, mub, sqb, N, bm[i], bv[i], decay, eps);
        state[i] = mu; state[dim + i] = inv; state[2 * dim + i] = sg;
        state[3 * dim + i] = mub; state[4 * dim + i] = sqb;
    }
    // single-block launches only (dim <= 1024) so this ordering is safe
    __syncthreads();
    if (i == 0) *Np = N + 1;
}

__global__ void ema_scan_kernel(float* __restrict__ state, const float* __restrict__ mbm, int K,
                                float decay, float eps, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float mu = state[0], inv = state[1], sg = state[2], mub = state[3], sqb = state[4];
    int N = *reinterpret_cast<int*>(state + 5);
    for (int k = 0; k < K; ++k) {
        out[4 * k + 0] = mu;
        out[4 * k + 1] = sg;
        ema_update_one(mu, inv, sg, mub, sqb, N, mbm[4 * k + 0], mbm[4 * k + 2], decay, eps);
        ++N;
        out[4 * k + 2] = mu;
        out[4 * k + 3] = inv;
    }
    state[0] = mu; state[1] = inv; state[2] = sg; state[3] = mub; state[4] = sqb;
    *reinterpret_cast<int*>(state + 5) = N;
}

template <bool INVERT>
__global__ void __launch_bounds__(256)
ema_apply_kernel(const float* __restrict__ state, int dim, const float* __restrict__ x,
                 float* __restrict__ out, long long total) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < total; i += stride) {
        const int f = (int)(i % dim);
        if (INVERT) out[i] = __fadd_rn(__fmul_rn(x[i], state[2 * dim + f]), state[f]);
        else out[i] = __fmul_rn(__fadd_rn(x[i], -state[f]), state[dim + f]);
    }
}

__global__ void __launch_bounds__(256)
env_returns_kernel(const float* __restrict__ r, const uint8_t* __restrict__ d,
                   float* __restrict__ er, float* __restrict__ trace, long long N, float gamma) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float v = __fadd_rn(r[i], __fmul_rn(gamma, er[i]));
    if (trace) trace[i] = v;
    er[i] = d[i] ? 0.f : v;
}

__global__ void __launch_bounds__(256)
post_step_kernel(const float* __restrict__ r, const uint8_t* __restrict__ d, float* __restrict__ rs,
                 uint8_t* __restrict__ ds, float* __restrict__ er, float* __restrict__ trace,
                 long long N, float gamma) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float ri = r[i];
    const uint8_t di = d[i];
    rs[i] = ri;
    ds[i] = di;
    const float v = __fadd_rn(ri, __fmul_rn(gamma, er[i]));
    if (trace) trace[i] = v;
    er[i] = di ? 0.f : v;
}

unsigned ew_grid(long long n, int per_thread) {
    long long b = (n + 256ll * per_thread - 1) / (256ll * per_thread);
    if (b < 1) b = 1;
    const long long cap = (long long)MLB_NUM_SMS * 16;
    return (unsigned)(b > cap ? cap : b);
}

}  // namespace

MLB_API size_t mlb_moments_workspace(long long n) { return (size_t)red_grid(n) * sizeof(MomPartial); }

static int moments_impl(void* stream, const float* x, long long n, float var_floor, float* out4,
                        mlb_metric* metric, void* ws, size_t ws_bytes) {
    MLB_REQUIRE(x && n > 0 && (out4 || metric));
    const unsigned g = red_grid(n);
    if (!ws || ws_bytes < g * sizeof(MomPartial)) return MLB_EWS;
    cudaStream_t s = mlb_stream(stream);
    MomPartial* part = reinterpret_cast<MomPartial*>(ws);
    moments_partial_kernel<<<g, RED_BLOCK, 0, s>>>(x, n, part);
    MLB_CHECK_LAUNCH();
    moments_final_kernel<<<1, RED_BLOCK, 0, s>>>(part, (int)g, (double)n, var_floor, out4, metric);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_moments_f32(void* stream, const float* x, long long n, float var_floor,
                            float* out4, void* ws, size_t ws_bytes) {
    return moments_impl(stream, x, n, var_floor, out4, nullptr, ws, ws_bytes);
}

MLB_API int mlb_metric_f32(void* stream, const float* x, long long n, mlb_metric* out,
                           void* ws, size_t ws_bytes) {
    return moments_impl(stream, x, n, 0.f, nullptr, out, ws, ws_bytes);
}

MLB_API int mlb_zscore_apply_f32(void* stream, const float* x, float* out, long long n,
                                 const float* mean_rstd) {
    MLB_REQUIRE(x && out && mean_rstd && n >= 0);
    if (n == 0) return MLB_OK;
    zscore_apply_kernel<<<ew_grid(n, 16), 256, 0, mlb_stream(stream)>>>(x, out, n, mean_rstd);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_zscore_f32(void* stream, const float* x, float* out, long long n, float* out4,
                           void* ws, size_t ws_bytes) {
    MLB_REQUIRE(out4);
    int rc = mlb_moments_f32(stream, x, n, 1e-5f, out4, ws, ws_bytes);
    if (rc) return rc;
    return mlb_zscore_apply_f32(stream, x, out, n, out4);
}

MLB_API int mlb_traj_moments_f32(void* stream, const float* x, int T, long long N, int C,
                                 double* out) {
    MLB_REQUIRE(x && out && T > 0 && N > 0 && C > 0 && T % C == 0);
    MLB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int block = N >= MLB_NUM_SMS * 256 ? 256 : 64;
    traj_moments_kernel<<<mlb_cdiv(N, block), block, 0, mlb_stream(stream)>>>(x, T, N, C, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_mb_moments_f32(void* stream, const double* traj_moments, const int32_t* perm,
                               int E, long long J, long long M, int Tp, float var_floor,
                               float* out, double* raw_out) {
    MLB_REQUIRE(traj_moments && perm && (out || raw_out) && E > 0 && J > 0 && M > 0 && J % M == 0 && Tp > 0);
    const unsigned g = (unsigned)(E * (J / M));
    mb_moments_kernel<<<g, 256, 0, mlb_stream(stream)>>>(traj_moments, perm, J, M, Tp, var_floor, out, raw_out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_moments_finalize_f32(void* stream, const double* raw, int K, double count,
                                     float var_floor, float* out) {
    MLB_REQUIRE(raw && out && K > 0 && count > 0);
    moments_finalize_kernel<<<mlb_cdiv(K, 128), 128, 0, mlb_stream(stream)>>>(raw, K, count, var_floor, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_ema_update_f32(void* stream, float* state, int dim, const float* batch_mean,
                               const float* batch_var, float decay, float eps) {
    MLB_REQUIRE(state && batch_mean && batch_var && dim > 0 && dim <= 1024);
    ema_update_kernel<<<1, ((dim + 31) / 32) * 32, 0, mlb_stream(stream)>>>(state, dim, batch_mean,
                                                                            batch_var, decay, eps);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_ema_scan_f32(void* stream, float* state, const float* mb_moments, int K,
                             float decay, float eps, float* out) {
    MLB_REQUIRE(state && mb_moments && out && K > 0);
    ema_scan_kernel<<<1, 32, 0, mlb_stream(stream)>>>(state, mb_moments, K, decay, eps, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_ema_normalize_f32(void* stream, const float* state, int dim, const float* x,
                                  float* out, long long rows) {
    MLB_REQUIRE(state && x && out && dim > 0 && rows >= 0);
    if (rows == 0) return MLB_OK;
    ema_apply_kernel<false><<<ew_grid(rows * dim, 4), 256, 0, mlb_stream(stream)>>>(state, dim, x, out, rows * dim);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_ema_invert_f32(void* stream, const float* state, int dim, const float* x,
                               float* out, long long rows) {
    MLB_REQUIRE(state && x && out && dim > 0 && rows >= 0);
    if (rows == 0) return MLB_OK;
    ema_apply_kernel<true><<<ew_grid(rows * dim, 4), 256, 0, mlb_stream(stream)>>>(state, dim, x, out, rows * dim);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_post_step_store_f32(void* stream, const float* rewards, const uint8_t* dones,
                                    float* reward_slab, uint8_t* done_slab, float* env_returns,
                                    float* trace, long long N, float gamma) {
    MLB_REQUIRE(rewards && dones && reward_slab && done_slab && env_returns && N >= 0);
    if (N == 0) return MLB_OK;
    cudaError_t e = launch_pdl(post_step_kernel, dim3(mlb_cdiv(N, 256)), dim3(256), 0, mlb_stream(stream), rewards, dones,
                               reward_slab, done_slab, env_returns, trace, N, gamma);
    if (e != cudaSuccess) return (int)e;
    return MLB_OK;
}

MLB_API int mlb_env_returns_f32(void* stream, const float* rewards, const uint8_t* dones,
                                float* env_returns, float* trace, long long N, float gamma) {
    MLB_REQUIRE(rewards && dones && env_returns && N >= 0);
    if (N == 0) return MLB_OK;
    env_returns_kernel<<<mlb_cdiv(N, 256), 256, 0, mlb_stream(stream)>>>(rewards, dones, env_returns, trace, N, gamma);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
