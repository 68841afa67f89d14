// K1: GAE(lambda) + returns -- reverse-time, column-parallel scan over [T, N] rollout buffers.
//
// Replaces compute_advantages (ml/algo_common.py:84-130), compute_returns (:45-81), the
// separate `returns = advantages + values` pass (ml/rollouts.py:769), the value-normaliser
// invert (ml/rollouts.py:726-741) and the four full-buffer Metric reductions
// (ml/rollouts.py:806-816).  One pass, every byte touched once: 17*T*N + 4*N bytes.
//
// Mapping: one thread owns VEC adjacent columns (VEC = 4 -> 128-bit loads of rewards/values
// and one 32-bit load of four done bytes) and walks time backwards with the two carries
// (next_advantage, next_value) in registers.  Time is consumed in chunks of U steps: all 3*U
// loads of a chunk are issued before the serial recurrence so >= U*VEC*9 bytes per thread are
// in flight (HBM latency hiding; the recurrence itself is 4 flops per element).
// Arithmetic is deliberately NOT contracted into FMAs (__fmul_rn/__fadd_rn) so the result is
// bit-identical to the float32 reference evaluated op by op.
#include <stdlib.h>

#include "common.cuh"

namespace {

template <int VEC> struct Vec;
template <> struct Vec<4> {
    static __device__ __forceinline__ void ldf(const float* p, float (&o)[4]) {
        float4 v = ld_stream_f4(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void ldd(const uint8_t* p, uint32_t& o) { o = ld_stream_u32(p); }
    static __device__ __forceinline__ void stf(float* p, const float (&o)[4]) {
        st_stream_f4(p, make_float4(o[0], o[1], o[2], o[3]));
    }
};
template <> struct Vec<2> {
    static __device__ __forceinline__ void ldf(const float* p, float (&o)[2]) {
        float2 v = __ldcs(reinterpret_cast<const float2*>(p)); o[0] = v.x; o[1] = v.y;
    }
    static __device__ __forceinline__ void ldd(const uint8_t* p, uint32_t& o) {
        o = __ldcs(reinterpret_cast<const unsigned short*>(p));
    }
    static __device__ __forceinline__ void stf(float* p, const float (&o)[2]) {
        __stcs(reinterpret_cast<float2*>(p), make_float2(o[0], o[1]));
    }
};
template <> struct Vec<1> {
    static __device__ __forceinline__ void ldf(const float* p, float (&o)[1]) { o[0] = __ldcs(p); }
    static __device__ __forceinline__ void ldd(const uint8_t* p, uint32_t& o) { o = __ldcs(p); }
    static __device__ __forceinline__ void stf(float* p, const float (&o)[1]) { __stcs(p, o[0]); }
};

struct GaeStatPartial {        // per-block partial of the 4 metric streams
    double s[4], ss[4];
    float mn[4], mx[4];
};

struct ThreadStats {
    double s[4], ss[4];
    float mn[4], mx[4];
    float cs[4], css[4];          // fp32 partials of the current time chunk (<= U*VEC terms)
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i] = 0.0; ss[i] = 0.0; mn[i] = INFINITY; mx[i] = -INFINITY; cs[i] = 0.f; css[i] = 0.f; }
    }
    __device__ __forceinline__ void add(int i, float x) {
        cs[i] += x; css[i] = fmaf(x, x, css[i]);
        mn[i] = fminf(mn[i], x); mx[i] = fmaxf(mx[i], x);
    }
    // a chunk holds at most 32 terms: the fp32 partial is exact to ~2e-6 relative, the running
    // totals over the T/U chunks of a column stay in fp64
    __device__ __forceinline__ void fold() {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i] += (double)cs[i]; ss[i] += (double)css[i]; cs[i] = 0.f; css[i] = 0.f; }
    }
};

template <int VEC, int U, bool STATS>
__global__ void __launch_bounds__(256)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
           const uint8_t* __restrict__ dones, const float* __restrict__ bootstrap,
           float* __restrict__ adv, float* __restrict__ ret,
           int T, long long N, float gamma, float gl, const float* __restrict__ vn,
           GaeStatPartial* __restrict__ partials) {
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const bool active = col < N;
    const float mu = vn ? vn[0] : 0.f;
    const float sigma = vn ? vn[2] : 1.f;   // EMA state layout (dim 1): mu, inv_sigma, sigma, ...

    ThreadStats st;
    if (STATS) st.init();

    if (active) {
        float na[VEC], nv[VEC];
        {
            float b[VEC];
            Vec<VEC>::ldf(bootstrap + col, b);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                na[j] = 0.f;
                nv[j] = __fadd_rn(__fmul_rn(b[j], sigma), mu);
            }
        }
        int t_hi = T;
        // main: chunks of U time steps.  Software pipeline: the 3*U loads of chunk k+1 are issued
        // BEFORE the serial recurrence of chunk k runs, so a thread always has a full chunk
        // (U*VEC*9 bytes) in flight underneath its arithmetic -- with N as the only parallel axis
        // (the scan is exact, so time cannot be split) this is what fills the HBM pipe for
        // N ~ 64K..256K columns.
        float r[2][U][VEC], v[2][U][VEC];
        uint32_t d[2][U];
        auto load_chunk = [&](int buf, int thi) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long off = (long long)(thi - 1 - u) * N + col;
                Vec<VEC>::ldf(rewards + off, r[buf][u]);
                Vec<VEC>::ldf(values + off, v[buf][u]);
                Vec<VEC>::ldd(dones + off, d[buf][u]);
            }
        };
        auto scan_chunk = [&](int buf, int thi) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long off = (long long)(thi - 1 - u) * N + col;
                float a[VEC], rt[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const bool done = ((d[buf][u] >> (8 * j)) & 0xffu) != 0u;
                    const float vv = __fadd_rn(__fmul_rn(v[buf][u][j], sigma), mu);
                    const float nvj = done ? 0.f : nv[j];
                    const float naj = done ? 0.f : na[j];
                    const float td = __fadd_rn(__fadd_rn(r[buf][u][j], __fmul_rn(gamma, nvj)), -vv);
                    a[j] = __fadd_rn(td, __fmul_rn(gl, naj));
                    rt[j] = __fadd_rn(a[j], vv);
                    na[j] = a[j];
                    nv[j] = vv;
                    if (STATS) { st.add(0, r[buf][u][j]); st.add(1, vv); st.add(2, rt[j]); st.add(3, a[j]); }
                }
                Vec<VEC>::stf(adv + off, a);
                if (ret) Vec<VEC>::stf(ret + off, rt);
            }
            if (STATS) st.fold();
        };
        if (t_hi >= U) load_chunk(0, t_hi);
        while (t_hi >= U) {
            if (t_hi - U >= U) load_chunk(1, t_hi - U);
            scan_chunk(0, t_hi);
            t_hi -= U;
            if (t_hi < U) break;
            if (t_hi - U >= U) load_chunk(0, t_hi - U);
            scan_chunk(1, t_hi);
            t_hi -= U;
        }
        // remainder (T % U steps)
        for (int t = t_hi - 1; t >= 0; --t) {
            const long long off = (long long)t * N + col;
            float r[VEC], v[VEC], a[VEC], rt[VEC];
            uint32_t d;
            Vec<VEC>::ldf(rewards + off, r);
            Vec<VEC>::ldf(values + off, v);
            Vec<VEC>::ldd(dones + off, d);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool done = ((d >> (8 * j)) & 0xffu) != 0u;
                const float vv = __fadd_rn(__fmul_rn(v[j], sigma), mu);
                const float nvj = done ? 0.f : nv[j];
                const float naj = done ? 0.f : na[j];
                const float td = __fadd_rn(__fadd_rn(r[j], __fmul_rn(gamma, nvj)), -vv);
                a[j] = __fadd_rn(td, __fmul_rn(gl, naj));
                rt[j] = __fadd_rn(a[j], vv);
                na[j] = a[j];
                nv[j] = vv;
                if (STATS) { st.add(0, r[j]); st.add(1, vv); st.add(2, rt[j]); st.add(3, a[j]); }
            }
            Vec<VEC>::stf(adv + off, a);
            if (ret) Vec<VEC>::stf(ret + off, rt);
        }
        if (STATS) st.fold();
    }

    if (STATS) {
        __shared__ double smd[32];
        __shared__ float smf[32];
        GaeStatPartial p;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            p.s[i] = block_sum_d(st.s[i], smd);
            p.ss[i] = block_sum_d(st.ss[i], smd);
            p.mn[i] = block_min_f(st.mn[i], smf);
            p.mx[i] = block_max_f(st.mx[i], smf);
        }
        if (threadIdx.x == 0) partials[blockIdx.x] = p;
    }
}

__global__ void __launch_bounds__(256)
gae_stats_finalize(const GaeStatPartial* __restrict__ partials, int nblocks, double count,
                   mlb_metric* __restrict__ out) {
    __shared__ double smd[32];
    __shared__ float smf[32];
    for (int i = 0; i < 4; ++i) {
        double s = 0.0, ss = 0.0;
        float mn = INFINITY, mx = -INFINITY;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
            s += partials[b].s[i]; ss += partials[b].ss[i];
            mn = fminf(mn, partials[b].mn[i]); mx = fmaxf(mx, partials[b].mx[i]);
        }
        s = block_sum_d(s, smd);
        ss = block_sum_d(ss, smd);
        mn = block_min_f(mn, smf);
        mx = block_max_f(mx, smf);
        if (threadIdx.x == 0) {
            const double mean = s / count;
            double m2 = ss - s * mean;
            if (m2 < 0.0) m2 = 0.0;
            out[i].mean = (float)mean;
            out[i].m2 = (float)m2;
            out[i].min = mn;
            out[i].max = mx;
            out[i].count = (int32_t)count;
        }
    }
}

template <int VEC, int U>
__global__ void __launch_bounds__(256)
returns_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
               const float* __restrict__ bootstrap, float* __restrict__ ret,
               int T, long long N, float gamma) {
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (col >= N) return;
    float nr[VEC];
    Vec<VEC>::ldf(bootstrap + col, nr);
    int t_hi = T;
    for (; t_hi >= U; t_hi -= U) {
        float r[U][VEC];
        uint32_t d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long off = (long long)(t_hi - 1 - u) * N + col;
            Vec<VEC>::ldf(rewards + off, r[u]);
            Vec<VEC>::ldd(dones + off, d[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long off = (long long)(t_hi - 1 - u) * N + col;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool done = ((d[u] >> (8 * j)) & 0xffu) != 0u;
                nr[j] = __fadd_rn(r[u][j], __fmul_rn(gamma, done ? 0.f : nr[j]));
            }
            Vec<VEC>::stf(ret + off, nr);
        }
    }
    for (int t = t_hi - 1; t >= 0; --t) {
        const long long off = (long long)t * N + col;
        float r[VEC];
        uint32_t d;
        Vec<VEC>::ldf(rewards + off, r);
        Vec<VEC>::ldd(dones + off, d);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const bool done = ((d >> (8 * j)) & 0xffu) != 0u;
            nr[j] = __fadd_rn(r[j], __fmul_rn(gamma, done ? 0.f : nr[j]));
        }
        Vec<VEC>::stf(ret + off, nr);
    }
}

struct ScanPlan { int vec; int block; unsigned grid; int deep; };

// Pick the widest vector that still leaves >= ~2 resident 256-thread CTAs per SM; narrower
// vectors (more threads) for mid-size N so all 148 SMs have loads in flight.
ScanPlan plan_scan(long long N, bool align16) {
    ScanPlan p;
    const long long want = (long long)MLB_NUM_SMS * 512;
    if (align16 && (N % 4 == 0) && N / 4 >= want) p.vec = 4;
    else if (align16 && (N % 2 == 0) && N / 2 >= want) p.vec = 2;
    else p.vec = 1;
    const long long threads = N / p.vec;
    p.block = threads >= (long long)MLB_NUM_SMS * 512 ? 256 : (threads >= MLB_NUM_SMS * 128 ? 128 : 64);
    static const int forced = [] { const char* v = getenv("MLB_GAE_BLOCK"); return v ? atoi(v) : 0; }();
    if (forced >= 32 && forced <= 1024 && forced % 32 == 0) p.block = forced;
    p.grid = mlb_cdiv(threads, p.block);
    // fewer than ~1024 threads per SM: a thread must keep more bytes in flight to cover the HBM
    // latency-bandwidth product (measured r2: N = 64K columns, U = 8 -> 0.69 of peak) -> 16-step chunks
    p.deep = (p.vec == 1 && threads < (long long)MLB_NUM_SMS * 1024) ? 1 : 0;
    return p;
}

}  // namespace

MLB_API size_t mlb_gae_workspace(int T, long long N) {
    (void)T;
    // worst case: VEC=1, 64-thread blocks
    return (size_t)mlb_cdiv(N, 64) * sizeof(GaeStatPartial);
}

MLB_API int mlb_gae_f32(void* stream, const float* rewards, const float* values,
                        const uint8_t* dones, const float* bootstrap, float* advantages,
                        float* returns, int T, long long N, float gamma, float gamma_lambda,
                        const float* vn_mu_sigma, mlb_metric* metrics, void* ws, size_t ws_bytes) {
    MLB_REQUIRE(rewards && values && dones && bootstrap && advantages);
    MLB_REQUIRE(T >= 0 && N >= 0);
    if (T == 0 || N == 0) return MLB_OK;
    const bool al = mlb_aligned16(rewards) && mlb_aligned16(values) && mlb_aligned16(advantages) &&
                    (returns == nullptr || mlb_aligned16(returns)) && mlb_aligned16(bootstrap) &&
                    ((reinterpret_cast<uintptr_t>(dones) & 3) == 0);
    const ScanPlan p = plan_scan(N, al);
    cudaStream_t s = mlb_stream(stream);
    GaeStatPartial* partials = nullptr;
    if (metrics) {
        if (ws == nullptr || ws_bytes < (size_t)p.grid * sizeof(GaeStatPartial)) return MLB_EWS;
        partials = reinterpret_cast<GaeStatPartial*>(ws);
    }
#define LAUNCH(V, UU)                                                                          \
    do {                                                                                       \
        if (metrics)                                                                           \
            gae_kernel<V, UU, true><<<p.grid, p.block, 0, s>>>(rewards, values, dones,         \
                bootstrap, advantages, returns, T, N, gamma, gamma_lambda, vn_mu_sigma,        \
                partials);                                                                     \
        else                                                                                   \
            gae_kernel<V, UU, false><<<p.grid, p.block, 0, s>>>(rewards, values, dones,        \
                bootstrap, advantages, returns, T, N, gamma, gamma_lambda, vn_mu_sigma,        \
                nullptr);                                                                      \
    } while (0)
    if (p.vec == 4) LAUNCH(4, 4);
    else if (p.vec == 2) LAUNCH(2, 8);
    else if (p.deep) LAUNCH(1, 16);
    else LAUNCH(1, 8);
#undef LAUNCH
    MLB_CHECK_LAUNCH();
    if (metrics) {
        gae_stats_finalize<<<1, 256, 0, s>>>(partials, (int)p.grid, (double)T * (double)N, metrics);
        MLB_CHECK_LAUNCH();
    }
    return MLB_OK;
}

MLB_API int mlb_returns_f32(void* stream, const float* rewards, const uint8_t* dones,
                            const float* bootstrap, float* returns, int T, long long N,
                            float gamma) {
    MLB_REQUIRE(rewards && dones && bootstrap && returns);
    MLB_REQUIRE(T >= 0 && N >= 0);
    if (T == 0 || N == 0) return MLB_OK;
    const bool al = mlb_aligned16(rewards) && mlb_aligned16(returns) && mlb_aligned16(bootstrap) &&
                    ((reinterpret_cast<uintptr_t>(dones) & 3) == 0);
    const ScanPlan p = plan_scan(N, al);
    cudaStream_t s = mlb_stream(stream);
    if (p.vec == 4) returns_kernel<4, 4><<<p.grid, p.block, 0, s>>>(rewards, dones, bootstrap, returns, T, N, gamma);
    else if (p.vec == 2) returns_kernel<2, 8><<<p.grid, p.block, 0, s>>>(rewards, dones, bootstrap, returns, T, N, gamma);
    else if (p.deep) returns_kernel<1, 16><<<p.grid, p.block, 0, s>>>(rewards, dones, bootstrap, returns, T, N, gamma);
    else returns_kernel<1, 8><<<p.grid, p.block, 0, s>>>(rewards, dones, bootstrap, returns, T, N, gamma);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
