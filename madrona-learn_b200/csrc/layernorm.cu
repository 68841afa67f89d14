// LayerNorm + ReLU forward / backward for the fp32 path.
//
// Replaces flax nn.LayerNorm (eps 1e-6, fast variance var = max(0, E[x^2] - E[x]^2), scale and
// bias; ml/models.py:46-56) followed by nn.relu (ml/models.py:117), and their autodiff.
// One warp per row (H <= 1024, H % 4 == 0): 128-bit loads, warp-shuffle row reductions, the
// per-feature dscale/dbias sums stay in registers across the rows a warp visits and are
// reduced through shared memory + one fp32 atomic per feature per block.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr float LN_EPS = 1e-6f;
constexpr int MAXV = 8;           // float4 vectors per lane: H <= 32*4*8 = 1024

// 4-element vector load / store in fp32 or bf16 storage (bf16 path: activations of the
// tensor-core MLP are bf16 in HBM, statistics and arithmetic stay fp32)
__device__ __forceinline__ float4 ld4(const float* p, int c) { return __ldg(reinterpret_cast<const float4*>(p) + c); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p, int c) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p) + c);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, int c, float4 v) { reinterpret_cast<float4*>(p)[c] = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, int c, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(p)[c] = u;
}

template <int NV, typename TY>
__global__ void __launch_bounds__(256)
ln_relu_fwd_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                   const float* __restrict__ bias, TY* __restrict__ y,
                   float* __restrict__ stats, long long rows, int H) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nvec = H / 4;
    float4 s[NV], b[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) { s[i] = __ldg(reinterpret_cast<const float4*>(scale) + c);
                        b[i] = __ldg(reinterpret_cast<const float4*>(bias) + c); }
    }
    const float invH = 1.f / (float)H;
    for (long long r = warp; r < rows; r += nwarps) {
        const float4* zr = reinterpret_cast<const float4*>(z + r * H);
        float4 v[NV];
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                v[i] = __ldg(zr + c);
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
        }
        sum = warp_sum(sum);
        sq = warp_sum(sq);
        const float mean = sum * invH;
        const float var = fmaxf(0.f, sq * invH - mean * mean);
        const float rstd = rsqrtf(var + LN_EPS);
        TY* yr = y + r * H;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                float4 o;
                o.x = fmaxf(0.f, (v[i].x - mean) * rstd * s[i].x + b[i].x);
                o.y = fmaxf(0.f, (v[i].y - mean) * rstd * s[i].y + b[i].y);
                o.z = fmaxf(0.f, (v[i].z - mean) * rstd * s[i].z + b[i].z);
                o.w = fmaxf(0.f, (v[i].w - mean) * rstd * s[i].w + b[i].w);
                st4(yr, c, o);
            }
        }
        if (stats && lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
    }
}

template <int NV, typename TD>
__global__ void __launch_bounds__(256, NV <= 2 ? 2 : 1)
ln_relu_bwd_kernel(const TD* __restrict__ dy, const float* __restrict__ z,
                   const float* __restrict__ stats, const float* __restrict__ scale,
                   const float* __restrict__ bias, TD* __restrict__ dz,
                   float* __restrict__ dscale, float* __restrict__ dbias, long long rows, int H) {
    extern __shared__ float sm[];           // [2][H] per-block feature sums
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nvec = H / 4;
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    float4 s[NV], b[NV], gs[NV], gb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        gs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        gb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nvec) { s[i] = __ldg(reinterpret_cast<const float4*>(scale) + c);
                        b[i] = __ldg(reinterpret_cast<const float4*>(bias) + c); }
    }
    const float invH = 1.f / (float)H;
    // Software-pipelined over the rows a warp visits: the loads of row r + nwarps are issued before the
    // reductions of row r (88 registers allow two 256-thread blocks per SM; one row per warp in flight left the
    // kernel latency-bound at 3.5 TB/s, two rows double the bytes in flight).
    float4 zc[NV], dc[NV];
    float2 stc = make_float2(0.f, 1.f);
    auto load_row = [&](long long r, float4 (&zz)[NV], float4 (&dd)[NV], float2& st) {
        const float4* zr = reinterpret_cast<const float4*>(z + r * H);
        const TD* dr = dy + r * H;
        st = __ldg(reinterpret_cast<const float2*>(stats) + r);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) { zz[i] = __ldg(zr + c); dd[i] = ld4(dr, c); }
        }
    };
    long long r = warp;
    if (r < rows) load_row(r, zc, dc, stc);
    while (r < rows) {
        const long long rn = r + nwarps;
        float4 zn[NV], dn[NV];
        float2 stn = make_float2(0.f, 1.f);
        constexpr bool PIPE = NV <= 2;             // wider rows: the second row in flight would spill
        if (PIPE && rn < rows) load_row(rn, zn, dn, stn);
        const float mean = stc.x, rstd = stc.y;
        float4 xh[NV], dx[NV];
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                const float4 zv = zc[i], dv = dc[i];
                float4 x, g;
                x.x = (zv.x - mean) * rstd; x.y = (zv.y - mean) * rstd;
                x.z = (zv.z - mean) * rstd; x.w = (zv.w - mean) * rstd;
                g.x = (x.x * s[i].x + b[i].x > 0.f) ? dv.x : 0.f;
                g.y = (x.y * s[i].y + b[i].y > 0.f) ? dv.y : 0.f;
                g.z = (x.z * s[i].z + b[i].z > 0.f) ? dv.z : 0.f;
                g.w = (x.w * s[i].w + b[i].w > 0.f) ? dv.w : 0.f;
                gs[i].x += g.x * x.x; gs[i].y += g.y * x.y; gs[i].z += g.z * x.z; gs[i].w += g.w * x.w;
                gb[i].x += g.x; gb[i].y += g.y; gb[i].z += g.z; gb[i].w += g.w;
                g.x *= s[i].x; g.y *= s[i].y; g.z *= s[i].z; g.w *= s[i].w;
                m1 += (g.x + g.y) + (g.z + g.w);
                m2 += (g.x * x.x + g.y * x.y) + (g.z * x.z + g.w * x.w);
                xh[i] = x; dx[i] = g;
            }
        }
        m1 = warp_sum(m1) * invH;
        m2 = warp_sum(m2) * invH;
        TD* or_ = dz + r * H;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                float4 o;
                o.x = rstd * (dx[i].x - m1 - xh[i].x * m2);
                o.y = rstd * (dx[i].y - m1 - xh[i].y * m2);
                o.z = rstd * (dx[i].z - m1 - xh[i].z * m2);
                o.w = rstd * (dx[i].w - m1 - xh[i].w * m2);
                st4(or_, c, o);
            }
        }
        if (PIPE) {
#pragma unroll
            for (int i = 0; i < NV; ++i) { zc[i] = zn[i]; dc[i] = dn[i]; }
            stc = stn;
        } else if (rn < rows) {
            load_row(rn, zc, dc, stc);
        }
        r = rn;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) {
            atomicAdd(&sm[4 * c + 0], gs[i].x); atomicAdd(&sm[4 * c + 1], gs[i].y);
            atomicAdd(&sm[4 * c + 2], gs[i].z); atomicAdd(&sm[4 * c + 3], gs[i].w);
            atomicAdd(&sm[H + 4 * c + 0], gb[i].x); atomicAdd(&sm[H + 4 * c + 1], gb[i].y);
            atomicAdd(&sm[H + 4 * c + 2], gb[i].z); atomicAdd(&sm[H + 4 * c + 3], gb[i].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        atomicAdd(dscale + i, sm[i]);
        atomicAdd(dbias + i, sm[H + i]);
    }
}

unsigned row_grid(long long rows) {
    long long b = (rows + 7) / 8;           // 8 warps per block
    const long long cap = (long long)MLB_NUM_SMS * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

template <typename TY>
static int ln_fwd_impl(void* stream, const float* z, const float* scale, const float* bias, TY* y,
                       float* stats, long long rows, int H) {
    MLB_REQUIRE(z && scale && bias && y && rows >= 0 && H > 0 && H % 4 == 0 && H <= 128 * MAXV);
    if (rows == 0) return MLB_OK;
    if (!(mlb_aligned16(z) && mlb_aligned16(y) && mlb_aligned16(scale) && mlb_aligned16(bias))) return MLB_EALIGN;
    cudaStream_t s = mlb_stream(stream);
    const unsigned g = row_grid(rows);
    const int nv = (H / 4 + 31) / 32;
    if (nv <= 1) ln_relu_fwd_kernel<1, TY><<<g, 256, 0, s>>>(z, scale, bias, y, stats, rows, H);
    else if (nv <= 2) ln_relu_fwd_kernel<2, TY><<<g, 256, 0, s>>>(z, scale, bias, y, stats, rows, H);
    else if (nv <= 4) ln_relu_fwd_kernel<4, TY><<<g, 256, 0, s>>>(z, scale, bias, y, stats, rows, H);
    else ln_relu_fwd_kernel<8, TY><<<g, 256, 0, s>>>(z, scale, bias, y, stats, rows, H);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_ln_relu_fwd_f32(void* stream, const float* z, const float* scale,
                                const float* bias, float* y, float* stats, long long rows, int H) {
    return ln_fwd_impl<float>(stream, z, scale, bias, y, stats, rows, H);
}

MLB_API int mlb_ln_relu_fwd_bf16(void* stream, const float* z, const float* scale,
                                 const float* bias, void* y_bf16, float* stats, long long rows, int H) {
    return ln_fwd_impl<__nv_bfloat16>(stream, z, scale, bias, reinterpret_cast<__nv_bfloat16*>(y_bf16),
                                      stats, rows, H);
}

template <typename TD>
static int ln_bwd_impl(void* stream, const TD* dy, const float* z, const float* stats,
                       const float* scale, const float* bias, TD* dz, float* dscale,
                       float* dbias, long long rows, int H) {
    MLB_REQUIRE(dy && z && stats && scale && bias && dz && dscale && dbias);
    MLB_REQUIRE(rows >= 0 && H > 0 && H % 4 == 0 && H <= 128 * MAXV);
    if (rows == 0) return MLB_OK;
    if (!(mlb_aligned16(z) && mlb_aligned16(dy) && mlb_aligned16(dz) && mlb_aligned16(scale) && mlb_aligned16(bias))) return MLB_EALIGN;
    cudaStream_t s = mlb_stream(stream);
    const unsigned g = row_grid(rows);
    const size_t smem = 2 * (size_t)H * sizeof(float);
    const int nv = (H / 4 + 31) / 32;
    if (nv <= 1) ln_relu_bwd_kernel<1, TD><<<g, 256, smem, s>>>(dy, z, stats, scale, bias, dz, dscale, dbias, rows, H);
    else if (nv <= 2) ln_relu_bwd_kernel<2, TD><<<g, 256, smem, s>>>(dy, z, stats, scale, bias, dz, dscale, dbias, rows, H);
    else if (nv <= 4) ln_relu_bwd_kernel<4, TD><<<g, 256, smem, s>>>(dy, z, stats, scale, bias, dz, dscale, dbias, rows, H);
    else ln_relu_bwd_kernel<8, TD><<<g, 256, smem, s>>>(dy, z, stats, scale, bias, dz, dscale, dbias, rows, H);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_ln_relu_bwd_f32(void* stream, const float* dy, const float* z, const float* stats,
                                const float* scale, const float* bias, float* dz, float* dscale,
                                float* dbias, long long rows, int H) {
    return ln_bwd_impl<float>(stream, dy, z, stats, scale, bias, dz, dscale, dbias, rows, H);
}

MLB_API int mlb_ln_relu_bwd_bf16(void* stream, const void* dy_bf16, const float* z, const float* stats,
                                 const float* scale, const float* bias, void* dz_bf16, float* dscale,
                                 float* dbias, long long rows, int H) {
    return ln_bwd_impl<__nv_bfloat16>(stream, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), z, stats,
                                      scale, bias, reinterpret_cast<__nv_bfloat16*>(dz_bf16), dscale, dbias,
                                      rows, H);
}
