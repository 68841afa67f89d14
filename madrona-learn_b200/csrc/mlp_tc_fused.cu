// Fused tensor-core MLP layer kernels (tcgen05 / TMEM / TMA), compute_dtype=bfloat16.
//
//   mlb_dense_ln_relu_fwd_tc   Y = relu(LayerNorm(X W) * scale + bias)
//       replaces nn.Dense -> nn.LayerNorm(eps 1e-6, fast variance) -> relu (ml/models.py:107-117)
//   mlb_dense_dx_lnbwd_tc      dZ_prev = LayerNormReLU'( dZ W^T )   (+ dscale_prev, dbias_prev)
//       replaces the autodiff of the same three ops for the previous layer (ml/ppo.py:276-281)
//
// Both keep a whole 128-row x H accumulator tile in TMEM (H <= 512 = all 512 columns), so the
// LayerNorm row statistics, the ReLU mask and the LayerNorm backward -- which need a full row --
// run in the epilogue straight out of TMEM: the pre-activation Z and the back-propagated dY are
// never written to HBM.  Epilogue thread t of warp w owns accumulator row 32*(w%4)+t
// (tcgen05.ld 32x32b), so row reductions are thread-local; the per-feature sums of the
// backward (dscale, dbias) use a 31-shuffle warp reduce-scatter per 32-column chunk.
//
// Forward (training) stores, per element, y (bf16) and xhat (bf16), plus rstd per row: that is
// what the backward needs (xhat for the LN Jacobian, sign(xhat*scale+bias) for the ReLU mask).
#include <stdlib.h>

#include "tc_common.cuh"
#include "mlp_tc_persist.cuh"

namespace {

using namespace tc;

constexpr int FUSED_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 column halves x 4 lane quadrants)
constexpr float LN_EPS = 1e-6f;

template <int STAGES>
struct FusedSmem {
    // stage = A tile (16 KB) + B tile (HN * 128 B)
    static __host__ __device__ constexpr int stage_bytes(int hn) { return BM * BK * 2 + hn * BK * 2; }
    static __host__ __device__ constexpr int bar_off(int hn) { return STAGES * stage_bytes(hn); }
    // + barriers (2*STAGES+1) + tmem slot + scale/bias (2*hn floats) + colsum (2*hn floats) + slack
    static __host__ __device__ constexpr int total(int hn) {
        return bar_off(hn) + (2 * STAGES + 2) * 8 + 16 + (4 * hn + 512) * 4 + 1024;
    }
};

// barrier of one 128-thread epilogue half-group (compile-time ids keep the barrier count low)
__device__ __forceinline__ void group_bar(int half) {
    if (half == 0) named_bar_sync(1, 128); else named_bar_sync(2, 128);
}

// 32-value warp reduce-scatter: on return lane i holds sum over the 32 lanes of v[i].
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int b = 16; b >= 1; b >>= 1) {
        const bool up = (lane & b) != 0;
#pragma unroll
        for (int j = 0; j < b; ++j) {
            const float send = up ? v[j] : v[j + b];
            const float keep = up ? v[j + b] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, b);
        }
    }
    return v[0];
}

// Shared mainloop: TMA producer (warp 0) + MMA issuer (warp 1); accumulator [128 x HN] in TMEM.
// A K-major [M rows, K]; B K-major [HN rows, K] loaded as HN/256 (or 1) boxes.
template <int STAGES>
__device__ __forceinline__ void mainloop(const CUtensorMap* tmA, const CUtensorMap* tmB, uint8_t* smem,
                                         uint64_t* full_bar, uint64_t* empty_bar, uint64_t* acc_bar,
                                         uint32_t tmem_base, int warp, int lane, int m0, int K, int HN) {
    const int num_kb = (K + BK - 1) / BK;
    const int stage_bytes = FusedSmem<STAGES>::stage_bytes(HN);
    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
                uint8_t* sa = smem + s * stage_bytes;
                uint8_t* sb = sa + BM * BK * 2;
                mbar_expect_tx(&full_bar[s], stage_bytes);
                tma_load_2d(tmA, &full_bar[s], sa, kb * BK, m0);
                for (int n = 0; n < HN; n += 256)            // box rows <= 256
                    tma_load_2d(tmB, &full_bar[s], sb + n * BK * 2, kb * BK, n);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const int n_inst = HN > 256 ? 256 : HN;          // N of one tcgen05.mma
            const uint32_t idesc = umma_idesc(false, false, n_inst);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], (kb / STAGES) & 1);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + s * stage_bytes);
                const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t ad = umma_desc(sa + k * 32, 16, 1024);
                    for (int n = 0; n < HN; n += 256) {
                        const uint64_t bd = umma_desc(sb + n * BK * 2 + k * 32, 16, 1024);
                        tcgen05_mma_f16(tmem_base + n, ad, bd, idesc, (kb | k) ? 1u : 0u);
                    }
                }
                tcgen05_commit(&empty_bar[s]);
            }
            tcgen05_commit(acc_bar);
        }
    }
}

struct Prologue {
    uint8_t* smem;
    uint64_t *full_bar, *empty_bar, *acc_bar, *aux_bar;
    float* fsm;            // 4*HN + 512 floats: scale | bias | colsum0 | colsum1 | row partials [2][128][2]
    uint32_t tmem_base;
};

template <int STAGES>
__device__ __forceinline__ Prologue prologue(uint8_t* smem_raw, const CUtensorMap* tmA,
                                             const CUtensorMap* tmB, int HN, uint32_t tmem_cols,
                                             const float* scale, const float* bias) {
    Prologue p;
    p.smem = align_smem_1024(smem_raw);
    p.full_bar = reinterpret_cast<uint64_t*>(p.smem + FusedSmem<STAGES>::bar_off(HN));
    p.empty_bar = p.full_bar + STAGES;
    p.acc_bar = p.empty_bar + STAGES;
    p.aux_bar = p.acc_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p.aux_bar + 1);
    p.fsm = reinterpret_cast<float*>(tmem_slot + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&p.full_bar[s], 1); mbar_init(&p.empty_bar[s], 1); }
        mbar_init(p.acc_bar, 1);
        mbar_init(p.aux_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < HN; i += blockDim.x) {
        p.fsm[i] = scale[i];
        p.fsm[HN + i] = bias[i];
        p.fsm[2 * HN + i] = 0.f;
        p.fsm[3 * HN + i] = 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    p.tmem_base = *tmem_slot;
    return p;
}

__device__ __forceinline__ void epilogue_done(uint32_t tmem_base, uint32_t tmem_cols, int warp) {
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                     : "memory");
    }
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        uint4 o;
        o.x = pack_bf16(v[j], v[j + 1]); o.y = pack_bf16(v[j + 2], v[j + 3]);
        o.z = pack_bf16(v[j + 4], v[j + 5]); o.w = pack_bf16(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(dst + j) = o;
    }
}

__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + j));
        v[j] = bf16lo(u.x); v[j + 1] = bf16hi(u.x); v[j + 2] = bf16lo(u.y); v[j + 3] = bf16hi(u.y);
        v[j + 4] = bf16lo(u.z); v[j + 5] = bf16hi(u.z); v[j + 6] = bf16lo(u.w); v[j + 7] = bf16hi(u.w);
    }
}

// ------------------------------------------------------------------------------------------
// forward: Y = relu(LN(X W))
// ------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(FUSED_THREADS, 2)
dense_ln_relu_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmXH,
                         const float* __restrict__ scale, const float* __restrict__ bias,
                         int has_xh, float* __restrict__ rstd_out, int M, int K, int HN,
                         uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    Prologue p = prologue<STAGES>(smem_raw, &tmA, &tmB, HN, tmem_cols, scale, bias);
    if (warp < 2) {
        mainloop<STAGES>(&tmA, &tmB, p.smem, p.full_bar, p.empty_bar, p.acc_bar, p.tmem_base, warp, lane, m0, K, HN);
    } else {
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                // column half handled by this warp
        const int rt = quad * 32 + lane;                 // row inside the tile = TMEM lane
        const int row = m0 + rt;
        const int et = (((warp - 2) & 3) << 5) | lane;   // thread id inside the half-group, 0..127
        const uint32_t taddr = p.tmem_base + ((uint32_t)(quad * 32) << 16);
        float* part = p.fsm + 4 * HN;                    // [2][128][2] row partials
        const int nchunks = HN / 32, num_panels = HN / 64;
        mbar_wait(p.acc_bar, 0);                         // all MMAs retired: operand stages are free
        tcgen05_fence_after();
        // pass 1: row statistics (fp32) over this half's columns, combined through shared memory;
        // fast variance like flax: var = max(0, E[z^2] - E[z]^2)
        float sum = 0.f, sq = 0.f;
        for (int ch = half * nchunks / 2; ch < (half + 1) * nchunks / 2; ++ch) {
            uint32_t r[32];
            tmem_ld32(taddr + ch * 32, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float z = __uint_as_float(r[j]); sum += z; sq = fmaf(z, z, sq); }
        }
        part[(half * 128 + rt) * 2] = sum;
        part[(half * 128 + rt) * 2 + 1] = sq;
        named_bar_sync(3, 256);
        sum += part[((half ^ 1) * 128 + rt) * 2];
        sq += part[((half ^ 1) * 128 + rt) * 2 + 1];
        const float invH = 1.f / (float)HN;
        const float mean = sum * invH;
        const float var = fmaxf(0.f, sq * invH - mean * mean);
        const float rstd = rsqrtf(var + LN_EPS);
        if (half == 0 && row < M && rstd_out) rstd_out[row] = rstd;
        // pass 2: normalise, scale/bias, ReLU -> bf16 into SWIZZLE_128B panels in the (now free)
        // operand stages -> TMA store.  One 32 KB panel buffer (Y | XH) per half-group.
        const float* s = p.fsm;
        const float* b = p.fsm + HN;
        uint8_t* buf = p.smem + half * 32768;
        const int p_lo = num_panels >= 2 ? half * num_panels / 2 : 0;
        const int p_hi = num_panels >= 2 ? (half + 1) * num_panels / 2 : (half == 0 ? 1 : 0);
        for (int pnl = p_lo; pnl < p_hi; ++pnl) {
            if (pnl > p_lo) {                             // buffer reuse: its previous store must have read it
                if (et == 0) tma_store_wait_read<0>();
                group_bar(half);
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int c = pnl * 64 + hf * 32;
                uint32_t r[32];
                tmem_ld32(taddr + c, r);
                float xh[32], y[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    xh[j] = (__uint_as_float(r[j]) - mean) * rstd;
                    y[j] = fmaxf(0.f, fmaf(xh[j], s[c + j], b[c + j]));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 o;
                    o.x = pack_bf16(y[8 * q], y[8 * q + 1]); o.y = pack_bf16(y[8 * q + 2], y[8 * q + 3]);
                    o.z = pack_bf16(y[8 * q + 4], y[8 * q + 5]); o.w = pack_bf16(y[8 * q + 6], y[8 * q + 7]);
                    *reinterpret_cast<uint4*>(buf + sw128(rt, hf * 4 + q)) = o;
                    if (has_xh) {
                        o.x = pack_bf16(xh[8 * q], xh[8 * q + 1]); o.y = pack_bf16(xh[8 * q + 2], xh[8 * q + 3]);
                        o.z = pack_bf16(xh[8 * q + 4], xh[8 * q + 5]); o.w = pack_bf16(xh[8 * q + 6], xh[8 * q + 7]);
                        *reinterpret_cast<uint4*>(buf + 16384 + sw128(rt, hf * 4 + q)) = o;
                    }
                }
            }
            fence_async_smem();
            group_bar(half);
            if (et == 0) {
                tma_store_2d(&tmY, buf, pnl * 64, m0);    // rows >= M are clipped by the tensor map
                if (has_xh) tma_store_2d(&tmXH, buf + 16384, pnl * 64, m0);
                tma_store_commit();
            }
        }
        if (et == 0) tma_store_wait<0>();
    }
    epilogue_done(p.tmem_base, tmem_cols, warp);
}

// ------------------------------------------------------------------------------------------
// backward: dZ_prev = LN'/ReLU'(dY),  dY = dZ W^T  (accumulator), per-feature dscale / dbias
// ------------------------------------------------------------------------------------------
template <int STAGES, int MINB>
__global__ void __launch_bounds__(FUSED_THREADS, MINB)
dense_dx_lnbwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmXH, const __grid_constant__ CUtensorMap tmDZ,
                      const float* __restrict__ scale, const float* __restrict__ bias,
                      const float* __restrict__ rstd_in, float* __restrict__ dscale,
                      float* __restrict__ dbias, int M, int K, int HN, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    Prologue p = prologue<STAGES>(smem_raw, &tmA, &tmB, HN, tmem_cols, scale, bias);
    if (warp < 2) {
        mainloop<STAGES>(&tmA, &tmB, p.smem, p.full_bar, p.empty_bar, p.acc_bar, p.tmem_base, warp, lane, m0, K, HN);
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int rt = quad * 32 + lane;
        const int row = m0 + rt;
        const int et256 = threadIdx.x - 64;              // 0..255 over both half-groups
        const bool valid = row < M;
        const uint32_t taddr = p.tmem_base + ((uint32_t)(quad * 32) << 16);
        const float* s = p.fsm;
        const float* b = p.fsm + HN;
        float* cs = p.fsm + 2 * HN;          // per-CTA dscale partial
        float* cb = p.fsm + 3 * HN;          // per-CTA dbias partial
        float* part = p.fsm + 4 * HN;        // [2][128][2] row partials
        const float rstd = valid ? rstd_in[row] : 0.f;
        const int nchunks = HN / 32, num_panels = HN / 64;
        const int ch_lo = half * nchunks / 2, ch_hi = (half + 1) * nchunks / 2;
        mbar_wait(p.acc_bar, 0);             // MMAs retired: operand stages are free
        tcgen05_fence_after();
        // xhat tile [128 x HN] bf16 -> SWIZZLE_128B panels in the freed operand stages (TMA load,
        // rows >= M zero-filled); dz is later written in place and TMA-stored from the same panels
        if (et256 == 0) {
            mbar_expect_tx(p.aux_bar, (uint32_t)num_panels * 16384u);
            for (int pnl = 0; pnl < num_panels; ++pnl)
                tma_load_2d(&tmXH, p.aux_bar, p.smem + pnl * 16384, pnl * 64, m0);
        }
        mbar_wait(p.aux_bar, 0);
        // pass 1: m1 = mean(dxhat), m2 = mean(dxhat * xhat); per-feature sums of du*xhat and du by
        // warp reduce-scatter.  (Writing dxhat back to TMEM with tcgen05.st to shorten pass 2 was
        // measured slower: the extra live registers spill at 2 CTAs/SM.)
        float m1 = 0.f, m2 = 0.f;
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
            const int c = ch * 32;
            uint32_t r[32];
            tmem_ld32(taddr + c, r);
            const uint8_t* pan = p.smem + (c >> 6) * 16384;
            const int hf = ch & 1;
            float gx[32], g[32];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 u = *reinterpret_cast<const uint4*>(pan + sw128(rt, hf * 4 + q));
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
                const float4 sa = *reinterpret_cast<const float4*>(s + c + 8 * q);
                const float4 sb = *reinterpret_cast<const float4*>(s + c + 8 * q + 4);
                const float4 ba = *reinterpret_cast<const float4*>(b + c + 8 * q);
                const float4 bb = *reinterpret_cast<const float4*>(b + c + 8 * q + 4);
                const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = 8 * q + e;
                    const float xh = (e & 1) ? bf16hi(w4[e >> 1]) : bf16lo(w4[e >> 1]);
                    const float dy = __uint_as_float(r[j]);
                    const float du = (fmaf(xh, sv[e], bv[e]) > 0.f) ? dy : 0.f;       // ReLU mask
                    const float dxh = du * sv[e];
                    m1 += dxh;
                    m2 = fmaf(dxh, xh, m2);
                    gx[j] = du * xh;
                    g[j] = du;
                }
            }
            const float csum = warp_reduce_scatter32(gx, lane);
            const float bsum = warp_reduce_scatter32(g, lane);
            atomicAdd(&cs[c + lane], csum);
            atomicAdd(&cb[c + lane], bsum);
        }
        part[(half * 128 + rt) * 2] = m1;
        part[(half * 128 + rt) * 2 + 1] = m2;
        named_bar_sync(3, 256);
        m1 += part[((half ^ 1) * 128 + rt) * 2];
        m2 += part[((half ^ 1) * 128 + rt) * 2 + 1];
        const float invH = 1.f / (float)HN;
        const float c1 = rstd * m1 * invH;               // dz = rstd*dxhat - rstd*m1 - xhat*(rstd*m2)
        const float c2 = rstd * m2 * invH;
        // pass 2: dz = rstd*dxhat - rstd*m1 - xhat*(rstd*m2), written over xhat in the panels
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
            const int c = ch * 32;
            uint32_t r[32];
            tmem_ld32(taddr + c, r);
            uint8_t* pan = p.smem + (c >> 6) * 16384;
            const int hf = ch & 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4* slot = reinterpret_cast<uint4*>(pan + sw128(rt, hf * 4 + q));
                const uint4 u = *slot;
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
                const float4 sa = *reinterpret_cast<const float4*>(s + c + 8 * q);
                const float4 sb = *reinterpret_cast<const float4*>(s + c + 8 * q + 4);
                const float4 ba = *reinterpret_cast<const float4*>(b + c + 8 * q);
                const float4 bb = *reinterpret_cast<const float4*>(b + c + 8 * q + 4);
                const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                float dz8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (e & 1) ? bf16hi(w4[e >> 1]) : bf16lo(w4[e >> 1]);
                    const float dy = __uint_as_float(r[8 * q + e]);
                    const float rs = (fmaf(xh, sv[e], bv[e]) > 0.f) ? rstd * sv[e] : 0.f;   // rstd * mask * scale
                    dz8[e] = fmaf(-c2, xh, fmaf(rs, dy, -c1));
                }
                *slot = make_uint4(pack_bf16(dz8[0], dz8[1]), pack_bf16(dz8[2], dz8[3]),
                                   pack_bf16(dz8[4], dz8[5]), pack_bf16(dz8[6], dz8[7]));
            }
        }
        fence_async_smem();
        named_bar_sync(3, 256);
        if (et256 == 0) {
            for (int pnl = 0; pnl < num_panels; ++pnl)
                tma_store_2d(&tmDZ, p.smem + pnl * 16384, pnl * 64, m0);
            tma_store_commit();
        }
        // the eight epilogue warps publish the CTA's per-feature partials (bar above ordered the atomics)
        for (int i = et256; i < HN; i += 256) {
            atomicAdd(dscale + i, cs[i]);
            atomicAdd(dbias + i, cb[i]);
        }
        if (et256 == 0) tma_store_wait<0>();
    }
    epilogue_done(p.tmem_base, tmem_cols, warp);
}

uint32_t tmem_cols_for(int hn) {
    uint32_t c = 32;
    while ((int)c < hn) c <<= 1;
    return c;
}

bool width_ok(int hn) { return hn >= 64 && hn % 64 == 0 && (hn <= 256 || hn == 512); }

}  // namespace

// X bf16 [M, K] (ldx) ; Wt bf16 [HN, K] (= W^T, ldw) ; scale/bias f32 [HN]
// Y bf16 [M, HN] ; XH (may be NULL) bf16 [M, HN] ; rstd (may be NULL) f32 [M]
MLB_API int mlb_dense_ln_relu_fwd_tc(void* stream, const void* X, const void* Wt, const float* scale,
                                     const float* bias, void* Y, void* XH, float* rstd, int M, int K,
                                     int HN, int ldx, int ldw) {
    MLB_REQUIRE(X && Wt && scale && bias && Y && M > 0 && K > 0 && width_ok(HN));
    MLB_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && mlb_aligned16(X) && mlb_aligned16(Wt) && mlb_aligned16(Y) &&
                (XH == nullptr || mlb_aligned16(XH)));
    if (tcp::persist_ok(M, K, HN))
        return tcp::launch_fwd_persist(mlb_stream(stream), X, Wt, scale, bias, Y, XH, rstd, M, K, HN, ldx, ldw);
    if (tcp::stream_ok(M, K, HN))
        return tcp::launch_fwd_stream(mlb_stream(stream), X, Wt, scale, bias, Y, XH, rstd, M, K, HN, ldx, ldw);
    CUtensorMap tA, tB;
    int rc = make_map(&tA, X, K, M, ldx, 64, 128);
    if (rc) return rc;
    rc = make_map(&tB, Wt, K, HN, ldw, 64, HN > 256 ? 256 : HN);
    if (rc) return rc;
    CUtensorMap tY, tXH;
    rc = make_map(&tY, Y, HN, M, HN, 64, 128);
    if (rc) return rc;
    rc = make_map(&tXH, XH ? XH : Y, HN, M, HN, 64, 128);
    if (rc) return rc;
    constexpr int ST = 2;
    const int smem = FusedSmem<ST>::total(HN);
    auto kern = dense_ln_relu_fwd_kernel<ST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<mlb_cdiv(M, BM), FUSED_THREADS, smem, mlb_stream(stream)>>>(
        tA, tB, tY, tXH, scale, bias, XH ? 1 : 0, rstd, M, K, HN, tmem_cols_for(HN));
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

// DZ_in bf16 [M, K] (lda) : gradient w.r.t. this layer's pre-activation (K = this layer's width,
//   or the head width for the last layer) ; W bf16 [HN, K] (ldw): this layer's kernel [in=HN, out=K]
// scale/bias/XH/rstd: the PREVIOUS layer's LayerNorm parameters and stashed forward state
// DZ_out bf16 [M, HN] ; dscale/dbias f32 [HN] accumulated with atomics (pre-zeroed)
MLB_API int mlb_dense_dx_lnbwd_tc(void* stream, const void* DZ_in, const void* W, const float* scale,
                                  const float* bias, const void* XH, const float* rstd, void* DZ_out,
                                  float* dscale, float* dbias, int M, int K, int HN, int lda, int ldw) {
    MLB_REQUIRE(DZ_in && W && scale && bias && XH && rstd && DZ_out && dscale && dbias);
    MLB_REQUIRE(M > 0 && K > 0 && width_ok(HN) && lda % 8 == 0 && ldw % 8 == 0);
    MLB_REQUIRE(mlb_aligned16(DZ_in) && mlb_aligned16(W) && mlb_aligned16(XH) && mlb_aligned16(DZ_out));
    if (tcp::persist_ok(M, K, HN))
        return tcp::launch_dx_persist(mlb_stream(stream), DZ_in, W, scale, bias, XH, rstd, DZ_out, dscale, dbias,
                                      M, K, HN, lda, ldw);
    if (tcp::stream_ok(M, K, HN))
        return tcp::launch_dx_stream(mlb_stream(stream), DZ_in, W, scale, bias, XH, rstd, DZ_out, dscale, dbias,
                                     M, K, HN, lda, ldw);
    CUtensorMap tA, tB;
    int rc = make_map(&tA, DZ_in, K, M, lda, 64, 128);
    if (rc) return rc;
    rc = make_map(&tB, W, K, HN, ldw, 64, HN > 256 ? 256 : HN);
    if (rc) return rc;
    CUtensorMap tXH, tDZ;
    rc = make_map(&tXH, XH, HN, M, HN, 64, 128);
    if (rc) return rc;
    rc = make_map(&tDZ, DZ_out, HN, M, HN, 64, 128);
    if (rc) return rc;
    constexpr int ST = 2;
    const int smem = FusedSmem<ST>::total(HN);
    static const int minb = [] { const char* v = getenv("MLB_DX_MINBLOCKS"); return v ? atoi(v) : 2; }();
    auto kern = minb == 1 ? dense_dx_lnbwd_kernel<ST, 1> : dense_dx_lnbwd_kernel<ST, 2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<mlb_cdiv(M, BM), FUSED_THREADS, smem, mlb_stream(stream)>>>(
        tA, tB, tXH, tDZ, scale, bias, rstd, dscale, dbias, M, K, HN, tmem_cols_for(HN));
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
