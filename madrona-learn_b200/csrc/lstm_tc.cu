// Fused LSTM step on the tensor cores (ml/rnn.py:10-45 MultiLayerLSTMCell / flax
// OptimizedLSTMCell, :91-96 the end-of-step reset of LSTM.sequence).
//
//     z = [x_t | h_{t-1}] [W_i | W_h]^T          one tcgen05 GEMM, K = in + RH, fp32 accumulators in TMEM
//     i, f, g, o = sigma(z_i + b), sigma(z_f + b), tanh(z_g + b), sigma(z_o + b)      \
//     c' = f c + i g ;  h' = o tanh(c')                                                 } epilogue, straight
//     h_seq = h' ;  carry = (c', h') zeroed where the step ended an episode            /  out of TMEM
//
// The gate pre-activations never touch HBM (the unfused path wrote z [M, 4 RH] fp32 once, added the
// recurrent product to it with atomics and read it back in a cell kernel: three launches per step).
//
// Layout: the weight rows are PERMUTED so that one 256-column accumulator unit holds all four gates of 64
// hidden units: packed row  nb*256 + g*64 + j  <->  gate g, hidden unit nb*64 + j  (mlb_lstm_pack_weights_bf16).
// A persistent CTA walks (row tile, unit) pairs; every ring stage carries one 16 KB activation k-block (from
// x_t for k < in, from h_{t-1} after) and the matching 32 KB k-block of the unit's weights (streamed from
// L2: the packed matrix is 4 RH x (in + RH) bf16 = 1 MB at RH = in = 256).  Two 256-column accumulator
// buffers: the tensor core fills unit u+1 while the sixteen epilogue warps run the cell math on unit u.
// Epilogue warp (quadrant q, group g): rows 32q.., hidden units 16g..16g+15 of the unit, one row per lane in
// registers; c, h and the gate stash move through a per-warp shared-memory transpose tile so that global
// accesses are 64-byte row segments (8 rows per instruction) instead of 32 scattered sectors.
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int L_THREADS = 576;           // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue
constexpr int L_STAGE = 16384 + 32768;
constexpr int L_MAX_STAGES = 4;
constexpr int L_STAGES = 3;              // 3 x 48 KB ring + 32 KB transpose tiles

struct LBars {
    uint64_t *full, *empty, *acc_full, *acc_empty;
    uint32_t* tmem_slot;
};

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.f / (1.f + __expf(-2.f * x)) - 1.f; }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 accumulator columns of this lane's row
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
}

// Per-warp transpose tile [32 rows x 64 B]: 16-byte slot q of row r at r*64 + ((q ^ ((r >> 1) & 3)) << 4)
// (conflict-free for both the row-per-lane and the 4-lanes-per-row access).
__device__ __forceinline__ uint4* tslot(uint8_t* wt, int r, int q) {
    return reinterpret_cast<uint4*>(wt + r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
}
// row-per-lane registers -> global [rows, ld] fp32, 16 columns: instruction k covers rows 8k..8k+7, four lanes
// per 64-byte row segment
template <bool STREAM>
__device__ __forceinline__ void tile_store16(uint8_t* wt, int lane, float* g, int ld, int rows_valid, const float (&v)[16]) {
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *tslot(wt, lane, q) = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                                         __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
    __syncwarp();
    const int c = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = k * 8 + (lane >> 2);
        const uint4 x = *tslot(wt, r, c);
        if (r < rows_valid) {
            uint4* dst = reinterpret_cast<uint4*>(g + (size_t)r * ld + c * 4);
            if (STREAM) __stcs(dst, x); else *dst = x;
        }
    }
}
// the same for a bf16 destination: 32 bytes per row, two lanes per row, instruction k covers rows 16k..16k+15
__device__ __forceinline__ void tile_store16_bf(uint8_t* wt, int lane, __nv_bfloat16* g, int ld, int rows_valid,
                                                const float (&v)[16]) {
    __syncwarp();
    *tslot(wt, lane, 0) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *tslot(wt, lane, 1) = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
    __syncwarp();
    const int c = lane & 1;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int r = k * 16 + (lane >> 1);
        const uint4 x = *tslot(wt, r, c);
        if (r < rows_valid) *reinterpret_cast<uint4*>(g + (size_t)r * ld + c * 8) = x;
    }
}
// global [rows, ld] fp32, 16 columns -> row-per-lane registers
__device__ __forceinline__ void tile_load16(uint8_t* wt, int lane, const float* g, int ld, int rows_valid, float (&v)[16]) {
    __syncwarp();
    const int c = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = k * 8 + (lane >> 2);
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows_valid) x = *reinterpret_cast<const uint4*>(g + (size_t)r * ld + c * 4);
        *tslot(wt, r, c) = x;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 x = *tslot(wt, lane, q);
        v[4 * q] = __uint_as_float(x.x); v[4 * q + 1] = __uint_as_float(x.y);
        v[4 * q + 2] = __uint_as_float(x.z); v[4 * q + 3] = __uint_as_float(x.w);
    }
}

__global__ void __launch_bounds__(L_THREADS, 1)
lstm_step_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                    const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias_packed,
                    const float* __restrict__ c_prev, const uint8_t* __restrict__ ends,
                    __nv_bfloat16* __restrict__ h_seq, float* __restrict__ c_carry, float* __restrict__ h_carry,
                    __nv_bfloat16* __restrict__ h_carry_bf, float* __restrict__ stash, int M, int in_dim, int RH,
                    int stages) {
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    uint8_t* smem = align_smem_1024(smem_raw);
    uint8_t* tiles = smem + stages * L_STAGE;                                // 16 x 2 KB transpose tiles
    float* bsm = reinterpret_cast<float*>(tiles + 16 * 2048);                // packed bias [4 RH]
    LBars bars;
    bars.full = reinterpret_cast<uint64_t*>(bsm + 4 * RH);
    bars.empty = bars.full + L_MAX_STAGES;
    bars.acc_full = bars.empty + L_MAX_STAGES;
    bars.acc_empty = bars.acc_full + 2;
    bars.tmem_slot = reinterpret_cast<uint32_t*>(bars.acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_x = in_dim / BK, num_kb = kb_x + RH / BK;
    const int nblk = RH / 64;                                                // 256-column units per row tile
    const int num_units = ((M + BM - 1) / BM) * nblk;
    const int my_units = ((int)blockIdx.x < num_units) ? (num_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto issue = [&](int it) {
        const int s = it % stages;
        const int u = blockIdx.x + (it / num_kb) * gridDim.x, kb = it % num_kb;
        const int tile = u / nblk, nb = u % nblk;
        uint8_t* dst = smem + s * L_STAGE;
        mbar_expect_tx(&bars.full[s], (uint32_t)L_STAGE);
        tma_load_2d(&tmW, &bars.full[s], dst + 16384, kb * BK, nb * 256);
        if (kb < kb_x) tma_load_2d(&tmX, &bars.full[s], dst, kb * BK, tile * BM);
        else tma_load_2d(&tmH, &bars.full[s], dst, (kb - kb_x) * BK, tile * BM);
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmH)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        for (int s = 0; s < L_MAX_STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars.acc_full[s], 1); mbar_init(&bars.acc_empty[s], 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(bars.tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();                                              // x_t, h_{t-1}, c, weights and bias come from predecessors
    for (int i = threadIdx.x; i < 4 * RH; i += blockDim.x) bsm[i] = bias_packed[i];
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *bars.tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int total = my_units * num_kb;
            for (int it = 0; it < total; ++it) {
                const int s = it % stages;
                if (it >= stages) mbar_wait_spin(&bars.empty[s], ((it / stages) & 1) ^ 1, 0x26);
                issue(it);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(false, false, 256);
            int it = 0;
            for (int i = 0; i < my_units; ++i) {
                const int buf = i & 1;
                mbar_wait_spin(&bars.acc_empty[buf], ((i >> 1) & 1) ^ 1, 0x25);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    mbar_wait_spin(&bars.full[s], (it / stages) & 1, 0x27);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + s * L_STAGE), sb = sa + 16384;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        tcgen05_mma_f16(d_tmem, umma_desc(sa + k * 32, 16, 1024), umma_desc(sb + k * 32, 16, 1024),
                                        idesc, (kb | k) ? 1u : 0u);
                    tcgen05_commit(&bars.empty[s]);
                }
                tcgen05_commit(&bars.acc_full[buf]);
            }
        }
    } else {
        // Epilogue warp (quad, grp): rows 32 quad.., hidden units 16 grp.. of the unit -- one accumulator row per
        // lane.  Every global array is row-major [M, RH], so a lane's 64 bytes sit 4 RH bytes from its
        // neighbour's: all loads and stores go through a per-warp [32 rows x 64 B] transpose tile so that one
        // instruction covers 8 rows x 64 contiguous bytes (4x fewer LSU wavefronts than row-per-lane access).
        const int quad = warp & 3, grp = (warp - 2) >> 2;
        uint8_t* wt = tiles + (warp - 2) * 2048;
        for (int i = 0; i < my_units; ++i) {
            const int buf = i & 1;
            const int u = blockIdx.x + i * gridDim.x;
            const int tile = u / nblk, nb = u % nblk;
            const int row0 = tile * BM + quad * 32;
            const int rows_valid = M - row0;                  // rows of this warp's block inside M
            const int j0 = nb * 64 + grp * 16;                // first hidden unit of this warp's group
            const bool ok = lane < rows_valid;
            const float keep = (ok && ends && ends[row0 + lane]) ? 0.f : 1.f;
            const uint32_t tq = tmem_base + buf * 256 + ((uint32_t)(quad * 32) << 16) + grp * 16;
            const float* bp = bsm + nb * 256 + grp * 16;      // packed bias: gate blocks of 64
            const size_t gofs = (size_t)row0 * RH + j0;
            float cp[16], a[16], ig[16];
            tile_load16(wt, lane, c_prev + gofs, RH, rows_valid, cp);          // (overlaps the accumulator wait)
            mbar_wait(&bars.acc_full[buf], (i >> 1) & 1, 0x21);
            tcgen05_fence_after();
            float* sp = stash ? stash + (size_t)row0 * 5 * RH + j0 : nullptr;
            // gate i, gate g  ->  ig = i * g
            tmem_ld16(tq, a);
#pragma unroll
            for (int k = 0; k < 16; ++k) { a[k] = sigm(a[k] + bp[k]); ig[k] = a[k]; }
            if (sp) tile_store16<true>(wt, lane, sp, 5 * RH, rows_valid, a);
            tmem_ld16(tq + 128, a);
#pragma unroll
            for (int k = 0; k < 16; ++k) { a[k] = tanh_fast(a[k] + bp[128 + k]); ig[k] *= a[k]; }
            if (sp) tile_store16<true>(wt, lane, sp + 2 * RH, 5 * RH, rows_valid, a);
            // gate f  ->  c' = f c + i g
            tmem_ld16(tq + 64, a);
#pragma unroll
            for (int k = 0; k < 16; ++k) { a[k] = sigm(a[k] + bp[64 + k]); cp[k] = a[k] * cp[k] + ig[k]; }
            if (sp) tile_store16<true>(wt, lane, sp + RH, 5 * RH, rows_valid, a);
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = keep * cp[k];
            tile_store16<false>(wt, lane, c_carry + gofs, RH, rows_valid, a);
#pragma unroll
            for (int k = 0; k < 16; ++k) ig[k] = tanh_fast(cp[k]);              // ig := tanh(c')
            if (sp) tile_store16<true>(wt, lane, sp + 4 * RH, 5 * RH, rows_valid, ig);
            // gate o  ->  h' = o tanh(c')
            tmem_ld16(tq + 192, a);
            tcgen05_fence_before();                           // last TMEM read of this warp for this unit
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.acc_empty[buf]);
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = sigm(a[k] + bp[192 + k]);
            if (sp) tile_store16<true>(wt, lane, sp + 3 * RH, 5 * RH, rows_valid, a);
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] *= ig[k];                         // h'
            tile_store16_bf(wt, lane, h_seq + gofs, RH, rows_valid, a);
            if (h_carry || h_carry_bf) {
#pragma unroll
                for (int k = 0; k < 16; ++k) a[k] *= keep;
                if (h_carry) tile_store16<false>(wt, lane, h_carry + gofs, RH, rows_valid, a);
                if (h_carry_bf) tile_store16_bf(wt, lane, h_carry_bf + gofs, RH, rows_valid, a);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// packed weights [4 RH, in + RH] bf16 and bias [4 RH] f32 from W_i^T [4 RH, in], W_h^T [4 RH, RH], b [4 RH]
__global__ void __launch_bounds__(256)
lstm_pack_kernel(const float* __restrict__ wi, const float* __restrict__ wh, const float* __restrict__ b,
                 __nv_bfloat16* __restrict__ wp, float* __restrict__ bp, int in_dim, int RH) {
    const int K = in_dim + RH;
    const long long n = (long long)4 * RH * K;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int rp = (int)(t / K), k = (int)(t - (long long)rp * K);
        const int nb = rp >> 8, g = (rp >> 6) & 3, j = rp & 63;
        const int r = g * RH + nb * 64 + j;
        wp[t] = __float2bfloat16_rn(k < in_dim ? wi[(long long)r * in_dim + k] : wh[(long long)r * RH + (k - in_dim)]);
        if (k == 0) bp[rp] = b[r];
    }
}

}  // namespace

MLB_API int mlb_lstm_pack_weights_bf16(void* stream, const float* wi_t, const float* wh_t, const float* bias,
                                       void* w_packed, float* bias_packed, int in_dim, int RH) {
    MLB_REQUIRE(wi_t && wh_t && bias && w_packed && bias_packed && in_dim > 0 && RH > 0 && RH % 64 == 0);
    lstm_pack_kernel<<<MLB_NUM_SMS * 2, 256, 0, mlb_stream(stream)>>>(wi_t, wh_t, bias,
        static_cast<__nv_bfloat16*>(w_packed), bias_packed, in_dim, RH);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_lstm_step_tc(void* stream, const void* x, int ldx, const void* h_prev, const void* w_packed,
                             const float* bias_packed, const float* c_prev, const uint8_t* ends, void* h_seq,
                             float* c_carry, float* h_carry, void* h_carry_bf16, float* stash, long long M,
                             int in_dim, int RH) {
    MLB_REQUIRE(x && h_prev && w_packed && bias_packed && c_prev && h_seq && c_carry && M >= 0);
    MLB_REQUIRE(in_dim > 0 && in_dim % 64 == 0 && RH > 0 && RH % 64 == 0 && RH <= 1024 && ldx % 8 == 0 && ldx >= in_dim);
    MLB_REQUIRE(M < (1ll << 31) && mlb_aligned16(x) && mlb_aligned16(h_prev) && mlb_aligned16(w_packed) &&
                mlb_aligned16(c_prev) && mlb_aligned16(h_seq) && mlb_aligned16(c_carry) &&
                (!h_carry || mlb_aligned16(h_carry)) && (!h_carry_bf16 || mlb_aligned16(h_carry_bf16)) &&
                (!stash || mlb_aligned16(stash)));
    if (M == 0) return MLB_OK;
    CUtensorMap tX, tH, tW;
    int rc;
    if ((rc = make_map(&tX, x, in_dim, M, ldx, 64, 128))) return rc;
    if ((rc = make_map(&tH, h_prev, RH, M, RH, 64, 128))) return rc;
    if ((rc = make_map(&tW, w_packed, in_dim + RH, 4 * RH, in_dim + RH, 64, 256))) return rc;
    const int stages = L_STAGES;
    const int smem = stages * L_STAGE + 16 * 2048 + 4 * RH * 4 + 256 + 1024;
    cudaError_t e = cudaFuncSetAttribute(lstm_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    int sms = MLB_NUM_SMS, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long units = ((M + BM - 1) / BM) * (RH / 64);
    const int grid = units < sms ? (int)units : sms;
    e = launch_pdl(lstm_step_tc_kernel, dim3(grid), dim3(L_THREADS), smem, mlb_stream(stream), tX, tH, tW, bias_packed,
                   c_prev, ends, static_cast<__nv_bfloat16*>(h_seq), c_carry, h_carry,
                   static_cast<__nv_bfloat16*>(h_carry_bf16), stash, (int)M, in_dim, RH, stages);
    if (e != cudaSuccess) return (int)e;
    return MLB_OK;
}
