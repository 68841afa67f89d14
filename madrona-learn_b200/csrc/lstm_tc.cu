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
// Epilogue warp (quadrant q, group g): rows 32q.., hidden units 16g..16g+15 of the unit, one row per lane:
// every global access is a whole 32-byte sector (8 fp32 / 16 bf16 per lane).
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int L_THREADS = 576;           // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue
constexpr int L_STAGE = 16384 + 32768;
constexpr int L_MAX_STAGES = 4;

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
    float* bsm = reinterpret_cast<float*>(smem + stages * L_STAGE);          // packed bias [4 RH]
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
        const int quad = warp & 3, grp = (warp - 2) >> 2;
        for (int i = 0; i < my_units; ++i) {
            const int buf = i & 1;
            const int u = blockIdx.x + i * gridDim.x;
            const int tile = u / nblk, nb = u % nblk;
            const int row = tile * BM + quad * 32 + lane;
            const bool ok = row < M;
            const float keep = (ok && ends && ends[row]) ? 0.f : 1.f;
            const uint32_t tq = tmem_base + buf * 256 + ((uint32_t)(quad * 32) << 16) + grp * 16;
            mbar_wait(&bars.acc_full[buf], (i >> 1) & 1, 0x21);
            tcgen05_fence_after();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                  // 8 hidden units at a time
                const int j0 = nb * 64 + grp * 16 + hf * 8;   // first hidden unit of this lane's group
                uint32_t zi[8], zf[8], zg[8], zo[8];
                tmem_ld8(tq + hf * 8, zi);
                tmem_ld8(tq + 64 + hf * 8, zf);
                tmem_ld8(tq + 128 + hf * 8, zg);
                tmem_ld8(tq + 192 + hf * 8, zo);
                float cp[8];
                if (ok) {
                    const float4 a = *reinterpret_cast<const float4*>(c_prev + (size_t)row * RH + j0);
                    const float4 b = *reinterpret_cast<const float4*>(c_prev + (size_t)row * RH + j0 + 4);
                    cp[0] = a.x; cp[1] = a.y; cp[2] = a.z; cp[3] = a.w; cp[4] = b.x; cp[5] = b.y; cp[6] = b.z; cp[7] = b.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) cp[k] = 0.f;
                }
                tmem_wait_ld();
                if (hf == 1) {                                // last TMEM read of this warp for this unit
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.acc_empty[buf]);
                }
                const float* bp = bsm + nb * 256 + grp * 16 + hf * 8;          // packed bias: gate blocks of 64
                float gi[8], gf[8], gg[8], go[8], tc8[8], c8[8], h8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    gi[k] = sigm(__uint_as_float(zi[k]) + bp[k]);
                    gf[k] = sigm(__uint_as_float(zf[k]) + bp[64 + k]);
                    gg[k] = tanh_fast(__uint_as_float(zg[k]) + bp[128 + k]);
                    go[k] = sigm(__uint_as_float(zo[k]) + bp[192 + k]);
                    c8[k] = gf[k] * cp[k] + gi[k] * gg[k];
                    tc8[k] = tanh_fast(c8[k]);
                    h8[k] = go[k] * tc8[k];
                }
                if (ok) {
                    const size_t o = (size_t)row * RH + j0;
                    *reinterpret_cast<uint4*>(h_seq + o) = make_uint4(pack_bf16(h8[0], h8[1]), pack_bf16(h8[2], h8[3]),
                                                                      pack_bf16(h8[4], h8[5]), pack_bf16(h8[6], h8[7]));
                    *reinterpret_cast<float4*>(c_carry + o) = make_float4(keep * c8[0], keep * c8[1], keep * c8[2], keep * c8[3]);
                    *reinterpret_cast<float4*>(c_carry + o + 4) = make_float4(keep * c8[4], keep * c8[5], keep * c8[6], keep * c8[7]);
                    if (h_carry) {
                        *reinterpret_cast<float4*>(h_carry + o) = make_float4(keep * h8[0], keep * h8[1], keep * h8[2], keep * h8[3]);
                        *reinterpret_cast<float4*>(h_carry + o + 4) = make_float4(keep * h8[4], keep * h8[5], keep * h8[6], keep * h8[7]);
                    }
                    if (h_carry_bf)
                        *reinterpret_cast<uint4*>(h_carry_bf + o) =
                            make_uint4(pack_bf16(keep * h8[0], keep * h8[1]), pack_bf16(keep * h8[2], keep * h8[3]),
                                       pack_bf16(keep * h8[4], keep * h8[5]), pack_bf16(keep * h8[6], keep * h8[7]));
                    if (stash) {
                        float* sp = stash + (size_t)row * 5 * RH + j0;
                        const float* src[5] = {gi, gf, gg, go, tc8};
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            __stcs(reinterpret_cast<float4*>(sp + q * RH), make_float4(src[q][0], src[q][1], src[q][2], src[q][3]));
                            __stcs(reinterpret_cast<float4*>(sp + q * RH + 4), make_float4(src[q][4], src[q][5], src[q][6], src[q][7]));
                        }
                    }
                }
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
    const int stages = L_MAX_STAGES;
    const int smem = stages * L_STAGE + 4 * RH * 4 + 256 + 1024;
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
