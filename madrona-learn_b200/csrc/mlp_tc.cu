// Tensor-core (tcgen05 / TMEM / TMA) bf16 GEMM for the policy/value MLP, compute_dtype=bfloat16.
//
// Replaces the dot_general of nn.Dense and its two autodiff transposes (ml/models.py:110-154,
// jax.value_and_grad at ml/ppo.py:276-281) on the 5th-gen tensor cores:
//
//     C[M, N] (+)= A[M, K] * B[N, K]^T        bf16 operands, fp32 accumulation in TMEM
//
// Operand "major-ness" covers all three products of a Dense layer without transposing
// activations in memory:
//   forward   Z  = X  W        A = X  [rows, in]   K-major     B = W^T [H, in]    K-major
//   dX        dX = dZ W^T      A = dZ [rows, H]    K-major     B = W   [in, H]    K-major
//   dW        dW = X^T dZ      A = X  [rows, in]   MN-major    B = dZ  [rows, H]  MN-major
//                              (reduction over rows; split-K over CTAs, fp32 red.global.add)
//
// Kernel anatomy (one CTA per 128 x BN output tile, 192 threads):
//   warp 0    TMA producer: cp.async.bulk.tensor 2-D boxes (64 x 128 K-major, 64 x 64 MN-major,
//             SWIZZLE_128B) into a STAGES-deep shared-memory ring, mbarrier expect_tx/complete_tx
//   warp 1    allocates TMEM (BN fp32 columns), then one elected lane issues
//             tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per 64-wide k-block and
//             tcgen05.commit's the stage's "empty" barrier / the tile's "accumulator full" barrier
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x32 (thread = one accumulator row), optional bias,
//             then fp32 store / bf16 store / fp32 atomic add (split-K)
// All mbarrier waits are bounded (trap instead of hanging the GPU).
#include "tc_common.cuh"
#include "mlp_tc_persist.cuh"

namespace {

using namespace tc;

constexpr int TC_THREADS = 192;

enum Epi { EPI_F32 = 0, EPI_BF16 = 1, EPI_ATOMIC = 2 };

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;            // 16 KB
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(TC_THREADS)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               void* __restrict__ Cout, const float* __restrict__ bias, int ldc, int M, int N,
               int K, int k_per_split) {
    using L = SmemLayout<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem_1024(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);
    const int num_kb = (k_end - k_begin + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {            // TMEM allocation: BN fp32 accumulator columns (power of 2 >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    pdl_wait();                    // operands / the accumulated output come from preceding kernels
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0 && num_kb > 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait_spin(&empty_bar[s], ph ^ 1);
                uint8_t* sa = smem + s * L::STAGE_BYTES;
                uint8_t* sb = sa + L::A_BYTES;
                mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
                const int k0 = k_begin + kb * BK;
                if (A_MN) {         // A stored [K, M]: two 64(m) x 64(k) boxes
                    tma_load_2d(&tmA, &full_bar[s], sa, m0, k0);
                    tma_load_2d(&tmA, &full_bar[s], sa + 8192, m0 + 64, k0);
                } else {            // A stored [M, K]: one 64(k) x 128(m) box
                    tma_load_2d(&tmA, &full_bar[s], sa, k0, m0);
                }
                if (B_MN) {         // B stored [K, N]: BN/64 boxes of 64(n) x 64(k)
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_2d(&tmB, &full_bar[s], sb + j * 8192, n0 + 64 * j, k0);
                } else {            // B stored [N, K]: one 64(k) x BN(n) box
                    tma_load_2d(&tmB, &full_bar[s], sb, k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0 && num_kb > 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1,
            // b=BF16 [10,13)=1, a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait_spin(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
                const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // K-major: 8-row groups 1024 B apart (SBO), K slice = +32 B inside the swizzle row
                    // MN-major: 64-wide MN atoms 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO),
                    //           K slice = 16 k-rows = +2048 B
                    const uint64_t ad = A_MN ? umma_desc(sa + k * 2048, 8192, 1024) : umma_desc(sa + k * 32, 16, 1024);
                    const uint64_t bd = B_MN ? umma_desc(sb + k * 2048, 8192, 1024) : umma_desc(sb + k * 32, 16, 1024);
                    tcgen05_mma_f16(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
                }
                tcgen05_commit(&empty_bar[s]);          // frees the smem stage when the MMAs retire
            }
            tcgen05_commit(acc_bar);                    // accumulator complete
        }
    } else {
        // ================= epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =================
        const int quad = warp & 3;
        const int row = m0 + quad * 32 + lane;
        if (num_kb > 0) {
            mbar_wait(acc_bar, 0);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
                const int col0 = n0 + c * 32;
                if (row < M) {
                    if (EPI == EPI_F32) {
                        float* dst = reinterpret_cast<float*>(Cout) + (long long)row * ldc + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < N) {
                                float4 v;
                                v.x = __uint_as_float(r[j]) + (bias ? bias[col0 + j] : 0.f);
                                v.y = __uint_as_float(r[j + 1]) + (bias ? bias[col0 + j + 1] : 0.f);
                                v.z = __uint_as_float(r[j + 2]) + (bias ? bias[col0 + j + 2] : 0.f);
                                v.w = __uint_as_float(r[j + 3]) + (bias ? bias[col0 + j + 3] : 0.f);
                                *reinterpret_cast<float4*>(dst + j) = v;
                            }
                        }
                    } else if (EPI == EPI_BF16) {
                        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(Cout) + (long long)row * ldc + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (col0 + j < N) {
                                uint4 v;
                                __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                                __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                                __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
                                __nv_bfloat162 p3 = __floats2bfloat162_rn(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
                                v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
                                v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
                                *reinterpret_cast<uint4*>(dst + j) = v;
                            }
                        }
                    } else {
                        // split-K reduction in L2: 128-bit vector reductions (4x fewer atomic ops)
                        float* dst = reinterpret_cast<float*>(Cout) + (long long)row * ldc + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (col0 + j < N)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                             ::"l"(dst + j), "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])),
                                               "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                                             : "memory");
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN)
                     : "memory");
    }
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
int launch_tc(cudaStream_t s, const CUtensorMap& tA, const CUtensorMap& tB, void* C, const float* bias,
              int ldc, int M, int N, int K, int splitk) {
    using L = SmemLayout<BN, STAGES>;
    auto kern = tc_gemm_kernel<BN, STAGES, A_MN, B_MN, EPI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return (int)e;
    int kps = (K + splitk - 1) / splitk;
    kps = (kps + BK - 1) / BK * BK;
    const int zs = (K + kps - 1) / kps;
    dim3 grid(mlb_cdiv(M, BM), mlb_cdiv(N, BN), zs);
    e = launch_pdl(kern, grid, dim3(TC_THREADS), L::TOTAL, s, tA, tB, C, bias, ldc, M, N, K, kps);
    return e == cudaSuccess ? MLB_OK : (int)e;
}

template <bool A_MN, bool B_MN>
int dispatch_tc(cudaStream_t s, const CUtensorMap& tA, const CUtensorMap& tB, void* C, const float* bias,
                int ldc, int M, int N, int K, int epi, int splitk, int bn) {
#define GO(BN_, ST_)                                                                            \
    do {                                                                                        \
        if (epi == EPI_F32) return launch_tc<BN_, ST_, A_MN, B_MN, EPI_F32>(s, tA, tB, C, bias, ldc, M, N, K, splitk);   \
        if (epi == EPI_BF16) return launch_tc<BN_, ST_, A_MN, B_MN, EPI_BF16>(s, tA, tB, C, bias, ldc, M, N, K, splitk); \
        return launch_tc<BN_, ST_, A_MN, B_MN, EPI_ATOMIC>(s, tA, tB, C, bias, ldc, M, N, K, splitk);                    \
    } while (0)
    if (bn == 64) GO(64, 4);
    if (bn == 128) {
        if (epi == EPI_ATOMIC) return launch_tc<128, 6, A_MN, B_MN, EPI_ATOMIC>(s, tA, tB, C, bias, ldc, M, N, K, splitk);
        GO(128, 4);
    }
    GO(256, 3);
#undef GO
}

}  // namespace

// C[M,N] (+)= A * B^T with bf16 operands.
//   a_mn == 0: A stored row-major [M, K] (lda elements);  a_mn == 1: A stored [K, M]
//   b_mn == 0: B stored row-major [N, K] (ldb);           b_mn == 1: B stored [K, N]
//   epi: 0 fp32 store (+bias), 1 bf16 store, 2 fp32 atomic add (C pre-initialised; splitk >= 1)
MLB_API int mlb_gemm_bf16_tc(void* stream, const void* A, const void* B, void* C, const float* bias,
                             int M, int N, int K, int lda, int ldb, int ldc, int a_mn, int b_mn,
                             int epi, int splitk) {
    MLB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && splitk >= 1 && epi >= 0 && epi <= 2);
    MLB_REQUIRE(!(splitk > 1 && epi != EPI_ATOMIC));
    MLB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && mlb_aligned16(A) && mlb_aligned16(B) && mlb_aligned16(C));
    MLB_REQUIRE((epi == EPI_BF16 ? ldc % 8 == 0 : ldc % 4 == 0) && N % 8 == 0);
    if (!a_mn && !b_mn && epi == EPI_F32 && bias && splitk == 1 && mlb_aligned16(bias) &&
        tcp::gemm_persist_ok(M, N, K, ldc))           // minibatch-size head GEMM: persistent variant
        return tcp::launch_gemm_bias_persist(mlb_stream(stream), A, B, bias, static_cast<float*>(C), M, N, K, lda,
                                             ldb, ldc);
    int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    if (epi == EPI_ATOMIC && bn > 128) bn = 128;   // split-K: more output tiles, fewer K-splits
    CUtensorMap tA, tB;
    int rc;
    if (a_mn) rc = make_map(&tA, A, M, K, lda, 64, 64);          // [K, M]: inner = M
    else rc = make_map(&tA, A, K, M, lda, 64, 128);              // [M, K]: inner = K
    if (rc) return rc;
    if (b_mn) rc = make_map(&tB, B, N, K, ldb, 64, 64);          // [K, N]: inner = N
    else rc = make_map(&tB, B, K, N, ldb, 64, bn);               // [N, K]: inner = K
    if (rc) return rc;
    cudaStream_t s = mlb_stream(stream);
    if (!a_mn && !b_mn) return dispatch_tc<false, false>(s, tA, tB, C, bias, ldc, M, N, K, epi, splitk, bn);
    if (a_mn && b_mn) return dispatch_tc<true, true>(s, tA, tB, C, bias, ldc, M, N, K, epi, splitk, bn);
    if (a_mn) return dispatch_tc<true, false>(s, tA, tB, C, bias, ldc, M, N, K, epi, splitk, bn);
    return dispatch_tc<false, true>(s, tA, tB, C, bias, ldc, M, N, K, epi, splitk, bn);
}

// ------------------------------------------------------------------------------------------
// fp32 -> bf16 casts (activations entering the tensor-core path, bf16 weight copies)
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            const float4 v = *reinterpret_cast<const float4*>(src + i);
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
            uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(dst + i) = o;
        } else {
            for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
        }
    }
}
// dst[c, r] = bf16(src[r, c]) and (optionally) dst2[r, c] = bf16(src[r, c]); 32x32 smem tiles
__global__ void __launch_bounds__(256)
cast_transpose_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst_t,
                           __nv_bfloat16* __restrict__ dst, int rows, int cols, int ld_src,
                           int ld_t, int ld_d) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        float v = (r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f;
        tile[i][threadIdx.x] = v;
        if (dst && r < rows && c < cols) dst[(long long)r * ld_d + c] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) dst_t[(long long)c * ld_t + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
}
}  // namespace

MLB_API int mlb_cast_f32_bf16(void* stream, const float* src, void* dst, long long n) {
    MLB_REQUIRE(src && dst && n >= 0 && mlb_aligned16(src) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0);
    if (n == 0) return MLB_OK;
    long long g = (n / 4 + 255) / 256;
    if (g > MLB_NUM_SMS * 16) g = MLB_NUM_SMS * 16;
    if (g < 1) g = 1;
    cast_bf16_kernel<<<(unsigned)g, 256, 0, mlb_stream(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

// bf16 copies of one fp32 weight matrix W [rows, cols]: dst_t = W^T [cols, rows] (ld_t), and
// dst = W [rows, cols] (ld_d; may be NULL)
MLB_API int mlb_cast_weight_bf16(void* stream, const float* src, void* dst_t, void* dst, int rows,
                                 int cols, int ld_src, int ld_t, int ld_d) {
    MLB_REQUIRE(src && dst_t && rows > 0 && cols > 0);
    dim3 grid(mlb_cdiv(cols, 32), mlb_cdiv(rows, 32));
    cast_transpose_bf16_kernel<<<grid, dim3(32, 8), 0, mlb_stream(stream)>>>(
        src, reinterpret_cast<__nv_bfloat16*>(dst_t), reinterpret_cast<__nv_bfloat16*>(dst), rows, cols,
        ld_src, ld_t, ld_d);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
