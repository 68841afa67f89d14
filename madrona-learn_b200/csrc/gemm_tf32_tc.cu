// Tensor-core GEMM for compute_dtype=float32: tcgen05.mma.kind::tf32 on fp32 operands read straight from the
// fp32 activations / master weights (no operand copies), fp32 accumulation in TMEM.
//
// The reference's default compute dtype is float32 (ml/cfg.py:96) and XLA:GPU evaluates an f32 dot_general at
// its default precision on the tensor cores as TF32; this is that product for nn.Dense and its two autodiff
// transposes (ml/models.py:110-154, jax.value_and_grad at ml/ppo.py:276-281), with mlb_gemm_f32's interface:
//
//     C[M, N] (+)= op(A)[M, K] * op(B)[K, N] (+ bias[N])        op(X) = X or X^T, row-major storage
//
//   transA = 0   A stored [M, K]   K-major operand     transA = 1   A stored [K, M]   MN-major operand
//   transB = 1   B stored [N, K]   K-major operand     transB = 0   B stored [K, N]   MN-major operand
//
// so none of the three products of a Dense layer (Z = X W, dX = dZ W^T, dW = X^T dZ) transposes anything in
// memory.  A k-block is 32 fp32 = one 128-byte swizzle row (K-major: one 32(k) x 128(m) SWIZZLE_128B box;
// MN-major: 32(mn) x 32(k) boxes, 4096 B apart, in the 32-byte-atom 128 B swizzle that is the only MN-major
// layout tcgen05 accepts for 32-bit operands), UMMA K = 8 (32 bytes).
//
// Three kernels share that operand path (all mbarrier waits are bounded):
//   tf32_gemm_kernel            one output tile per CTA (anatomy of mlp_tc.cu: warp 0 TMA producer, warp 1 TMEM
//                               allocation + single-thread MMA issue, warps 2-5 epilogue): the split-K / accumulate
//                               reductions (red.global.add.v4) and problems with fewer tiles than SMs
//   tf32_gemm_persist_kernel    one CTA per SM walking its tiles with the accumulator double-buffered in TMEM and
//                               the output through swizzled smem panels + TMA store: Z = X W, dX = dZ W^T at
//                               minibatch size (mlb_gemm_tf32_tc picks it when tiles > SMs)
//   the same with LN = true     mlb_dense_ln_relu_fwd_tf32: LayerNorm + ReLU of the layer in the epilogue (eight
//                               epilogue warps, row statistics out of TMEM)
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int T32_THREADS = 192;
constexpr int BK32 = 32;               // 32 fp32 = 128 bytes
constexpr int UMMA_K32 = 8;            // 8 tf32 = 32 bytes per instruction

template <int BN, int STAGES>
struct Smem32 {
    static constexpr int A_BYTES = BM * BK32 * 4;          // 16 KB
    static constexpr int B_BYTES = BN * BK32 * 4;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + (3 * STAGES + 1) * 8 + 16 + 1024;
};

__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// shared-memory descriptor with layout type SWIZZLE_128B_BASE32B (= 1, bits [61,64)): 32-byte chunks swizzled
// inside a 128-byte row by (row & 3) -- what TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
}

// ATOMIC: fp32 vector reductions into a pre-initialised C (accumulate and / or split-K)
// ROUND:  the tensor core drops the low 13 mantissa bits of an fp32 operand (a biased conversion: every product
//         shrinks by ~1e-3).  With ROUND the four epilogue warps, idle during the main loop, convert a landed
//         stage in place with cvt.rna.tf32.f32 (round to nearest) before the MMA thread reads it: the operand
//         error becomes zero-mean and a 256-deep dot averages it out (rel-L2 2.9e-4 instead of 8.7e-4 against
//         the exact product, measured).  The pipeline is then TMA -> round -> MMA per stage; the conversion is
//         bound by shared-memory bandwidth (read + write of the stage), so ROUND = 1 converts only the
//         ACTIVATION operand of the store-epilogue products (A of Z = X W and dX = dZ W^T; both operands if A
//         is MN-major) and leaves the weight operand, two thirds of a forward stage, and the reduction
//         (dW-type) products to the hardware; ROUND = 2 converts both operands always.
template <int BN, int STAGES, bool A_MN, bool B_MN, bool ATOMIC, int ROUND>
__global__ void __launch_bounds__(T32_THREADS)
tf32_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, float* __restrict__ C, const float* __restrict__ bias, int ldc, int M, int N, int K,
                 int k_per_split) {
    using L = Smem32<BN, STAGES>;
    // ROUND = 1 leaves the reduction (accumulate / split-K: the dW-type products) to the hardware conversion:
    // there both operands are activations, the conversion pass doubles the shared-memory traffic of a kernel
    // that is bound by it (TMA write + convert read + convert write + MMA read per stage: 37 -> 43 us), and a
    // uniform -1e-3 scale on a gradient is invisible to Adam and to the global-norm clip
    constexpr bool DO_ROUND = ROUND == 2 || (ROUND == 1 && !ATOMIC);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem_1024(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* ready_bar = empty_bar + STAGES;              // DO_ROUND: stage converted, 128 arrivals
    uint64_t* acc_bar = ready_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);
    const int num_kb = (k_end - k_begin + BK32 - 1) / BK32;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 128);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && num_kb > 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait_spin(&empty_bar[s], ph ^ 1, 0x71);
                uint8_t* sa = smem + s * L::STAGE_BYTES;
                uint8_t* sb = sa + L::A_BYTES;
                mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
                const int k0 = k_begin + kb * BK32;
                if (A_MN) {         // A stored [K, M]: four 32(m) x 32(k) boxes
#pragma unroll
                    for (int j = 0; j < BM / 32; ++j) tma_load_2d(&tmA, &full_bar[s], sa + j * 4096, m0 + 32 * j, k0);
                } else {            // A stored [M, K]: one 32(k) x 128(m) box
                    tma_load_2d(&tmA, &full_bar[s], sa, k0, m0);
                }
                if (B_MN) {         // B stored [K, N]: BN/32 boxes of 32(n) x 32(k)
#pragma unroll
                    for (int j = 0; j < BN / 32; ++j) tma_load_2d(&tmB, &full_bar[s], sb + j * 4096, n0 + 32 * j, k0);
                } else {            // B stored [N, K]: one 32(k) x BN(n) box
                    tma_load_2d(&tmB, &full_bar[s], sb, k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && num_kb > 0) {
            // instruction descriptor: c = F32 [4,6) = 1, a = b = TF32 (format 2) [7,10) / [10,13), a_major bit 15,
            // b_major bit 16, N >> 3 [17,23), M >> 4 [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait_spin(DO_ROUND ? &ready_bar[s] : &full_bar[s], ph, 0x72);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
                const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK32 / UMMA_K32; ++k) {
                    // K-major (SWIZZLE_128B): 8-row groups 1024 B apart (SBO), K slice = +32 B inside the swizzle row
                    // MN-major (SWIZZLE_128B_BASE32B, the only MN-major layout of 32-bit operands): 32-wide MN
                    //           atoms 4096 B apart (LBO), 4-k-row groups 512 B apart (SBO), K slice = 8 k-rows
                    //           = +1024 B
                    const uint64_t ad = A_MN ? umma_desc_mn32(sa + k * 1024, 4096, 512) : umma_desc(sa + k * 32, 16, 1024);
                    const uint64_t bd = B_MN ? umma_desc_mn32(sb + k * 1024, 4096, 512) : umma_desc(sb + k * 32, 16, 1024);
                    tcgen05_mma_tf32(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
                }
                tcgen05_commit(&empty_bar[s]);
            }
            tcgen05_commit(acc_bar);
        }
    } else {
        const int quad = warp & 3;
        const int row = m0 + quad * 32 + lane;
        if (DO_ROUND) {
            // in-place round-to-nearest of every landed stage (element-wise: the swizzle does not matter);
            // the generic-proxy writes are fenced towards the async proxy the MMA reads through
            const int et = threadIdx.x - 64;                           // 0..127
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], (kb / STAGES) & 1, 0x74);
                const uint32_t base = smem_u32(smem + s * L::STAGE_BYTES) + et * 16;
#pragma unroll 8
                for (int v = 0; v < ((ROUND == 2 || A_MN) ? L::STAGE_BYTES : L::A_BYTES) / (128 * 16); ++v) {
                    uint32_t a, b, c, d;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + v * 2048));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(__uint_as_float(a)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(__uint_as_float(b)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(__uint_as_float(c)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(d)));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                 ::"r"(base + v * 2048), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
                }
                fence_async_smem();
                mbar_arrive(&ready_bar[s]);
            }
        }
        if (num_kb > 0) {
            mbar_wait(acc_bar, 0, 0x73);
            tcgen05_fence_after();
            const bool add_bias = bias != nullptr && blockIdx.z == 0;
            if (ATOMIC) {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int col0 = n0 + c * 32;
                    if (col0 >= N) break;                                   // warp-uniform
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
                    if (row < M) {
                        float* dst = C + (long long)row * ldc + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < N) {                             // N % 4 == 0
                                float4 v;
                                v.x = __uint_as_float(r[j]);     v.y = __uint_as_float(r[j + 1]);
                                v.z = __uint_as_float(r[j + 2]); v.w = __uint_as_float(r[j + 3]);
                                if (add_bias) {
                                    const float4 b = *reinterpret_cast<const float4*>(bias + col0 + j);
                                    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                                }
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                             ::"l"(dst + j), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                            }
                        }
                    }
                }
            } else {
                // Coalesced output through TMA: the accumulator is complete, so every operand stage is free; each
                // warp stages its 32 rows x 32 columns (thread = row) as a SWIZZLE_128B panel in the stage memory
                // (conflict-free 16-byte shared stores) and one lane hands the panel to cp.async.bulk.tensor --
                // full 128-byte rows instead of 32 scattered 16-byte pieces per store instruction; rows >= M and
                // columns >= N are clipped by the tensor map.  Two panels per warp: the store of chunk c drains
                // while chunk c + 1 is read out of TMEM.
                uint8_t* panels = smem + quad * 8192;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int col0 = n0 + c * 32;
                    if (col0 >= N) break;                                   // warp-uniform
                    uint8_t* panel = panels + (c & 1) * 4096;
                    if (c >= 2) {
                        if (lane == 0) tma_store_wait_read<1>();            // the store that last read this panel
                        __syncwarp();
                    }
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
                    const uint32_t pa = smem_u32(panel);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v;
                        v.x = __uint_as_float(r[4 * j]);     v.y = __uint_as_float(r[4 * j + 1]);
                        v.z = __uint_as_float(r[4 * j + 2]); v.w = __uint_as_float(r[4 * j + 3]);
                        if (add_bias && col0 + 4 * j < N) {
                            const float4 b = *reinterpret_cast<const float4*>(bias + col0 + 4 * j);
                            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                        }
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                     ::"r"(pa + sw128(lane, j)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, panel, col0, m0 + quad * 32);
                        tma_store_commit();
                    }
                }
                if (lane == 0) tma_store_wait<0>();
                __syncwarp();
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

// ------------------------------------------------------------------------------------------------------------
// Persistent variant of the store-epilogue kernel (no accumulate): one CTA per SM walks output tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ...  (tile t -> m-tile t / n_tiles, n-tile t % n_tiles: the n-tiles of one
// row block run back to back and share its A rows through L2).  The ncu capture of the one-tile-per-CTA kernel
// (profiles/r2_tf32_gemm_ncu_raw.csv) shows every unit below 25 % busy -- L2 19 %, DRAM 24 %, tensor pipe 22 %,
// 9 % of the warp slots -- i.e. the launch is a chain of per-CTA latencies (prologue -> operand stream -> MMA
// -> TMEM read-out -> store drain, 12 us per tile, nothing overlapped).  Here the accumulator is double-buffered
// in TMEM (2 x BN columns), so while the four epilogue warps read tile i out (tcgen05.ld -> swizzled smem
// panels -> TMA store) the producer / MMA warps already stream and multiply tile i + 1, and the prologue is paid
// once per SM instead of once per tile.  Operand rounding (ROUND) has its own four warps (the epilogue warps are
// busy with the previous tile).  Roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue,
// warps 6-9 rounding.
// ------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool LN = false>
struct Smem32P {
    static constexpr int A_BYTES = BM * BK32 * 4;
    static constexpr int B_BYTES = BN * BK32 * 4;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPW = LN ? 8 : 4;                          // epilogue warps
    static constexpr int PANEL_OFF = STAGES * STAGE_BYTES;          // per epilogue warp: 2 panels x 4 KB
    static constexpr int XCH_OFF = PANEL_OFF + EPW * 8192;          // LN: row-statistics exchange [2][2][128] float2
    static constexpr int BAR_OFF = XCH_OFF + (LN ? 4096 : 0);
    static constexpr int TOTAL = BAR_OFF + (3 * STAGES + 4) * 8 + 16 + 1024;
    static constexpr int THREADS_NOROUND = 64 + 32 * EPW;
};

// LN: the Dense -> LayerNorm -> ReLU layer of compute_dtype=float32 in one launch (ml/models.py:107-117): BN
// covers the whole feature row (N == BN), so the epilogue thread that owns an accumulator row computes the row's
// mean / rstd out of TMEM (pass 1), then re-reads it and emits z (the pre-activation the fp32 backward
// differentiates through; skipped when tmC is unused) and y = relu((z - mean) * rstd * scale + bias) as two
// panel streams (pass 2), plus stats[row] = {mean, rstd}.  Same arithmetic as mlb_ln_relu_fwd_f32 (eps 1e-6,
// fast variance); saves that kernel's 8 * rows * H bytes and its launch.  The LayerNorm epilogue is ~10
// instructions per element, so it runs on EIGHT warps (two per TMEM lane quadrant, each half of the columns;
// the two partial row sums meet through shared memory) -- with four the epilogue, not the operand stream,
// set the tile period (13 us instead of 7.5).
template <int BN, int STAGES, bool A_MN, bool B_MN, int ROUND, bool LN>
__global__ void __launch_bounds__((LN ? 320 : 192) + (ROUND ? 128 : 0))
tf32_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmY,
                         const float* __restrict__ bias, const float* __restrict__ ln_scale,
                         const float* __restrict__ ln_bias, float* __restrict__ stats, int store_z, int M, int N,
                         int K, int n_tiles, int num_tiles) {
    using L = Smem32P<BN, STAGES, LN>;
    constexpr uint32_t TCOLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem_1024(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* ready_bar = empty_bar + STAGES;
    uint64_t* tfull_bar = ready_bar + STAGES;              // [2] accumulator complete (MMA -> epilogue)
    uint64_t* tempty_bar = tfull_bar + 2;                  // [2] accumulator drained (4 epilogue warps -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + BK32 - 1) / BK32;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
        if (LN) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 128);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], L::EPW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: one continuous stream of k-blocks over all of this CTA's tiles
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_spin(&empty_bar[s], ((it / STAGES) & 1) ^ 1, 0x75);
                    uint8_t* sa = smem + s * L::STAGE_BYTES;
                    uint8_t* sb = sa + L::A_BYTES;
                    mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
                    const int k0 = kb * BK32;
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < BM / 32; ++j) tma_load_2d(&tmA, &full_bar[s], sa + j * 4096, m0 + 32 * j, k0);
                    } else {
                        tma_load_2d(&tmA, &full_bar[s], sa, k0, m0);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j) tma_load_2d(&tmB, &full_bar[s], sb + j * 4096, n0 + 32 * j, k0);
                    } else {
                        tma_load_2d(&tmB, &full_bar[s], sb, k0, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            uint32_t it = 0, li = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++li) {
                const uint32_t acc = li & 1;
                mbar_wait_spin(&tempty_bar[acc], ((li >> 1) & 1) ^ 1, 0x76);       // epilogue drained this buffer
                tcgen05_fence_after();
                const uint32_t tacc = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_spin(ROUND ? &ready_bar[s] : &full_bar[s], (it / STAGES) & 1, 0x77);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
                    const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK32 / UMMA_K32; ++k) {
                        const uint64_t ad = A_MN ? umma_desc_mn32(sa + k * 1024, 4096, 512) : umma_desc(sa + k * 32, 16, 1024);
                        const uint64_t bd = B_MN ? umma_desc_mn32(sb + k * 1024, 4096, 512) : umma_desc(sb + k * 32, 16, 1024);
                        tcgen05_mma_tf32(tacc, ad, bd, idesc, (kb | k) ? 1u : 0u);
                    }
                    tcgen05_commit(&empty_bar[s]);
                }
                tcgen05_commit(&tfull_bar[acc]);
            }
        }
    } else if (warp < 2 + L::EPW) {
        // ================= epilogue (warps 2..5 [6..9] -> TMEM lane quadrants 2, 3, 0, 1 [again])
        const int quad = warp & 3;
        const int ew = warp - 2;
        uint8_t* panels = smem + L::PANEL_OFF + ew * 8192;
        uint32_t li = 0, cnt = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++li) {
            const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
            const uint32_t acc = li & 1;
            mbar_wait(&tfull_bar[acc], (li >> 1) & 1, 0x78);
            tcgen05_fence_after();
            if (LN) {
                constexpr int CH = BN / 64;                                 // 32-column chunks per warp (half a row)
                const int half = ew >> 2;
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + (uint32_t)(half * CH * 32);
                float sum = 0.f, sq = 0.f;
#pragma unroll 1
                for (int c = 0; c < CH; ++c) {                              // pass 1: row statistics of this half
                    uint32_t r[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float v = __uint_as_float(r[j]); sum += v; sq = fmaf(v, v, sq); }
                }
                // the partner warp of this quadrant holds the other half: exchange through shared memory
                // (double-buffered by tile parity; the pair meets at a 64-thread named barrier)
                float2* xch = reinterpret_cast<float2*>(smem + L::XCH_OFF) + (li & 1) * 256;
                xch[half * 128 + quad * 32 + lane] = make_float2(sum, sq);
                named_bar_sync(1 + quad, 64);
                const float2 o = xch[(half ^ 1) * 128 + quad * 32 + lane];
                sum += o.x; sq += o.y;
                const float invH = 1.f / (float)BN;
                const float mean = sum * invH;
                const float rstd = rsqrtf(fmaxf(0.f, sq * invH - mean * mean) + 1e-6f);
                const int row = m0 + quad * 32 + lane;
                if (half == 0 && stats != nullptr && row < M)
                    *reinterpret_cast<float2*>(stats + 2 * (long long)row) = make_float2(mean, rstd);
                uint8_t* pz = panels;
                uint8_t* py = panels + 4096;
                const uint32_t paz = smem_u32(pz), pay = smem_u32(py);
#pragma unroll 1
                for (int c = 0; c < CH; ++c) {                              // pass 2: z and y panels
                    const int col0 = (half * CH + c) * 32;
                    if (cnt >= 1) {
                        if (lane == 0) tma_store_wait_read<0>();            // the panels' previous stores were read
                        __syncwarp();
                    }
                    ++cnt;
                    uint32_t r[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 sc = __ldg(reinterpret_cast<const float4*>(ln_scale + col0) + j);
                        const float4 bi = __ldg(reinterpret_cast<const float4*>(ln_bias + col0) + j);
                        float4 v, y;
                        v.x = __uint_as_float(r[4 * j]);     v.y = __uint_as_float(r[4 * j + 1]);
                        v.z = __uint_as_float(r[4 * j + 2]); v.w = __uint_as_float(r[4 * j + 3]);
                        y.x = fmaxf(0.f, (v.x - mean) * rstd * sc.x + bi.x);
                        y.y = fmaxf(0.f, (v.y - mean) * rstd * sc.y + bi.y);
                        y.z = fmaxf(0.f, (v.z - mean) * rstd * sc.z + bi.z);
                        y.w = fmaxf(0.f, (v.w - mean) * rstd * sc.w + bi.w);
                        if (store_z)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(paz + sw128(lane, j)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                     ::"r"(pay + sw128(lane, j)), "f"(y.x), "f"(y.y), "f"(y.z), "f"(y.w) : "memory");
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {                                        // one bulk group per chunk (z + y)
                        if (store_z) tma_store_2d(&tmC, pz, col0, m0 + quad * 32);
                        tma_store_2d(&tmY, py, col0, m0 + quad * 32);
                        tma_store_commit();
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n0 + c * 32;
                if (col0 >= N) break;                                       // warp-uniform
                uint8_t* panel = panels + (cnt & 1) * 4096;
                if (cnt >= 2) {
                    if (lane == 0) tma_store_wait_read<1>();                // the store that last read this panel
                    __syncwarp();
                }
                ++cnt;
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
                const uint32_t pa = smem_u32(panel);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v;
                    v.x = __uint_as_float(r[4 * j]);     v.y = __uint_as_float(r[4 * j + 1]);
                    v.z = __uint_as_float(r[4 * j + 2]); v.w = __uint_as_float(r[4 * j + 3]);
                    if (bias != nullptr && col0 + 4 * j < N) {
                        const float4 b = *reinterpret_cast<const float4*>(bias + col0 + 4 * j);
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                 ::"r"(pa + sw128(lane, j)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, panel, col0, m0 + quad * 32);
                    tma_store_commit();
                }
            }
            }
            // every lane's tcgen05.ld of this accumulator has completed (tmem_ld32 waits): hand it back
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        if (lane == 0) tma_store_wait<0>();
        __syncwarp();
    } else if (ROUND) {
        // ================= operand rounding (warps 6..9): cvt.rna.tf32.f32 in place on every landed stage
        const int et = threadIdx.x - L::THREADS_NOROUND;                    // 0..127
        uint32_t it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full_bar[s], (it / STAGES) & 1, 0x79);
                const uint32_t base = smem_u32(smem + s * L::STAGE_BYTES) + et * 16;
#pragma unroll 8
                for (int v = 0; v < ((ROUND == 2 || A_MN) ? L::STAGE_BYTES : L::A_BYTES) / (128 * 16); ++v) {
                    uint32_t a, b, c, d;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + v * 2048));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(__uint_as_float(a)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(__uint_as_float(b)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(__uint_as_float(c)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(d)));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                 ::"r"(base + v * 2048), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
                }
                fence_async_smem();
                mbar_arrive(&ready_bar[s]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

// 2-D fp32 row-major tensor [outer, inner] with row stride ld (elements); box = [box_outer, 32 inner]
int make_map_f32(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_outer,
                 bool mn_major) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return MLB_EINVAL;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MLB_OK : MLB_EINVAL;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, bool ATOMIC, int ROUND>
int launch32(cudaStream_t s, const CUtensorMap& tA, const CUtensorMap& tB, float* C, const float* bias, int ldc,
             int M, int N, int K, int splitk) {
    CUtensorMap tC{};
    if (!ATOMIC) {                 // output panels: 32 columns x 32 rows, SWIZZLE_128B
        const int rc = make_map_f32(&tC, C, N, M, ldc, 32, false);
        if (rc) return rc;
    }
    using L = Smem32<BN, STAGES>;
    auto kern = tf32_gemm_kernel<BN, STAGES, A_MN, B_MN, ATOMIC, ROUND>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return (int)e;
    int kps = (K + splitk - 1) / splitk;
    kps = (kps + BK32 - 1) / BK32 * BK32;
    const int zs = (K + kps - 1) / kps;
    dim3 grid(mlb_cdiv(M, BM), mlb_cdiv(N, BN), zs);
    e = launch_pdl(kern, grid, dim3(T32_THREADS), L::TOTAL, s, tA, tB, tC, C, bias, ldc, M, N, K, kps);
    return e == cudaSuccess ? MLB_OK : (int)e;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int ROUND>
int launch32p(cudaStream_t s, const CUtensorMap& tA, const CUtensorMap& tB, float* C, const float* bias, int ldc,
              int M, int N, int K) {
    using L = Smem32P<BN, STAGES, false>;
    CUtensorMap tC{};
    const int rc = make_map_f32(&tC, C, N, M, ldc, 32, false);
    if (rc) return rc;
    auto kern = tf32_gemm_persist_kernel<BN, STAGES, A_MN, B_MN, ROUND, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return (int)e;
    const int n_tiles = (int)mlb_cdiv(N, BN);
    const long long tiles = (long long)mlb_cdiv(M, BM) * n_tiles;
    const unsigned grid = (unsigned)(tiles < MLB_NUM_SMS ? tiles : MLB_NUM_SMS);
    e = launch_pdl(kern, dim3(grid), dim3(L::THREADS_NOROUND + (ROUND ? 128 : 0)), L::TOTAL, s, tA, tB, tC, tC, bias,
                   (const float*)nullptr, (const float*)nullptr, (float*)nullptr, 0, M, N, K, n_tiles, (int)tiles);
    return e == cudaSuccess ? MLB_OK : (int)e;
}

// Dense (B = W stored [K, H], MN-major) + LayerNorm + ReLU, H == BN
template <int BN, int STAGES, int ROUND>
int launch32ln(cudaStream_t s, const float* x, const float* w, const float* scale, const float* bias, float* z,
               float* y, float* stats, long long rows, int K, int ldx) {
    using L = Smem32P<BN, STAGES, true>;
    CUtensorMap tA, tB, tZ{}, tY;
    int rc = make_map_f32(&tA, x, K, rows, ldx, BM, false);
    if (!rc) rc = make_map_f32(&tB, w, BN, K, BN, 32, true);
    if (!rc) rc = make_map_f32(&tY, y, BN, rows, BN, 32, false);
    if (!rc && z) rc = make_map_f32(&tZ, z, BN, rows, BN, 32, false);
    if (rc) return rc;
    auto kern = tf32_gemm_persist_kernel<BN, STAGES, false, true, ROUND, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return (int)e;
    const long long tiles = mlb_cdiv(rows, BM);
    const unsigned grid = (unsigned)(tiles < MLB_NUM_SMS ? tiles : MLB_NUM_SMS);
    e = launch_pdl(kern, dim3(grid), dim3(L::THREADS_NOROUND + (ROUND ? 128 : 0)), L::TOTAL, s, tA, tB, z ? tZ : tY, tY,
                   (const float*)nullptr, scale, bias, stats, z ? 1 : 0, (int)rows, BN, K, 1, (int)tiles);
    return e == cudaSuccess ? MLB_OK : (int)e;
}

template <bool A_MN, bool B_MN, int ROUND>
int dispatch32p(cudaStream_t s, const CUtensorMap& tA, const CUtensorMap& tB, float* C, const float* bias, int ldc,
                int M, int N, int K, int bn) {
    if (bn == 32) return launch32p<32, 6, A_MN, B_MN, ROUND>(s, tA, tB, C, bias, ldc, M, N, K);
    if (bn == 64) return launch32p<64, 6, A_MN, B_MN, ROUND>(s, tA, tB, C, bias, ldc, M, N, K);
    if (bn == 128) return launch32p<128, 5, A_MN, B_MN, ROUND>(s, tA, tB, C, bias, ldc, M, N, K);
    return launch32p<256, 4, A_MN, B_MN, ROUND>(s, tA, tB, C, bias, ldc, M, N, K);
}

template <bool A_MN, bool B_MN, int ROUND>
int dispatch32(cudaStream_t s, const CUtensorMap& tA, const CUtensorMap& tB, float* C, const float* bias, int ldc,
               int M, int N, int K, bool atomic, int splitk, int bn, int stages256) {
#define GO(BN_, ST_)                                                                                              \
    do {                                                                                                          \
        if (atomic) return launch32<BN_, ST_, A_MN, B_MN, true, ROUND>(s, tA, tB, C, bias, ldc, M, N, K, splitk); \
        return launch32<BN_, ST_, A_MN, B_MN, false, ROUND>(s, tA, tB, C, bias, ldc, M, N, K, splitk);            \
    } while (0)
    if (bn == 32) GO(32, 6);
    if (bn == 64) GO(64, 6);
    if (bn == 128) GO(128, 6);
    // 256-wide tiles (the Dense forward / dX): 2 stages = 96 KB, two CTAs per SM (256 TMEM columns each), so one
    // CTA's epilogue runs under the other's main loop; 4 stages = 192 KB, one CTA per SM
    if (stages256 == 2) GO(256, 2);
    GO(256, 4);
#undef GO
}

// immutable after first use (read once): MLB_TF32_ROUND = 0 hardware conversion only, 1 (default) activations
// rounded to nearest, 2 both operands; MLB_TF32_STAGES = 2 selects the two-CTAs-per-SM pipeline of the 256-wide
// tiles (measured slower than 4 stages / one CTA: 55 vs 51 us on 65536 x 256 x 256)
int tf32_knob(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

MLB_API int mlb_gemm_tf32_ok(int M, int N, int K, int lda, int ldb, int ldc, const void* A, const void* B,
                             const void* C) {
    return M > 0 && N > 0 && K > 0 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 &&
           mlb_aligned16(A) && mlb_aligned16(B) && mlb_aligned16(C);
}

// Same contract as mlb_gemm_f32 (split-K needs accumulate != 0; with accumulate the launch picks its own
// split so that the output tiles x K-slices fill the SMs -- `splitk` is a lower bound).
MLB_API int mlb_gemm_tf32_tc(void* stream, const float* A, const float* B, float* C, const float* bias, int M,
                             int N, int K, int lda, int ldb, int ldc, int transA, int transB, int accumulate,
                             int splitk) {
    MLB_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0 && splitk >= 1);
    if (M == 0 || N == 0) return MLB_OK;
    MLB_REQUIRE(!(splitk > 1 && !accumulate));
    MLB_REQUIRE(mlb_gemm_tf32_ok(M, N, K, lda, ldb, ldc, A, B, C));
    MLB_REQUIRE(!bias || mlb_aligned16(bias));
    const bool a_mn = transA != 0, b_mn = transB == 0;
    int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
    static const int dw_bn = tf32_knob("MLB_TF32_DW_BN", 128);
    if (accumulate && bn > dw_bn) bn = dw_bn;      // reductions: more output tiles, fewer K-slices
    if (accumulate) {
        const long long tiles = (long long)mlb_cdiv(M, BM) * mlb_cdiv(N, bn);
        long long want = MLB_NUM_SMS / (tiles > 0 ? tiles : 1);
        const long long cap = K / 256;             // >= 8 k-blocks per slice
        if (want > cap) want = cap;
        if (want > splitk) splitk = (int)want;
    }
    CUtensorMap tA, tB;
    int rc;
    if (a_mn) rc = make_map_f32(&tA, A, M, K, lda, 32, true);           // [K, M]: inner = M, 32(m) x 32(k) boxes
    else rc = make_map_f32(&tA, A, K, M, lda, BM, false);                // [M, K]: inner = K, 32(k) x 128(m) box
    if (rc) return rc;
    if (b_mn) rc = make_map_f32(&tB, B, N, K, ldb, 32, true);           // [K, N]: inner = N
    else rc = make_map_f32(&tB, B, K, N, ldb, bn, false);                // [N, K]: inner = K
    if (rc) return rc;
    cudaStream_t s = mlb_stream(stream);
    const bool atomic = accumulate != 0;
    static const int round = tf32_knob("MLB_TF32_ROUND", 1), st256 = tf32_knob("MLB_TF32_STAGES", 4);
    static const int persist = tf32_knob("MLB_TF32_PERSIST", 1);
    // more output tiles than SMs and a plain store: the persistent kernel (accumulator double-buffered in TMEM)
    if (persist && !atomic && (long long)mlb_cdiv(M, BM) * mlb_cdiv(N, bn) > MLB_NUM_SMS) {
#define DISPATCHP(R_)                                                                                          \
    do {                                                                                                       \
        if (!a_mn && !b_mn) return dispatch32p<false, false, R_>(s, tA, tB, C, bias, ldc, M, N, K, bn);        \
        if (a_mn && b_mn) return dispatch32p<true, true, R_>(s, tA, tB, C, bias, ldc, M, N, K, bn);            \
        if (a_mn) return dispatch32p<true, false, R_>(s, tA, tB, C, bias, ldc, M, N, K, bn);                   \
        return dispatch32p<false, true, R_>(s, tA, tB, C, bias, ldc, M, N, K, bn);                             \
    } while (0)
        if (round == 1) DISPATCHP(1);
        if (round >= 2) DISPATCHP(2);
        DISPATCHP(0);
#undef DISPATCHP
    }
#define DISPATCH(R_)                                                                                                    \
    do {                                                                                                                \
        if (!a_mn && !b_mn) return dispatch32<false, false, R_>(s, tA, tB, C, bias, ldc, M, N, K, atomic, splitk, bn, st256); \
        if (a_mn && b_mn) return dispatch32<true, true, R_>(s, tA, tB, C, bias, ldc, M, N, K, atomic, splitk, bn, st256);     \
        if (a_mn) return dispatch32<true, false, R_>(s, tA, tB, C, bias, ldc, M, N, K, atomic, splitk, bn, st256);            \
        return dispatch32<false, true, R_>(s, tA, tB, C, bias, ldc, M, N, K, atomic, splitk, bn, st256);                      \
    } while (0)
    if (round == 1) DISPATCH(1);
    if (round >= 2) DISPATCH(2);
    DISPATCH(0);
#undef DISPATCH
}

// One MLP layer of compute_dtype=float32 in one launch: z = x W (tcgen05 kind::tf32), y = relu(LayerNorm(z)).
//   x [rows, K] (row stride ldx), w [K, H] row-major (the flax Dense kernel), scale / bias [H],
//   z [rows, H] or NULL (rollout inference does not need the pre-activation), y [rows, H],
//   stats [rows][2] = {mean, rstd} or NULL.  H in {64, 128, 256} (one accumulator tile = the whole feature row).
MLB_API int mlb_dense_ln_relu_fwd_tf32(void* stream, const float* x, const float* w, const float* scale,
                                       const float* bias, float* z, float* y, float* stats, long long rows, int K,
                                       int H, int ldx) {
    MLB_REQUIRE(x && w && scale && bias && y && rows >= 0 && K > 0 && ldx >= K && ldx % 4 == 0);
    MLB_REQUIRE(H == 64 || H == 128 || H == 256);
    MLB_REQUIRE(rows <= 0x7FFFFFFF);
    if (rows == 0) return MLB_OK;
    MLB_REQUIRE(mlb_aligned16(x) && mlb_aligned16(w) && mlb_aligned16(y) && mlb_aligned16(scale) &&
                mlb_aligned16(bias) && (!z || mlb_aligned16(z)) && (!stats || (reinterpret_cast<uintptr_t>(stats) & 7) == 0));
    static const int round = tf32_knob("MLB_TF32_ROUND", 1);
    cudaStream_t s = mlb_stream(stream);
#define LNGO(R_)                                                                                          \
    do {                                                                                                  \
        if (H == 64) return launch32ln<64, 6, R_>(s, x, w, scale, bias, z, y, stats, rows, K, ldx);        \
        if (H == 128) return launch32ln<128, 4, R_>(s, x, w, scale, bias, z, y, stats, rows, K, ldx);      \
        return launch32ln<256, 3, R_>(s, x, w, scale, bias, z, y, stats, rows, K, ldx);                    \
    } while (0)
    if (round == 1) LNGO(1);
    if (round >= 2) LNGO(2);
    LNGO(0);
#undef LNGO
}
