// fp32 SIMT GEMM for compute_dtype=float32 (the reference's default, ml/cfg.py:96).
//
// Replaces the XLA dot_general of nn.Dense (ml/models.py:110-115,129-135,148-154) and its
// autodiff transposes (jax.value_and_grad, ml/ppo.py:276-281) with true-fp32 FFMA math so the
// fp32 path agrees with the reference's CPU (non-TF32) numerics to ~1e-6.  The tensor-core
// (tcgen05, bf16) path lives in mlp_tc.cu.
//
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N])
//   op(A) = A (row-major, lda) or A^T (A stored [K,M]);  op(B) likewise.
//   splitk > 1: K is cut in `splitk` slices (grid.z), partial products are reduced with
//   fp32 atomics into C, which must be pre-initialised (used for dW = X^T dZ, K = rows).
//
// Tiling: BM x BN block tile, BK = 16, 256 threads, TM x TN register tile, shared tiles stored
// k-major, next tile prefetched into registers while the current one is consumed.
#include "common.cuh"

namespace {

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             const float* __restrict__ bias, int M, int N, int K, int lda, int ldb, int ldc,
             int accumulate, int k_per_split) {
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
    constexpr int PA = 4, PB = 4;
    __shared__ __align__(16) float As[2][BK][BM + PA];
    __shared__ __align__(16) float Bs[2][BK][BN + PB];
    constexpr int A_PER = BM * BK / 256;
    constexpr int B_PER = BN * BK / 256;

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    // 8-wide register tiles are split in two float4 halves BM/2 (BN/2) apart so that the 16
    // lanes of a half-warp read one contiguous 256 B run of the shared tile (no bank conflict)
    auto row_of = [&](int i) { return TM == 8 ? (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * TM + i; };
    auto col_of = [&](int j) { return TN == 8 ? (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4)) : tx * TN + j; };

    float ra[A_PER], rb[B_PER];
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int e = tid + i * 256;
            int m, k;
            if (TA) { m = e % BM; k = e / BM; }       // A stored [K, M]: contiguous along m
            else    { k = e % BK; m = e / BK; }       // A stored [M, K]: contiguous along k
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < kend) v = TA ? __ldg(A + (long long)gk * lda + gm)
                                            : __ldg(A + (long long)gm * lda + gk);
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int e = tid + i * 256;
            int n, k;
            if (TB) { k = e % BK; n = e / BK; }       // B stored [N, K]: contiguous along k
            else    { n = e % BN; k = e / BN; }       // B stored [K, N]: contiguous along n
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < kend) v = TB ? __ldg(B + (long long)gn * ldb + gk)
                                            : __ldg(B + (long long)gk * ldb + gn);
            rb[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int e = tid + i * 256;
            int m, k;
            if (TA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
            As[buf][k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int e = tid + i * 256;
            int n, k;
            if (TB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
            Bs[buf][k][n] = rb[i];
        }
    };

    if (kbeg < kend) {
        load_tiles(kbeg);
        store_tiles(0);
        __syncthreads();
        int buf = 0;
        for (int k0 = kbeg; k0 < kend; k0 += BK) {
            const bool more = k0 + BK < kend;
            if (more) load_tiles(k0 + BK);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][row_of(i)]);
                    a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][col_of(j)]);
                    b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (more) {
                store_tiles(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }

    const bool atomic = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + row_of(i);
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + col_of(j);
            if (gn >= N) continue;
            float v = acc[i][j];
            float* c = C + (long long)gm * ldc + gn;
            if (atomic) {
                if (bias && blockIdx.z == 0) v += bias[gn];
                atomicAdd(c, v);
            } else {
                if (bias) v += bias[gn];
                if (accumulate) v += *c;
                *c = v;
            }
        }
    }
}

template <int BM, int BN, int TM, int TN>
int launch(cudaStream_t s, const float* A, const float* B, float* C, const float* bias, int M,
           int N, int K, int lda, int ldb, int ldc, int ta, int tb, int accumulate, int splitk) {
    const dim3 grid(mlb_cdiv(M, BM), mlb_cdiv(N, BN), splitk);
    int kps = (K + splitk - 1) / splitk;
    kps = (kps + BK - 1) / BK * BK;
#define GO(TA_, TB_) sgemm_kernel<BM, BN, TM, TN, TA_, TB_><<<grid, 256, 0, s>>>( \
        A, B, C, bias, M, N, K, lda, ldb, ldc, accumulate, kps)
    if (!ta && !tb) GO(false, false);
    else if (!ta && tb) GO(false, true);
    else if (ta && !tb) GO(true, false);
    else GO(true, true);
#undef GO
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MLB_OK : (int)e;
}

}  // namespace

MLB_API int mlb_gemm_f32(void* stream, const float* A, const float* B, float* C,
                         const float* bias, int M, int N, int K, int lda, int ldb, int ldc,
                         int transA, int transB, int accumulate, int splitk) {
    MLB_REQUIRE(A && B && C && M >= 0 && N >= 0 && K >= 0 && splitk >= 1);
    if (M == 0 || N == 0) return MLB_OK;
    MLB_REQUIRE(!(splitk > 1 && !accumulate));   // split-K reduces into a pre-initialised C
    cudaStream_t s = mlb_stream(stream);
    if (N <= 32) return launch<128, 32, 4, 4>(s, A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, accumulate, splitk);
    if (N <= 64 || M <= 64) return launch<64, 64, 4, 4>(s, A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, accumulate, splitk);
    return launch<128, 128, 8, 8>(s, A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, accumulate, splitk);
}
