// Internal launchers of the persistent fused-layer kernels (mlp_tc_persist.cu), dispatched from the
// C-ABI entry points in mlp_tc_fused.cu.
#pragma once
#include <cuda_runtime.h>

namespace tcp {

int sm_count();
// MLB_TC_PERSIST=0 disables, =force uses the persistent kernels for every supported shape
bool persist_ok(int M, int K, int HN);
int launch_fwd_persist(cudaStream_t st, const void* X, const void* Wt, const float* scale, const float* bias,
                       void* Y, void* XH, float* rstd, int M, int K, int HN, int ldx, int ldw);
// C[M, N] fp32 = A[M, K] * Bt[N, K]^T + bias  (bias != NULL)
bool gemm_persist_ok(int M, int N, int K, int ldc);
int launch_gemm_bias_persist(cudaStream_t st, const void* A, const void* Bt, const float* bias, float* C,
                             int M, int N, int K, int lda, int ldb, int ldc);
int launch_dx_persist(cudaStream_t st, const void* DZ_in, const void* W, const float* scale, const float* bias,
                      const void* XH, const float* rstd, void* DZ_out, float* dscale, float* dbias, int M,
                      int K, int HN, int lda, int ldw);

// weight-streaming persistent kernels for HN = 512 (two 256-column accumulator halves per row tile)
bool stream_ok(int M, int K, int HN);
int launch_fwd_stream(cudaStream_t st, const void* X, const void* Wt, const float* scale, const float* bias,
                      void* Y, void* XH, float* rstd, int M, int K, int HN, int ldx, int ldw);
int launch_dx_stream(cudaStream_t st, const void* DZ_in, const void* W, const float* scale, const float* bias,
                     const void* XH, const float* rstd, void* DZ_out, float* dscale, float* dbias, int M,
                     int K, int HN, int lda, int ldw);

}  // namespace tcp
