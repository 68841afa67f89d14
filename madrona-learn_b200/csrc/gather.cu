// K5: minibatch gather straight from the rollout store (bit-exact indexing).
//
// The reference first relayouts the whole store [C, T', P, B, ...] -> [P, C*B, T', ...]
// (ml/rollouts.py:788-804), then per minibatch does take(axis 0) + swapaxes(0, 1)
// (ml/rollouts.py:319-329).  With trajectory j = c*B + b both steps collapse to
//     out[s, m, :] = store[j/B, s, j%B, :],  j = idx[m]      (P = 1)
// so the relayout copy never exists.  Rows are copied as 16-byte vectors when the row size
// and pointers allow it; a warp covers consecutive bytes of consecutive rows of one time
// step, so global stores are fully coalesced and loads are coalesced within each row.
#include "common.cuh"

namespace {

template <typename V>
__global__ void __launch_bounds__(256)
gather_kernel(const V* __restrict__ store, const int32_t* __restrict__ idx, V* __restrict__ out,
              int Tp, long long B, long long M, long long row_vecs) {
    const int s = blockIdx.y;
    const long long total = M * row_vecs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long m = e / row_vecs, k = e - m * row_vecs;
        const long long j = idx[m];
        const long long c = j / B, b = j - c * B;
        out[((long long)s * M + m) * row_vecs + k] =
            __ldg(store + (((c * Tp + s) * B) + b) * row_vecs + k);
    }
}

template <typename V>
int launch_gather(cudaStream_t st, const void* store, const int32_t* idx, void* out, int Tp,
                  long long B, long long M, long long row_vecs) {
    const long long total = M * row_vecs;
    long long gx = (total + 255) / 256;
    const long long cap = (long long)MLB_NUM_SMS * 8;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    gather_kernel<V><<<dim3((unsigned)gx, (unsigned)Tp), 256, 0, st>>>(
        reinterpret_cast<const V*>(store), idx, reinterpret_cast<V*>(out), Tp, B, M, row_vecs);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MLB_OK : (int)e;
}

int gather_dispatch(cudaStream_t st, const void* store, const int32_t* idx, void* out, int Tp,
                    long long B, long long M, long long row_bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(store) | reinterpret_cast<uintptr_t>(out);
    if (row_bytes % 16 == 0 && (a & 15) == 0)
        return launch_gather<uint4>(st, store, idx, out, Tp, B, M, row_bytes / 16);
    if (row_bytes % 8 == 0 && (a & 7) == 0)
        return launch_gather<uint2>(st, store, idx, out, Tp, B, M, row_bytes / 8);
    if (row_bytes % 4 == 0 && (a & 3) == 0)
        return launch_gather<uint32_t>(st, store, idx, out, Tp, B, M, row_bytes / 4);
    return launch_gather<uint8_t>(st, store, idx, out, Tp, B, M, row_bytes);
}

}  // namespace

MLB_API int mlb_mb_gather(void* stream, const void* store, const int32_t* idx, void* out, int C,
                          int Tp, long long B, long long M, long long row_bytes) {
    if (M == 0) return MLB_OK;
    MLB_REQUIRE(store && idx && out && C > 0 && Tp > 0 && B > 0 && M > 0 && row_bytes > 0);
    MLB_REQUIRE(Tp <= 65535);
    return gather_dispatch(mlb_stream(stream), store, idx, out, Tp, B, M, row_bytes);
}

MLB_API int mlb_mb_gather_rnn(void* stream, const void* store, const int32_t* idx, void* out,
                              int C, long long B, long long M, long long row_bytes) {
    if (M == 0) return MLB_OK;
    MLB_REQUIRE(store && idx && out && C > 0 && B > 0 && M > 0 && row_bytes > 0);
    // [C, B, row] is the Tp == 1 case of the step store: out[m] = store[j/B, j%B] = flat[j]
    return gather_dispatch(mlb_stream(stream), store, idx, out, 1, B, M, row_bytes);
}
