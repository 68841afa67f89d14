// PBT policy-batch reorder: _compute_reorder_chunks (ml/rollouts.py:1107-1190).
//
// Agents (assignments int32 [S], values in [0, P)) are stably sorted by policy; every policy's run
// is cut into full chunks of C (packed first, in policy order) and one partial chunk placed at
// partial_base + p*C.  Outputs: to_policy_idxs [B, C] (padding slots repeat the chunk's first
// entry; an entirely empty chunk keeps the out-of-range sentinel S, which the reference's gather
// clips) and to_sim_idxs [S] (the inverse map).  The reference does this with argsort + nonzero +
// scatters; here it is a stable counting sort: per-block histograms, one scan block, one scatter
// pass with warp match-any ranks -- integer-only, bit-exact (golden: the reference's own KATs).
#include "common.cuh"

namespace {

constexpr int RO_BLOCK = 256;

__global__ void __launch_bounds__(RO_BLOCK)
reorder_hist_kernel(const int32_t* __restrict__ a, long long S, int P, int32_t* __restrict__ block_hist) {
    extern __shared__ int32_t h[];
    for (int p = threadIdx.x; p < P; p += blockDim.x) h[p] = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * RO_BLOCK + threadIdx.x;
    if (i < S) atomicAdd(&h[min(max(a[i], 0), P - 1)], 1);
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) block_hist[(long long)blockIdx.x * P + p] = h[p];
}

// tab[p] = {full_count, full_start, partial_start, count}; block_hist -> exclusive prefix over blocks
__global__ void __launch_bounds__(1024)
reorder_scan_kernel(int32_t* __restrict__ block_hist, int nblocks, int P, int C, int32_t* __restrict__ tab) {
    extern __shared__ int32_t sm[];              // counts[P] | full_counts[P]
    int32_t* counts = sm;
    int32_t* fullc = sm + P;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        int32_t run = 0;
        for (int b = 0; b < nblocks; ++b) {
            const int32_t v = block_hist[(long long)b * P + p];
            block_hist[(long long)b * P + p] = run;
            run += v;
        }
        counts[p] = run;
        fullc[p] = (run / C) * C;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t partial_base = 0;
        for (int p = 0; p < P; ++p) partial_base += fullc[p];
        int32_t full_start = 0;
        for (int p = 0; p < P; ++p) {
            tab[4 * p + 0] = fullc[p];
            tab[4 * p + 1] = full_start;
            tab[4 * p + 2] = partial_base + p * C - fullc[p];
            tab[4 * p + 3] = counts[p];
            full_start += fullc[p];
        }
    }
}

__global__ void __launch_bounds__(RO_BLOCK)
reorder_scatter_kernel(const int32_t* __restrict__ a, long long S, int P, const int32_t* __restrict__ block_off,
                       const int32_t* __restrict__ tab, int32_t* __restrict__ to_policy,
                       int32_t* __restrict__ to_sim, long long slots) {
    extern __shared__ int32_t wc[];               // [warps][P] per-warp counts
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < nw * P; k += blockDim.x) wc[k] = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * RO_BLOCK + threadIdx.x;
    const bool on = i < S;
    const int p = on ? min(max(a[i], 0), P - 1) : -1 - lane;       // inactive lanes match nobody
    const unsigned m = __match_any_sync(0xffffffffu, p);
    const int rank_w = __popc(m & ((1u << lane) - 1u));
    if (on && rank_w == 0) wc[warp * P + p] = __popc(m);
    __syncthreads();
    if (!on) return;
    int offs = block_off[(long long)blockIdx.x * P + p] + rank_w;
    for (int w = 0; w < warp; ++w) offs += wc[w * P + p];
    const int fullc = tab[4 * p], pos = offs < fullc ? tab[4 * p + 1] + offs : tab[4 * p + 2] + offs;
    to_sim[i] = pos;
    if (pos < slots) to_policy[pos] = (int32_t)i;       // B too small for the layout: never write outside
}

__global__ void __launch_bounds__(256)
reorder_fill_kernel(int32_t* __restrict__ x, long long n, int32_t v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

// padding slots of a chunk repeat its first entry (ml/rollouts.py:1184-1186)
__global__ void __launch_bounds__(256)
reorder_pad_kernel(int32_t* __restrict__ to_policy, long long B, int C, int32_t S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C || i % C == 0) return;
    if (to_policy[i] == S) to_policy[i] = to_policy[(i / C) * C];
}

// out[k, :] = src[clip(idx[k], 0, n_src - 1), :]   (x.at[idx].get(mode='clip'), ml/rollouts.py:143-152)
template <typename V>
__global__ void __launch_bounds__(256)
gather_rows_clip_kernel(const V* __restrict__ src, const int32_t* __restrict__ idx, V* __restrict__ out,
                        long long n_idx, long long n_src, int row_vecs) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_idx * row_vecs) return;
    const long long k = e / row_vecs, c = e - k * row_vecs;
    long long j = idx[k];
    j = j < 0 ? 0 : (j >= n_src ? n_src - 1 : j);
    out[e] = __ldg(src + j * row_vecs + c);
}

}  // namespace

MLB_API size_t mlb_reorder_chunks_workspace(long long S, int P) {
    const long long nb = (S + RO_BLOCK - 1) / RO_BLOCK;
    return (size_t)(nb * P + 4 * (long long)P) * sizeof(int32_t);
}

MLB_API int mlb_reorder_chunks(void* stream, const int32_t* assignments, long long S, int P, int C,
                               long long B, int32_t* to_policy, int32_t* to_sim, void* ws,
                               size_t ws_bytes) {
    MLB_REQUIRE(assignments && to_policy && to_sim && S > 0 && P > 0 && P <= 1024 && C > 0 && B > 0);
    MLB_REQUIRE(S < (1ll << 31) && B * C < (1ll << 31));
    if (!ws || ws_bytes < mlb_reorder_chunks_workspace(S, P)) return MLB_EWS;
    const long long nb = (S + RO_BLOCK - 1) / RO_BLOCK;
    // every agent must have a slot: sum of full chunks + one partial chunk per policy
    MLB_REQUIRE(B * C >= S);
    cudaStream_t st = mlb_stream(stream);
    int32_t* block_hist = static_cast<int32_t*>(ws);
    int32_t* tab = block_hist + nb * P;
    reorder_fill_kernel<<<mlb_cdiv(B * C, 256), 256, 0, st>>>(to_policy, B * C, (int32_t)S);
    MLB_CHECK_LAUNCH();
    reorder_hist_kernel<<<(unsigned)nb, RO_BLOCK, P * sizeof(int32_t), st>>>(assignments, S, P, block_hist);
    MLB_CHECK_LAUNCH();
    reorder_scan_kernel<<<1, 1024, 2 * P * sizeof(int32_t), st>>>(block_hist, (int)nb, P, C, tab);
    MLB_CHECK_LAUNCH();
    reorder_scatter_kernel<<<(unsigned)nb, RO_BLOCK, (RO_BLOCK / 32) * P * sizeof(int32_t), st>>>(
        assignments, S, P, block_hist, tab, to_policy, to_sim, B * C);
    MLB_CHECK_LAUNCH();
    reorder_pad_kernel<<<mlb_cdiv(B * C, 256), 256, 0, st>>>(to_policy, B, C, (int32_t)S);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_gather_rows_clip(void* stream, const void* src, const int32_t* idx, void* out,
                                 long long n_idx, long long n_src, long long row_bytes) {
    if (n_idx == 0) return MLB_OK;
    MLB_REQUIRE(src && idx && out && n_idx > 0 && n_src > 0 && row_bytes > 0);
    cudaStream_t st = mlb_stream(stream);
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out);
    if (row_bytes % 16 == 0 && (al & 15) == 0) {
        const int rv = (int)(row_bytes / 16);
        gather_rows_clip_kernel<uint4><<<mlb_cdiv(n_idx * rv, 256), 256, 0, st>>>(
            static_cast<const uint4*>(src), idx, static_cast<uint4*>(out), n_idx, n_src, rv);
    } else if (row_bytes % 4 == 0 && (al & 3) == 0) {
        const int rv = (int)(row_bytes / 4);
        gather_rows_clip_kernel<uint32_t><<<mlb_cdiv(n_idx * rv, 256), 256, 0, st>>>(
            static_cast<const uint32_t*>(src), idx, static_cast<uint32_t*>(out), n_idx, n_src, rv);
    } else {
        gather_rows_clip_kernel<uint8_t><<<mlb_cdiv(n_idx * row_bytes, 256), 256, 0, st>>>(
            static_cast<const uint8_t*>(src), idx, static_cast<uint8_t*>(out), n_idx, n_src, (int)row_bytes);
    }
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
