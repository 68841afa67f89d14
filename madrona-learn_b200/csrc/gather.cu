// K5: minibatch gather straight from the rollout store (bit-exact indexing).
//
// The reference first relayouts the whole store [C, T', P, B, ...] -> [P, C*B, T', ...]
// (ml/rollouts.py:788-804), then per minibatch does take(axis 0) + swapaxes(0, 1)
// (ml/rollouts.py:319-329).  With trajectory j = c*B + b both steps collapse to
//     out[s, m, :] = store[j/B, s, j%B, :],  j = idx[m]      (P = 1)
// so the relayout copy never exists.  Rows are copied as 16-byte vectors when the row size
// and pointers allow it; a warp covers consecutive bytes of consecutive rows of one time
// step, so global stores are fully coalesced and loads are coalesced within each row.
#include <cuda_bf16.h>

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

// All leaves of a minibatch in ONE launch.  Per leaf the vector width (16/8/4/1 bytes) is chosen on
// the host; the observation leaf can additionally be emitted as bf16 (the tensor-core forward's
// A operand), which replaces a separate cast pass over the minibatch.
constexpr int MAX_LEAVES = 8;
struct LeafDev {
    const uint8_t* store;
    uint8_t* out;
    __nv_bfloat16* out_bf16;
    int row_vecs, vec_bytes;
};
struct LeafPack { LeafDev l[MAX_LEAVES]; int n; };

__global__ void __launch_bounds__(256)
gather_multi_kernel(const __grid_constant__ LeafPack P, const int32_t* __restrict__ idx, int Tp, int B, int M) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = blockIdx.y;
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int li = 0; li < P.n; ++li) {
        const LeafDev& L = P.l[li];
        const int rv = L.row_vecs, total = M * rv;
        for (int e = t0; e < total; e += stride) {
            const int m = e / rv, k = e - m * rv;
            const int j = idx[m];
            const int c = j / B, b = j - c * B;
            const long long src = (((long long)c * Tp + s) * B + b) * rv + k;
            const long long dst = ((long long)s * M + m) * rv + k;
            switch (L.vec_bytes) {
                case 16: {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(L.store) + src);
                    if (L.out) reinterpret_cast<uint4*>(L.out)[dst] = v;
                    if (L.out_bf16) {
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(__uint_as_float(v.x), __uint_as_float(v.y));
                        const __nv_bfloat162 hi = __floats2bfloat162_rn(__uint_as_float(v.z), __uint_as_float(v.w));
                        reinterpret_cast<uint2*>(L.out_bf16)[dst] =
                            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                    }
                    break;
                }
                case 8: reinterpret_cast<uint2*>(L.out)[dst] = __ldg(reinterpret_cast<const uint2*>(L.store) + src); break;
                case 4: reinterpret_cast<uint32_t*>(L.out)[dst] = __ldg(reinterpret_cast<const uint32_t*>(L.store) + src); break;
                default: L.out[dst] = __ldg(L.store + src); break;
            }
        }
    }
}

// Index-exact data-parallel minibatch gather (SURVEY 8e "index-exact"): idx holds GLOBAL trajectory
// ids j = c*Bg + bg of the world-sharded run (Bg = world * B); worlds [r*B, (r+1)*B) live in rank r's
// store, which is mapped into this process over NVLink (symmetric memory).  Rows owned by a peer
// are fetched by plain peer loads: >= 256-byte observation rows, so the NVLink requests are
// full-width.  Same addressing as gather_multi_kernel once (owner, c, b) are known.
constexpr int MAX_PEERS = 8;
struct PeerStores { const uint8_t* base[MAX_PEERS][MAX_LEAVES]; int world; };

__global__ void __launch_bounds__(256)
gather_multi_peer_kernel(const __grid_constant__ LeafPack P, const __grid_constant__ PeerStores S,
                         const int32_t* __restrict__ idx, int Tp, int B, int M) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = blockIdx.y;
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int Bg = B * S.world;
    for (int li = 0; li < P.n; ++li) {
        const LeafDev& L = P.l[li];
        const int rv = L.row_vecs, total = M * rv;
        for (int e = t0; e < total; e += stride) {
            const int m = e / rv, k = e - m * rv;
            const int j = idx[m];
            const int c = j / Bg, bg = j - c * Bg;
            const int owner = bg / B, b = bg - owner * B;
            const uint8_t* store = S.base[owner][li];
            const long long src = (((long long)c * Tp + s) * B + b) * rv + k;
            const long long dst = ((long long)s * M + m) * rv + k;
            switch (L.vec_bytes) {
                case 16: {
                    const uint4 v = *(reinterpret_cast<const uint4*>(store) + src);
                    if (L.out) reinterpret_cast<uint4*>(L.out)[dst] = v;
                    if (L.out_bf16) {
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(__uint_as_float(v.x), __uint_as_float(v.y));
                        const __nv_bfloat162 hi = __floats2bfloat162_rn(__uint_as_float(v.z), __uint_as_float(v.w));
                        reinterpret_cast<uint2*>(L.out_bf16)[dst] =
                            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                    }
                    break;
                }
                case 8: reinterpret_cast<uint2*>(L.out)[dst] = *(reinterpret_cast<const uint2*>(store) + src); break;
                case 4: reinterpret_cast<uint32_t*>(L.out)[dst] = *(reinterpret_cast<const uint32_t*>(store) + src); break;
                default: L.out[dst] = store[src]; break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Owner-affine split of the global minibatches (index-exact data-parallel mode).
//
// Global minibatch (e, k) = perm[e, k*Mp : (k+1)*Mp] (Mp = world * M trajectory ids, identical on every
// rank).  WHICH rank trains on which of its trajectories does not change the update (the loss is a sum
// over the minibatch, its statistics are global), so instead of handing rank r the r-th contiguous slice
// -- (world-1)/world of whose rows live on other GPUs -- every rank keeps the trajectories it OWNS (in
// permutation order, at most M) and only the binomial imbalance moves: the surplus of the over-represented
// ranks, in rank order, fills the deficits of the under-represented ones, in rank order.  Every rank
// evaluates the same assignment and extracts its own M ids, so the union over ranks is exactly the
// reference's minibatch.  Expected remote fraction ~ sqrt((world-1) / (2 pi M)) instead of (world-1)/world.
// One block per (epoch, minibatch); owner(j) = (j mod (world*B)) / B.
// ------------------------------------------------------------------------------------------
constexpr int ASSIGN_THREADS = 1024;

__global__ void __launch_bounds__(ASSIGN_THREADS)
dp_assign_kernel(const int32_t* __restrict__ perm, long long perm_ld, int nmb, int world, int rank, int B, int M,
                 int32_t* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int wsum[32][MLB_MAX_PEERS];          // per-warp totals per owner
    __shared__ int cnt[MLB_MAX_PEERS], sur_pre[MLB_MAX_PEERS + 1], def_pre[MLB_MAX_PEERS + 1];
    const int e = blockIdx.x / nmb, k = blockIdx.x - e * nmb;
    const int Mp = world * M;
    const int32_t* ids = perm + (long long)e * perm_ld + (long long)k * Mp;
    int32_t* o = out + ((long long)e * nmb + k) * M;
    const int ipt = (Mp + ASSIGN_THREADS - 1) / ASSIGN_THREADS;         // ids per thread (contiguous: stable)
    const int lo = threadIdx.x * ipt, hi = min(Mp, lo + ipt);
    const int Bg = world * B;
    int local[MLB_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < MLB_MAX_PEERS; ++r) local[r] = 0;
    for (int i = lo; i < hi; ++i) {
        const int ow = (ids[i] % Bg) / B;
#pragma unroll
        for (int r = 0; r < MLB_MAX_PEERS; ++r) local[r] += (ow == r);
    }
    // exclusive scan over threads, per owner: warp shuffles, then the warp totals through shared memory
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int pre[MLB_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < MLB_MAX_PEERS; ++r) {
        int v = local[r];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += t;
        }
        pre[r] = v - local[r];
        if (lane == 31) wsum[warp][r] = v;
    }
    __syncthreads();
    if (threadIdx.x < MLB_MAX_PEERS) {
        int run = 0;
        for (int w = 0; w < ASSIGN_THREADS / 32; ++w) { const int t = wsum[w][threadIdx.x]; wsum[w][threadIdx.x] = run; run += t; }
        cnt[threadIdx.x] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, d = 0;
        for (int r = 0; r < world; ++r) {
            sur_pre[r] = a; def_pre[r] = d;
            a += max(cnt[r] - M, 0);
            d += max(M - cnt[r], 0);
        }
        sur_pre[world] = a; def_pre[world] = d;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MLB_MAX_PEERS; ++r) pre[r] += wsum[warp][r];
    for (int i = lo; i < hi; ++i) {
        const int id = ids[i];
        const int ow = (id % Bg) / B;
        int p = 0;
#pragma unroll
        for (int r = 0; r < MLB_MAX_PEERS; ++r) if (ow == r) { p = pre[r]; pre[r] += 1; }
        if (p < M) {
            if (ow == rank) o[p] = id;                              // kept by its owner
        } else {
            const int q = sur_pre[ow] + (p - M);                    // q-th element of the global surplus list
            int dst = 0;
            while (dst + 1 < world && def_pre[dst + 1] <= q) ++dst; // the deficit rank it goes to
            if (dst == rank) o[min(cnt[dst], M) + (q - def_pre[dst])] = id;
        }
    }
}

}  // namespace

MLB_API int mlb_mb_gather_multi(void* stream, const mlb_gather_leaf* leaves_host, int num_leaves,
                                const int32_t* idx, int C, int Tp, long long B, long long M) {
    if (M == 0 || num_leaves == 0) return MLB_OK;
    MLB_REQUIRE(leaves_host && idx && num_leaves > 0 && num_leaves <= MAX_LEAVES && C > 0 && Tp > 0 &&
                Tp <= 65535 && B > 0 && M > 0 && B < (1ll << 31) && (long long)C * B < (1ll << 31));
    LeafPack P;
    P.n = num_leaves;
    long long max_total = 0;
    for (int i = 0; i < num_leaves; ++i) {
        const mlb_gather_leaf& h = leaves_host[i];
        MLB_REQUIRE(h.store && (h.out || h.out_bf16) && h.row_bytes > 0);
        const uintptr_t a = reinterpret_cast<uintptr_t>(h.store) | reinterpret_cast<uintptr_t>(h.out);
        int vb = 1;
        if (h.row_bytes % 16 == 0 && (a & 15) == 0) vb = 16;
        else if (h.row_bytes % 8 == 0 && (a & 7) == 0) vb = 8;
        else if (h.row_bytes % 4 == 0 && (a & 3) == 0) vb = 4;
        MLB_REQUIRE(!h.out_bf16 || (vb == 16 && (reinterpret_cast<uintptr_t>(h.out_bf16) & 7) == 0));
        MLB_REQUIRE(h.out || vb == 16);
        const long long rv = h.row_bytes / vb;
        MLB_REQUIRE(M * rv < (1ll << 31));
        P.l[i] = LeafDev{static_cast<const uint8_t*>(h.store), static_cast<uint8_t*>(h.out),
                         static_cast<__nv_bfloat16*>(h.out_bf16), (int)rv, vb};
        if (M * rv > max_total) max_total = M * rv;
    }
    long long gx = (max_total + 255) / 256;
    const long long cap = (long long)MLB_NUM_SMS * 8 / (Tp < 8 ? Tp : 8) + 1;
    if (gx > cap) gx = cap;
    cudaError_t e = launch_pdl(gather_multi_kernel, dim3((unsigned)gx, (unsigned)Tp), dim3(256), 0, mlb_stream(stream), P, idx,
                               Tp, (int)B, (int)M);
    if (e != cudaSuccess) return (int)e;
    return MLB_OK;
}

MLB_API int mlb_mb_gather_multi_peer(void* stream, const mlb_gather_leaf* leaves_host, int num_leaves,
                                     const void* const* peer_stores_host, int world, const int32_t* idx,
                                     int C, int Tp, long long B, long long M) {
    if (M == 0 || num_leaves == 0) return MLB_OK;
    MLB_REQUIRE(leaves_host && idx && peer_stores_host && num_leaves > 0 && num_leaves <= MAX_LEAVES && C > 0 &&
                Tp > 0 && Tp <= 65535 && B > 0 && M > 0 && world >= 1 && world <= MAX_PEERS &&
                (long long)C * B * world < (1ll << 31));
    LeafPack P;
    PeerStores S;
    P.n = num_leaves;
    S.world = world;
    long long max_total = 0;
    for (int i = 0; i < num_leaves; ++i) {
        const mlb_gather_leaf& h = leaves_host[i];
        MLB_REQUIRE((h.out || h.out_bf16) && h.row_bytes > 0);
        uintptr_t a = reinterpret_cast<uintptr_t>(h.out);
        for (int r = 0; r < world; ++r) {
            const void* st = peer_stores_host[(size_t)r * num_leaves + i];
            MLB_REQUIRE(st);
            a |= reinterpret_cast<uintptr_t>(st);
            S.base[r][i] = static_cast<const uint8_t*>(st);
        }
        int vb = 1;
        if (h.row_bytes % 16 == 0 && (a & 15) == 0) vb = 16;
        else if (h.row_bytes % 8 == 0 && (a & 7) == 0) vb = 8;
        else if (h.row_bytes % 4 == 0 && (a & 3) == 0) vb = 4;
        MLB_REQUIRE(!h.out_bf16 || (vb == 16 && (reinterpret_cast<uintptr_t>(h.out_bf16) & 7) == 0));
        MLB_REQUIRE(h.out || vb == 16);
        const long long rv = h.row_bytes / vb;
        MLB_REQUIRE(M * rv < (1ll << 31));
        P.l[i] = LeafDev{nullptr, static_cast<uint8_t*>(h.out), static_cast<__nv_bfloat16*>(h.out_bf16), (int)rv, vb};
        if (M * rv > max_total) max_total = M * rv;
    }
    long long gx = (max_total + 255) / 256;
    const long long cap = (long long)MLB_NUM_SMS * 8 / (Tp < 8 ? Tp : 8) + 1;
    if (gx > cap) gx = cap;
    cudaError_t e = launch_pdl(gather_multi_peer_kernel, dim3((unsigned)gx, (unsigned)Tp), dim3(256), 0,
                               mlb_stream(stream), P, S, idx, Tp, (int)B, (int)M);
    if (e != cudaSuccess) return (int)e;
    return MLB_OK;
}

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

// out [E, nmb, M]: this rank's trajectory ids of every global minibatch (see dp_assign_kernel).
MLB_API int mlb_dp_assign_minibatches(void* stream, const int32_t* perm, long long perm_ld, int E, int nmb,
                                      int world, int rank, long long B, long long M, int32_t* out) {
    MLB_REQUIRE(perm && out && E > 0 && nmb > 0 && world >= 1 && world <= MLB_MAX_PEERS && rank >= 0 && rank < world);
    MLB_REQUIRE(B > 0 && M > 0 && (long long)world * M < (1ll << 30) && (long long)world * B < (1ll << 31) &&
                perm_ld >= (long long)nmb * world * M);
    cudaError_t e = launch_pdl(dp_assign_kernel, dim3((unsigned)(E * nmb)), dim3(ASSIGN_THREADS), 0, mlb_stream(stream),
                               perm, perm_ld, nmb, world, rank, (int)B, (int)M, out);
    return e == cudaSuccess ? MLB_OK : (int)e;
}
