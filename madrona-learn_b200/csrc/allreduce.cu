// Gradient all-reduce fused with the global-norm reduction, over NVLink peer memory.
//
// Data-parallel PPO sums the flat gradient arena (~150 K floats for the 3x256 MLP)
// over ranks once per minibatch, then clip_by_global_norm needs sum(g^2) of the reduced
// gradient (ml/ppo.py:84-90).  At this size an NCCL all-reduce is pure latency (19 us at 2 GPUs,
// 38 us at 8, measured in the update graph) and is followed by two more latency-bound launches
// for the norm.  Here every rank's arena lives in symmetric memory (peer-mapped over
// NVLink/NVSwitch); ONE kernel per rank
//   1. signals "my gradients are complete" into every peer's signal pad and waits for all peers,
//   2. one-shot reduces: reads every rank's arena with system-scope loads in RANK ORDER (so all
//      ranks compute bit-identical sums), writes the reduced gradient to a local buffer and
//      accumulates sum(g^2) in fp64,
//   3. the last block to finish publishes sum(g^2), signals "done reading" to every peer and
//      waits for theirs, so the arena may be overwritten as soon as the kernel completes.
// Signals carry a monotonically increasing epoch kept on the device, so the kernel can be
// replayed from a CUDA graph.  All waits are bounded (trap, never hang the GPU).
#include "common.cuh"

namespace {

constexpr int AR_BLOCK = 256;
constexpr int AR_MAX_GRID = 128;

struct PeerTable {
    int rank, world;
    const float* grads[MLB_MAX_PEERS];
    uint32_t* signals[MLB_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// epochs are compared with wrap-around arithmetic
__device__ __forceinline__ void wait_epoch(const uint32_t* slot, uint32_t e) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(slot) - e) < 0) {
        if (clock64() - t0 > 60000000000ll) __trap();         // ~30 s: a peer never arrived
    }
}

// state[0] = epoch of the last completed call, state[1] = block arrival counter
__global__ void __launch_bounds__(AR_BLOCK)
allreduce_sumsq_kernel(const __grid_constant__ PeerTable T, float* __restrict__ out, long long n,
                       double* __restrict__ sumsq_out, uint32_t* __restrict__ state,
                       double* __restrict__ partials) {
    const uint32_t e = state[0] + 1;           // every block reads it before the last block advances it
    uint32_t* my_sig = T.signals[T.rank];
    if (blockIdx.x == 0 && threadIdx.x < T.world) {
        __threadfence_system();
        st_release_sys(T.signals[threadIdx.x] + T.rank, e);                 // phase 0: data ready
    }
    if (threadIdx.x < T.world) wait_epoch(my_sig + threadIdx.x, e);
    __syncthreads();

    double acc = 0.0;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        // all peers' loads are issued before the first add (one NVLink round trip, not `world`);
        // the sum itself stays in rank order
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < T.world; r0 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (r0 + u < T.world) v[u] = ld_sys_v4(T.grads[r0 + u] + 4 * i);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (r0 + u < T.world) {
                    if (r0 + u == 0) s = v[u];
                    else { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
                }
        }
        reinterpret_cast<float4*>(out)[i] = s;
        acc += (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z + (double)s.w * s.w;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s = ld_sys_f32(T.grads[0] + i);
        for (int r = 1; r < T.world; ++r) s += ld_sys_f32(T.grads[r] + i);
        out[i] = s;
        acc += (double)s * s;
    }
    __shared__ double smd[32];
    __shared__ bool last;
    acc = block_sum_d(acc, smd);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc;
        __threadfence();
        last = atomicAdd(&state[1], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    // last block of this rank: every block has finished reading the peers' arenas
    if (threadIdx.x < T.world) {
        __threadfence_system();
        st_release_sys(T.signals[threadIdx.x] + MLB_MAX_PEERS + T.rank, e);   // phase 1: done reading
    }
    double tot = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) tot += __ldcg(partials + b);
    tot = block_sum_d(tot, smd);
    if (threadIdx.x < T.world) wait_epoch(my_sig + MLB_MAX_PEERS + threadIdx.x, e);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sumsq_out) *sumsq_out = tot;
        state[1] = 0;
        state[0] = e;
    }
}

// ------------------------------------------------------------------------------------------
// NVLS variant: the NVSwitch reduces.  Rank r owns slice r of the arena: one
// multimem.ld_reduce per 16 bytes returns the sum over all ranks computed in the switch, and one
// multimem.st broadcasts it into every rank's `reduced` buffer -- each element is reduced exactly
// once, so all ranks see identical bits (the scheme of NCCL's NVLS all-reduce), with 1/world of
// the one-shot variant's NVLink traffic.  Two cross-rank barriers: gradients ready / slices
// written (the second also means nobody reads the arenas any more).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mc_ld_reduce_v4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_v4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float mc_ld_reduce_f32(const float* mc) {
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f32(float* mc, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}

// state[0] = epoch, state[1], state[2] = block arrival counters.  All blocks must be co-resident
// (grid <= AR_MAX_GRID < #SMs): they wait for the peers' slices inside the kernel.
__global__ void __launch_bounds__(AR_BLOCK)
allreduce_nvls_kernel(const __grid_constant__ PeerTable T, const float* __restrict__ mc_grads,
                      float* __restrict__ mc_out, const float* __restrict__ out_local, long long n,
                      double* __restrict__ sumsq_out, uint32_t* __restrict__ state,
                      double* __restrict__ partials) {
    const uint32_t e = state[0] + 1;
    uint32_t* my_sig = T.signals[T.rank];
    if (blockIdx.x == 0 && threadIdx.x < T.world) {
        __threadfence_system();
        st_release_sys(T.signals[threadIdx.x] + T.rank, e);                 // phase 0: gradients ready
    }
    if (threadIdx.x < T.world) wait_epoch(my_sig + threadIdx.x, e);
    __syncthreads();

    const long long n4 = n >> 2;
    const long long per = (n4 + T.world - 1) / T.world;
    const long long lo = T.rank * per, hi = min(n4, lo + per);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = lo + t0; i < hi; i += stride) mc_st_v4(mc_out + 4 * i, mc_ld_reduce_v4(mc_grads + 4 * i));
    if (T.rank == 0)
        for (long long i = 4 * n4 + t0; i < n; i += stride) mc_st_f32(mc_out + i, mc_ld_reduce_f32(mc_grads + i));

    __shared__ double smd[32];
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(&state[1], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < T.world)                                       // phase 1: my slice is everywhere
        st_release_sys(T.signals[threadIdx.x] + MLB_MAX_PEERS + T.rank, e);
    if (threadIdx.x < T.world) wait_epoch(my_sig + MLB_MAX_PEERS + threadIdx.x, e);
    __syncthreads();

    double acc = 0.0;
    for (long long i = t0; i < n4; i += stride) {
        const float4 v = ld_sys_v4(out_local + 4 * i);
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (long long i = 4 * n4 + t0; i < n; i += stride) { const float v = ld_sys_f32(out_local + i); acc += (double)v * v; }
    acc = block_sum_d(acc, smd);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc;
        __threadfence();
        last = atomicAdd(&state[2], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    double tot = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) tot += __ldcg(partials + b);
    tot = block_sum_d(tot, smd);
    if (threadIdx.x == 0) {
        if (sumsq_out) *sumsq_out = tot;
        state[1] = 0;
        state[2] = 0;
        state[0] = e;
    }
}

}  // namespace

MLB_API int mlb_allreduce_nvls_f32(void* stream, const mlb_peer_table* peers_host, const float* mc_grads,
                                   float* mc_out, const float* out_local, long long n, double* sumsq_out,
                                   uint32_t* state, void* ws, size_t ws_bytes) {
    MLB_REQUIRE(peers_host && mc_grads && mc_out && out_local && state && n > 0 && ws &&
                ws_bytes >= AR_MAX_GRID * sizeof(double));
    MLB_REQUIRE(peers_host->world >= 1 && peers_host->world <= MLB_MAX_PEERS && peers_host->rank >= 0 &&
                peers_host->rank < peers_host->world && mlb_aligned16(mc_grads) && mlb_aligned16(mc_out) &&
                mlb_aligned16(out_local));
    PeerTable T;
    T.rank = peers_host->rank;
    T.world = peers_host->world;
    for (int r = 0; r < T.world; ++r) {
        MLB_REQUIRE(peers_host->signals[r]);
        T.grads[r] = peers_host->grads[r];
        T.signals[r] = peers_host->signals[r];
    }
    long long g = mlb_cdiv(n >> 2, AR_BLOCK);
    if (g > AR_MAX_GRID) g = AR_MAX_GRID;
    static const int cap_nvls = mlb_coresident_cap(allreduce_nvls_kernel, AR_BLOCK, 0);  // grid-wide tickets inside
    if (g > cap_nvls) g = cap_nvls;
    if (g < 1) g = 1;
    allreduce_nvls_kernel<<<(unsigned)g, AR_BLOCK, 0, mlb_stream(stream)>>>(
        T, mc_grads, mc_out, out_local, n, sumsq_out, state, static_cast<double*>(ws));
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API size_t mlb_allreduce_workspace(void) { return AR_MAX_GRID * sizeof(double); }

MLB_API int mlb_allreduce_sumsq_f32(void* stream, const mlb_peer_table* peers_host, float* out,
                                    long long n, double* sumsq_out, uint32_t* state, void* ws,
                                    size_t ws_bytes) {
    MLB_REQUIRE(peers_host && out && state && n > 0 && ws && ws_bytes >= AR_MAX_GRID * sizeof(double));
    MLB_REQUIRE(peers_host->world >= 1 && peers_host->world <= MLB_MAX_PEERS && peers_host->rank >= 0 &&
                peers_host->rank < peers_host->world && mlb_aligned16(out));
    PeerTable T;
    T.rank = peers_host->rank;
    T.world = peers_host->world;
    for (int r = 0; r < T.world; ++r) {
        MLB_REQUIRE(peers_host->grads[r] && peers_host->signals[r] && mlb_aligned16(peers_host->grads[r]));
        T.grads[r] = peers_host->grads[r];
        T.signals[r] = peers_host->signals[r];
    }
    long long g = mlb_cdiv(n >> 2, AR_BLOCK);
    if (g > AR_MAX_GRID) g = AR_MAX_GRID;
    static const int cap_p2p = mlb_coresident_cap(allreduce_sumsq_kernel, AR_BLOCK, 0);
    if (g > cap_p2p) g = cap_p2p;
    if (g < 1) g = 1;
    allreduce_sumsq_kernel<<<(unsigned)g, AR_BLOCK, 0, mlb_stream(stream)>>>(
        T, out, n, sumsq_out, state, static_cast<double*>(ws));
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
