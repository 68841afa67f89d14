// Shared device/host helpers for libmlb200 (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mlb200.h"

#define MLB_API extern "C" __attribute__((visibility("default")))

// Every entry point: enqueue-only, no sync, no allocation; returns 0, a cudaError_t (>0),
// or a negative MLB_E* code.
#define MLB_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return (int)e__; } while (0)

#define MLB_REQUIRE(cond) do { if (!(cond)) return MLB_EINVAL; } while (0)

static inline cudaStream_t mlb_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: a kernel launched with launch_pdl() may start (prologue: barrier
// init, TMEM allocation, tensor-map prefetch, resident-weight load) while its predecessor in the
// stream is still draining its last tiles; pdl_wait() blocks until the predecessor has completed
// and flushed, and must precede every access to data the predecessor (or anything before it)
// produced.  pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as soon as every
// CTA of this grid has started.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("MLB_PDL"); return !(v && v[0] == '0'); }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}



// Kernels that meet at device-wide spin barriers need every CTA of the grid resident at once.  They are launched
// non-cooperatively (inside PDL chains / CUDA graphs), so the grid is capped by what the CURRENT device can hold:
// its queried SM count (not the compile-time MLB_NUM_SMS) x the occupancy of that kernel.  Cached per process
// (one process per GPU).  A concurrent kernel on another stream can still delay a CTA; the barriers' bounded
// spins then trap instead of hanging.
template <typename Kern>
inline int mlb_coresident_cap(Kern kern, int block, size_t smem) {
    int dev = 0, sms = 0, per = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) return 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, block, smem) != cudaSuccess || per < 1) return 1;
    return sms * per;
}

static inline unsigned mlb_cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

constexpr int MLB_NUM_SMS = 148;   // B200: 2 dies x 74 SMs

static inline bool mlb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------
// streaming (evict-first) 128-bit / 32-bit global accesses
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
    return __ldcs(reinterpret_cast<const unsigned int*>(p));
}

// ---------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of a double; result valid in thread 0.  `sm` needs >= 32 doubles.
__device__ __forceinline__ double block_sum_d(double v, double* sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sm[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (w == 0) {
        r = lane < nw ? sm[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_min_f(float v, float* sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_min(v);
    if (lane == 0) sm[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = v;
    if (w == 0) {
        r = lane < nw ? sm[lane] : INFINITY;
        r = warp_min(r);
    }
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_max_f(float v, float* sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) sm[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = v;
    if (w == 0) {
        r = lane < nw ? sm[lane] : -INFINITY;
        r = warp_max(r);
    }
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------
// Threefry-2x32 (20 rounds) -- JAX's PRNG core (published algorithm, Random123).
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mlb_rotl32(uint32_t x, int r) {
    return (x << r) | (x >> (32 - r));
}
__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1,
                                                      uint32_t& x0, uint32_t& x1) {
    const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    const int R0[4] = {13, 15, 26, 6};
    const int R1[4] = {17, 29, 16, 24};
    x0 += ks[0];
    x1 += ks[1];
#pragma unroll
    for (int g = 0; g < 5; ++g) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (g & 1) ? R1[i] : R0[i];
            x0 += x1;
            x1 = mlb_rotl32(x1, r);
            x1 ^= x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
}

// random_bits(key, 32, shape)[idx] for a flat array of `size` elements.
// partitionable == 0: jax < 0.5 layout (counters iota(size) padded to even, split in halves)
// partitionable == 1: jax >= 0.5 layout (64-bit iota as (hi, lo); bits = y0 ^ y1)
__host__ __device__ __forceinline__ uint32_t threefry_bits_at(uint32_t k0, uint32_t k1,
                                                              uint64_t idx, uint64_t size,
                                                              int partitionable) {
    if (partitionable) {
        uint32_t x0 = (uint32_t)(idx >> 32), x1 = (uint32_t)idx;
        threefry2x32(k0, k1, x0, x1);
        return x0 ^ x1;
    }
    const uint64_t padded = size + (size & 1);
    const uint64_t half = padded >> 1;
    uint32_t x0, x1;
    if (idx < half) {
        x0 = (uint32_t)idx;
        const uint64_t j = idx + half;
        x1 = j < size ? (uint32_t)j : 0u;     // the pad element is counter 0
        threefry2x32(k0, k1, x0, x1);
        return x0;
    }
    x0 = (uint32_t)(idx - half);
    x1 = (uint32_t)idx;
    threefry2x32(k0, k1, x0, x1);
    return x1;
}

// random.split(key, num)[i] -> (o0, o1)
__host__ __device__ __forceinline__ void threefry_split_at(uint32_t k0, uint32_t k1, uint32_t i,
                                                           uint32_t num, int partitionable,
                                                           uint32_t& o0, uint32_t& o1) {
    if (partitionable) {
        uint32_t x0 = 0, x1 = i;
        threefry2x32(k0, k1, x0, x1);
        o0 = x0; o1 = x1;
        return;
    }
    o0 = threefry_bits_at(k0, k1, 2ull * i, 2ull * num, 0);
    o1 = threefry_bits_at(k0, k1, 2ull * i + 1, 2ull * num, 0);
}
