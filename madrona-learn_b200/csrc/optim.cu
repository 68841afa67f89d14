// Fused multi-tensor optimiser over a flat fp32 parameter arena.
//
// Replaces optax.chain(clip_by_global_norm, adam) + apply_updates (ml/ppo.py:84-90,283-286),
// the kernel re-projection  w <- ||w0|| * w / ||w||  (ml/ppo.py:303-310, norms from
// ml/train_state.py:413-423) and the LayerNorm renormalisation
// (b, s) <- sqrt(F / (b.b + s.s)) * (b, s)  (ml/ppo.py:312-338).
// Three launches per step regardless of the number of parameter tensors:
//   mlb_sumsq_f32      global gradient norm (double accumulation, deterministic tree)
//   mlb_adam_step_f32  clip scale + Adam moments + parameter update (28 B/param)
//   mlb_renorm_segments one block per re-projected tensor / LayerNorm pair
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

struct SumPartial { double s; };

constexpr int OPT_BLOCK = 256;

unsigned opt_grid(long long n) {
    long long b = (n + OPT_BLOCK * 8 - 1) / (OPT_BLOCK * 8);
    const long long cap = (long long)MLB_NUM_SMS * 4;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

__global__ void __launch_bounds__(OPT_BLOCK)
sumsq_partial_kernel(const float* __restrict__ x, long long n, double* __restrict__ part) {
    double s = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = x[i];
        s += (double)v * (double)v;
    }
    __shared__ double smd[32];
    s = block_sum_d(s, smd);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(OPT_BLOCK)
sumsq_final_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nparts; b += blockDim.x) s += part[b];
    __shared__ double smd[32];
    s = block_sum_d(s, smd);
    if (threadIdx.x == 0) *out = s;
}

// optax.adam: m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; t += 1
//             u = -lr * (m / (1-b1^t)) / (sqrt(v / (1-b2^t)) + eps) ; p += u
__global__ void __launch_bounds__(OPT_BLOCK)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, const int* __restrict__ step,
            const double* __restrict__ grad_sumsq, float lr, float b1, float b2, float eps,
            float max_grad_norm, float grad_scale) {
    const int t = *step + 1;
    float scale = grad_scale;
    if (grad_sumsq && max_grad_norm > 0.f) {
        // the reduced gradient is grad_scale * g; its norm is grad_scale * sqrt(sumsq)
        const float gn = (float)(sqrt(*grad_sumsq) * (double)fabsf(grad_scale));
        if (!(gn < max_grad_norm)) scale *= max_grad_norm / gn;   // optax: (x / g_norm) * c
    }
    const float bc1 = 1.f - powf(b1, (float)t);
    const float bc2 = 1.f - powf(b2, (float)t);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * scale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float u = -lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
        p[i] += u;
    }
}

// One thread-block CLUSTER (8 CTAs, distributed shared memory) per segment:
//   kind 1 -> w *= target / ||w|| ; kind 2 -> (s,b) *= sqrt(target / (s.s+b.b)) ; kind 0 -> copy only
// Each CTA reduces its slice, the 8 partial sums are exchanged through DSMEM
// (cluster.map_shared_rank) after one cluster barrier, then every CTA rescales its slice and --
// when a bf16 copy table is given -- refreshes the tensor-core operand copies W (row-major) and
// W^T of the weight it just re-projected, so the optimiser step needs no separate cast pass.
constexpr int RENORM_CLUSTER = 8;

__global__ void __cluster_dims__(RENORM_CLUSTER, 1, 1) __launch_bounds__(512)
renorm_kernel(float* __restrict__ p, const mlb_segment* __restrict__ segs,
              const mlb_bf16_copy* __restrict__ copies, int* __restrict__ step) {
    cg::cluster_group cluster = cg::this_cluster();
    const int seg = blockIdx.x / RENORM_CLUSTER;
    const int rank = (int)cluster.block_rank();
    const mlb_segment sg = segs[seg];
    if (blockIdx.x == 0 && threadIdx.x == 0 && step) *step += 1;
    float* w = p + sg.offset;
    const long long per = (sg.length + RENORM_CLUSTER - 1) / RENORM_CLUSTER;
    const long long lo = rank * per, hi = min(sg.length, lo + per);
    __shared__ double smd[32];
    __shared__ double partial;
    __shared__ float fac;
    double s = 0.0;
    if (sg.kind != 0)
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) { const float v = w[i]; s += (double)v * v; }
    s = block_sum_d(s, smd);
    if (threadIdx.x == 0) partial = s;
    cluster.sync();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int r = 0; r < RENORM_CLUSTER; ++r) tot += *cluster.map_shared_rank(&partial, r);
        fac = sg.kind == 1 ? (float)((double)sg.target / sqrt(tot))
                           : (sg.kind == 2 ? (float)sqrt((double)sg.target / tot) : 1.f);
    }
    __syncthreads();
    const float f = fac;
    mlb_bf16_copy cp;
    cp.dst = nullptr; cp.dst_t = nullptr; cp.rows = cp.cols = cp.ld_t = cp.ld_d = 0;
    if (copies) cp = copies[seg];
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float v = w[i];
        if (sg.kind != 0) { v *= f; w[i] = v; }
        if (cp.dst_t) {
            const int r = (int)(i / cp.cols), c = (int)(i - (long long)r * cp.cols);
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            reinterpret_cast<__nv_bfloat16*>(cp.dst_t)[(long long)c * cp.ld_t + r] = h;
            if (cp.dst) reinterpret_cast<__nv_bfloat16*>(cp.dst)[(long long)r * cp.ld_d + c] = h;
        }
    }
    cluster.sync();      // keep every CTA's shared memory alive until all remote reads are done
}

// ------------------------------------------------------------------------------------------
// The whole optimiser step in ONE launch (the parameter arena is ~150 K floats: every phase of
// the three-launch path above is a latency-bound kernel).  The grid (<= #SMs blocks, all
// co-resident) owns contiguous slices of the arena and meets at two device-wide barriers:
//   A  sum(g^2) partials                                  -> barrier -> clip scale
//   B  Adam moments + update, per-(block, segment) sum(p^2) -> barrier
//   C  re-projection / LayerNorm renorm factor per segment (partials re-reduced in block order:
//      deterministic, so data-parallel ranks stay bit-identical), rescale, bf16 W / W^T refresh
// Segments must tile the arena exactly, in offset order.
// ------------------------------------------------------------------------------------------
constexpr int FO_BLOCK = 256;
constexpr int FO_MAX_SEGS = 32;

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// state[0] = arrival count, state[1] = generation.  Bounded spin: trap, never hang.
__device__ __forceinline__ void grid_barrier(uint32_t* state) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t gen = ld_acquire_gpu(&state[1]);
        if (atomicAdd(&state[0], 1u) == gridDim.x - 1) {
            state[0] = 0;
            __threadfence();
            st_release_gpu(&state[1], gen + 1);
        } else {
            const long long t0 = clock64();
            while (ld_acquire_gpu(&state[1]) == gen)
                if (clock64() - t0 > 4000000000ll) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FO_BLOCK)
optimizer_fused_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, const mlb_segment* __restrict__ segs, int nseg,
                       const mlb_bf16_copy* __restrict__ copies, int* __restrict__ step,
                       double* __restrict__ grad_sumsq, int have_sumsq, float lr, float b1, float b2,
                       float eps, float max_grad_norm, float grad_scale, uint32_t* __restrict__ sync_state,
                       double* __restrict__ gpart, double* __restrict__ segpart, float* __restrict__ zero_after) {
    __shared__ double smd[32];
    __shared__ double bcast;
    // No early launch_dependents here: this kernel rewrites the parameters and their bf16 operand
    // copies, which the tensor-core kernels further down the chain load in their PRE-wait prologues
    // (resident-weight TMA, LayerNorm scale/bias).  The trigger is issued at the very end, after the
    // last parameter write has been fenced, so no successor -- direct or transitive -- can start
    // before the new weights are in place.
    pdl_wait();
    const long long slice = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * slice, hi = min(n, lo + slice);
    // ---- A: global gradient norm --------------------------------------------------------
    double ssq;
    if (have_sumsq) ssq = *grad_sumsq;
    else {
        double s = 0.0;
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) { const float x = g[i]; s += (double)x * x; }
        s = block_sum_d(s, smd);
        if (threadIdx.x == 0) gpart[blockIdx.x] = s;
        grid_barrier(sync_state);
        double t = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += __ldcg(gpart + b);
        t = block_sum_d(t, smd);                 // same fixed order in every block
        if (threadIdx.x == 0) { bcast = t; if (blockIdx.x == 0) *grad_sumsq = t; }
        __syncthreads();
        ssq = bcast;
    }
    const int t = *step + 1;
    float scale = grad_scale;
    if (max_grad_norm > 0.f) {
        const float gn = (float)(sqrt(ssq) * (double)fabsf(grad_scale));
        if (!(gn < max_grad_norm)) scale *= max_grad_norm / gn;   // optax: (x / g_norm) * c
    }
    const float bc1 = 1.f - powf(b1, (float)t);
    const float bc2 = 1.f - powf(b2, (float)t);
    // ---- B: Adam + per-(block, segment) sum of squares of the updated parameters ---------
    int s0 = 0;
    while (s0 < nseg - 1 && segs[s0].offset + segs[s0].length <= lo) ++s0;
    for (int si = s0; si < nseg && segs[si].offset < hi; ++si) {
        const long long a = max(lo, (long long)segs[si].offset), e = min(hi, (long long)(segs[si].offset + segs[si].length));
        double s = 0.0;
        for (long long i = a + threadIdx.x; i < e; i += blockDim.x) {
            const float gi = g[i] * scale;
            const float mi = b1 * m[i] + (1.f - b1) * gi;
            const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
            m[i] = mi; v[i] = vi;
            const float pn = p[i] + (-lr * (mi / bc1) / (sqrtf(vi / bc2) + eps));
            p[i] = pn;
            s += (double)pn * pn;
        }
        s = block_sum_d(s, smd);
        if (threadIdx.x == 0) segpart[(long long)blockIdx.x * FO_MAX_SEGS + si] = s;
    }
    grid_barrier(sync_state);
    // ---- C: renorm factor per segment, rescale, bf16 operand copies -----------------------
    for (int si = s0; si < nseg && segs[si].offset < hi; ++si) {
        const mlb_segment sg = segs[si];
        const long long a = max(lo, (long long)sg.offset), e = min(hi, (long long)(sg.offset + sg.length));
        float f = 1.f;
        if (sg.kind != 0) {
            const int bfirst = (int)(sg.offset / slice), blast = (int)((sg.offset + sg.length - 1) / slice);
            double tot = 0.0;
            for (int b = bfirst + threadIdx.x; b <= blast; b += blockDim.x)
                tot += __ldcg(segpart + (long long)b * FO_MAX_SEGS + si);
            tot = block_sum_d(tot, smd);
            if (threadIdx.x == 0)
                bcast = sg.kind == 1 ? (double)sg.target / sqrt(tot) : sqrt((double)sg.target / tot);
            __syncthreads();
            f = (float)bcast;
            __syncthreads();
        }
        mlb_bf16_copy cp;
        cp.dst = nullptr; cp.dst_t = nullptr; cp.rows = cp.cols = cp.ld_t = cp.ld_d = 0;
        if (copies) cp = copies[si];
        if (sg.kind == 0 && !cp.dst_t) continue;
        for (long long i = a + threadIdx.x; i < e; i += blockDim.x) {
            float x = p[i];
            if (sg.kind != 0) { x *= f; p[i] = x; }
            if (cp.dst_t) {
                const long long k = i - sg.offset;
                const int r = (int)(k / cp.cols), c = (int)(k - (long long)r * cp.cols);
                const __nv_bfloat16 h = __float2bfloat16_rn(x);
                reinterpret_cast<__nv_bfloat16*>(cp.dst_t)[(long long)c * cp.ld_t + r] = h;
                if (cp.dst) reinterpret_cast<__nv_bfloat16*>(cp.dst)[(long long)r * cp.ld_d + c] = h;
            }
        }
    }
    // the gradient arena is consumed: clear this block's slice for the next minibatch's accumulation
    if (zero_after)
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) zero_after[i] = 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) *step = t;
    __threadfence();
    pdl_launch_dependents();
}

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long rows, int ld, int ncols, float* __restrict__ out) {
    // blockDim = (32 cols, 8 row-lanes); each block strides over rows
    const int c = blockIdx.y * 32 + threadIdx.x;
    float s = 0.f;
    if (c < ncols)
        for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y)
            s += ld_as_float(x + r * ld + c);
    __shared__ float sm[8][33];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < ncols) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

}  // namespace

MLB_API int mlb_fill_zero(void* stream, void* p, size_t bytes) {
    MLB_REQUIRE(p || bytes == 0);
    if (bytes == 0) return MLB_OK;
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, mlb_stream(stream));
    return e == cudaSuccess ? MLB_OK : (int)e;
}

// rows x cols block of a row-major fp32 matrix with row stride ld (elements): cudaMemset2DAsync, a memset node
// in a captured graph like mlb_fill_zero
MLB_API int mlb_fill_zero_2d(void* stream, float* p, int rows, int cols, int ld) {
    MLB_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols && (p || rows == 0 || cols == 0));
    if (rows == 0 || cols == 0) return MLB_OK;
    cudaError_t e = cudaMemset2DAsync(p, (size_t)ld * 4, 0, (size_t)cols * 4, (size_t)rows, mlb_stream(stream));
    return e == cudaSuccess ? MLB_OK : (int)e;
}

MLB_API int mlb_copy_bytes(void* stream, const void* src, void* dst, size_t bytes) {
    MLB_REQUIRE((src && dst) || bytes == 0);
    if (bytes == 0) return MLB_OK;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, mlb_stream(stream));
    return e == cudaSuccess ? MLB_OK : (int)e;
}

MLB_API size_t mlb_sumsq_workspace(long long n) { return (size_t)opt_grid(n) * sizeof(double); }

MLB_API int mlb_sumsq_f32(void* stream, const float* x, long long n, double* out, void* ws,
                          size_t ws_bytes) {
    MLB_REQUIRE(x && out && n > 0);
    const unsigned g = opt_grid(n);
    if (!ws || ws_bytes < g * sizeof(double)) return MLB_EWS;
    cudaStream_t s = mlb_stream(stream);
    sumsq_partial_kernel<<<g, OPT_BLOCK, 0, s>>>(x, n, reinterpret_cast<double*>(ws));
    MLB_CHECK_LAUNCH();
    sumsq_final_kernel<<<1, OPT_BLOCK, 0, s>>>(reinterpret_cast<double*>(ws), (int)g, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_adam_step_f32(void* stream, float* params, const float* grads, float* m, float* v,
                              long long n, const int32_t* step, const double* grad_sumsq,
                              float lr, float b1, float b2, float eps, float max_grad_norm,
                              float grad_scale) {
    MLB_REQUIRE(params && grads && m && v && step && n > 0);
    adam_kernel<<<opt_grid(n), OPT_BLOCK, 0, mlb_stream(stream)>>>(params, grads, m, v, n, step,
        grad_sumsq, lr, b1, b2, eps, max_grad_norm, grad_scale);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_renorm_segments(void* stream, float* params, const mlb_segment* segments_dev,
                                int num_segments, int32_t* step, const mlb_bf16_copy* copies_dev) {
    MLB_REQUIRE(params && segments_dev && num_segments > 0);
    renorm_kernel<<<num_segments * RENORM_CLUSTER, 512, 0, mlb_stream(stream)>>>(params, segments_dev,
                                                                                copies_dev, step);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API size_t mlb_optimizer_fused_workspace(void) {
    return (size_t)MLB_NUM_SMS * (1 + FO_MAX_SEGS) * sizeof(double) + 16;
}

MLB_API int mlb_optimizer_step_fused(void* stream, float* params, const float* grads, float* m, float* v,
                                     long long n, const mlb_segment* segments_dev, int num_segments,
                                     const mlb_bf16_copy* copies_dev, int32_t* step, double* grad_sumsq,
                                     int have_sumsq, float lr, float b1, float b2, float eps,
                                     float max_grad_norm, float grad_scale, uint32_t* sync_state, void* ws,
                                     size_t ws_bytes, float* zero_after) {
    MLB_REQUIRE(params && grads && m && v && step && grad_sumsq && segments_dev && sync_state && n > 0);
    MLB_REQUIRE(num_segments > 0 && num_segments <= FO_MAX_SEGS);
    if (!ws || ws_bytes < mlb_optimizer_fused_workspace()) return MLB_EWS;
    long long gsz = (n + 1023) / 1024;
    if (gsz > MLB_NUM_SMS) gsz = MLB_NUM_SMS;
    static const int cap = mlb_coresident_cap(optimizer_fused_kernel, FO_BLOCK, 0);      // two grid barriers inside
    if (gsz > cap) gsz = cap;
    if (gsz < 1) gsz = 1;
    double* gpart = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 15) & ~uintptr_t(15));
    double* segpart = gpart + MLB_NUM_SMS;
    cudaError_t e = launch_pdl(optimizer_fused_kernel, dim3((unsigned)gsz), dim3(FO_BLOCK), 0, mlb_stream(stream),
        params, grads, m, v, n, segments_dev, num_segments, copies_dev, step, grad_sumsq, have_sumsq, lr, b1, b2,
        eps, max_grad_norm, grad_scale, sync_state, gpart, segpart, zero_after);
    if (e != cudaSuccess) return (int)e;
    return MLB_OK;
}

MLB_API int mlb_colsum_f32(void* stream, const float* x, long long rows, int ld, int ncols,
                           float* out) {
    MLB_REQUIRE(x && out && rows >= 0 && ncols > 0 && ld >= ncols);
    if (rows == 0) return MLB_OK;
    long long g = (rows + 63) / 64;
    if (g > MLB_NUM_SMS * 4) g = MLB_NUM_SMS * 4;
    colsum_kernel<float><<<dim3((unsigned)g, (unsigned)((ncols + 31) / 32)), dim3(32, 8), 0, mlb_stream(stream)>>>(x, rows, ld, ncols, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_colsum_bf16(void* stream, const void* x, long long rows, int ld, int ncols, float* out) {
    MLB_REQUIRE(x && out && rows >= 0 && ncols > 0 && ld >= ncols);
    if (rows == 0) return MLB_OK;
    long long g = (rows + 63) / 64;
    if (g > MLB_NUM_SMS * 4) g = MLB_NUM_SMS * 4;
    colsum_kernel<__nv_bfloat16><<<dim3((unsigned)g, (unsigned)((ncols + 31) / 32)), dim3(32, 8), 0, mlb_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), rows, ld, ncols, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
