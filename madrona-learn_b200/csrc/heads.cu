// Actor/critic head epilogues: discrete action sampling (rollout) and the fused PPO loss +
// gradient w.r.t. the head outputs (learner).
//
// Replaces DiscreteActionDistributions.sample / action_stats (ml/dists.py:26-77), the
// rollout PRNG key chain (ml/rollouts.py:878-880), and _ppo_update's loss_fn
// (ml/ppo.py:129-262: z-scored advantages, clipped surrogate, value loss with optional clip /
// huber / value-normaliser, entropy bonus) together with its autodiff down to the head
// outputs.  `head` is the [rows, ld] output of the fused actor+critic head GEMM: columns
// [0, sumA) are the concatenated logits, column sumA is the critic value.
//
// A block stages a [128 x ld] tile of head rows in shared memory with coalesced 128-bit
// loads, each thread then owns one row (conflict-free, stride ld+1), and gradients go back
// through the same tile so global stores are coalesced too.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int LOSS_RB_MIN = 32;          // smallest rows-per-block of the loss kernel (workspace sizing)
constexpr int MAXC = MLB_MAX_ACTION_COMPONENTS;

struct Layout {
    int A;
    int off[MAXC];
    int nb[MAXC];
    float obj_scale[MAXC];
    float ent_scale[MAXC];
    // continuous action group (ContinuousActionDistributions, ml/dists.py:211-284): component i = action
    // dimension i, raw mean in column i, raw std in column A + i; std = (hi - lo) * sigmoid(raw + 2) + lo
    int continuous;
    float std_lo, std_hi;
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
// jax.scipy.stats.norm.logpdf(x, loc, scale) = -(z^2)/2 - log(scale) - log(sqrt(2 pi)), z = (x - loc) / scale
__device__ __forceinline__ float normal_logpdf(float x, float mu, float sd) {
    const float z = (x - mu) / sd;
    return -0.5f * z * z - logf(sd) - 0.918938533204672742f;
}

// Distributional critic (DreamerV3Critic, ml/models.py:157-174 -> SymExpTwoHotDistribution,
// ml/dists.py:119-208): V logits over fixed symexp-spaced bins.  V == 1 is the plain critic.
// kind 1: HL-Gauss critic (HLGaussCritic / HLGaussDist, ml/models.py:177-306): V logits over linearly spaced
// bin centres (`bins`), V + 1 bin bounds and the smoothness of the Gaussian histogram target.
struct CriticBins {
    int V;
    int kind;
    float smooth;
    float bins[MLB_MAX_CRITIC_BINS];
    float bounds[MLB_MAX_CRITIC_BINS + 1];
};

// SymExpTwoHotDistribution.mean (ml/dists.py:143-170): softmax-weighted bins, summed symmetrically
// around the midpoint so the estimate is exactly 0 at the zero-initialised critic.
__device__ __forceinline__ float twohot_mean(const float* l, const CriticBins& cb) {
    const int V = cb.V, mid = (V - 1) / 2;
    float mx = -INFINITY;
    for (int k = 0; k < V; ++k) mx = fmaxf(mx, l[k]);
    float se = 0.f;
    for (int k = 0; k < V; ++k) se += expf(l[k] - mx);
    const float inv = 1.f / se;
    float acc = 0.f;
    for (int k = 0; k < mid; ++k)
        acc += expf(l[mid - 1 - k] - mx) * inv * cb.bins[mid - 1 - k] + expf(l[mid + 1 + k] - mx) * inv * cb.bins[mid + 1 + k];
    return expf(l[mid] - mx) * inv * cb.bins[mid] + acc;
}

__global__ void rollout_keys_kernel(uint32_t* __restrict__ prng_key, uint32_t* __restrict__ policy_key,
                                    int part) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t k0 = prng_key[0], k1 = prng_key[1];
    uint32_t n0, n1, s0, s1, p0, p1;
    threefry_split_at(k0, k1, 0, 2, part, n0, n1);     // prng_key, step_key = split(prng_key)
    threefry_split_at(k0, k1, 1, 2, part, s0, s1);
    threefry_split_at(s0, s1, 0, 1, part, p0, p1);     // step_keys = split(step_key, 1)
    prng_key[0] = n0; prng_key[1] = n1;
    policy_key[0] = p0; policy_key[1] = p1;
}

// Staging of a [rb x ld] tile of head rows (ld % 4 == 0, 16-byte aligned rows): flat float4 /
// packed accesses over the whole row width; the shared tile keeps an odd row stride `ts` so the
// row-owning threads read it without bank conflicts.  Only the first `ncols` columns are kept.
__device__ __forceinline__ void stage_in(float* tile, int ts, const float* __restrict__ g, long long row0,
                                         long long rows, int ld, int ncols, int rb) {
    const int nrow = (int)min((long long)rb, rows - row0);
    const int vpr = ld >> 2;                                   // float4 per row
    const float4* src = reinterpret_cast<const float4*>(g + row0 * ld);
    const bool p2 = (vpr & (vpr - 1)) == 0;
    const int sh = 31 - __clz(vpr);
    for (int e = threadIdx.x; e < nrow * vpr; e += blockDim.x) {
        const int r = p2 ? (e >> sh) : e / vpr, c = (e - r * vpr) << 2;
        if (c < ncols) {
            const float4 v = __ldg(src + e);
            float* t = tile + r * ts + c;
            t[0] = v.x;
            if (c + 1 < ncols) t[1] = v.y;
            if (c + 2 < ncols) t[2] = v.z;
            if (c + 3 < ncols) t[3] = v.w;
        }
    }
}

// Gradients back to global: fp32 or bf16 rows of width ld; columns >= ncols are written as zeros.
__device__ __forceinline__ void stage_out(const float* tile, int ts, float* __restrict__ g, long long row0,
                                          long long rows, int ld, int ncols, int rb) {
    const int nrow = (int)min((long long)rb, rows - row0);
    const int vpr = ld >> 2;
    float4* dst = reinterpret_cast<float4*>(g + row0 * ld);
    const bool p2 = (vpr & (vpr - 1)) == 0;
    const int sh = 31 - __clz(vpr);
    for (int e = threadIdx.x; e < nrow * vpr; e += blockDim.x) {
        const int r = p2 ? (e >> sh) : e / vpr, c = (e - r * vpr) << 2;
        const float* t = tile + r * ts + c;
        dst[e] = make_float4(c < ncols ? t[0] : 0.f, c + 1 < ncols ? t[1] : 0.f, c + 2 < ncols ? t[2] : 0.f,
                             c + 3 < ncols ? t[3] : 0.f);
    }
}
__device__ __forceinline__ void stage_out(const float* tile, int ts, __nv_bfloat16* __restrict__ g, long long row0,
                                          long long rows, int ld, int ncols, int rb) {
    const int nrow = (int)min((long long)rb, rows - row0);
    const int vpr = ld >> 2;
    uint2* dst = reinterpret_cast<uint2*>(g + row0 * ld);
    const bool p2 = (vpr & (vpr - 1)) == 0;
    const int sh = 31 - __clz(vpr);
    for (int e = threadIdx.x; e < nrow * vpr; e += blockDim.x) {
        const int r = p2 ? (e >> sh) : e / vpr, c = (e - r * vpr) << 2;
        const float* t = tile + r * ts + c;
        const __nv_bfloat162 lo = __floats2bfloat162_rn(c < ncols ? t[0] : 0.f, c + 1 < ncols ? t[1] : 0.f);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(c + 2 < ncols ? t[2] : 0.f, c + 3 < ncols ? t[3] : 0.f);
        dst[e] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}

__device__ __forceinline__ float uniform_from_bits(uint32_t bits) {
    // jax.random.uniform(minval=tiny, maxval=1): 23 mantissa bits -> [1,2) - 1, clamped to tiny
    const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
    const float tiny = 1.17549435e-38f;
    return fmaxf(tiny, f * (1.0f - tiny) + tiny);
}

// One thread per (row, action component): the Gumbel-max draw of a component needs
// nb threefry evaluations, so spreading components over threads cuts the serial chain 6x and
// fills the SMs at rollout batch sizes (8192 rows -> 49152 threads).
__global__ void __launch_bounds__(128)
sample_kernel(const float* __restrict__ head, int ld, const uint32_t* __restrict__ policy_key,
              Layout L, long long rows, int part, int deterministic,
              int32_t* __restrict__ actions, float* __restrict__ log_probs,
              float* __restrict__ values, int vcol, CriticBins cb) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * L.A) return;
    const long long row = t / L.A;
    const int i = (int)(t - row * L.A);
    const float* l = head + row * ld;
    const int off = L.off[i], nb = L.nb[i];
    float mx = -INFINITY;
    for (int j = 0; j < nb; ++j) mx = fmaxf(mx, __ldg(l + off + j));
    float se = 0.f;
    for (int j = 0; j < nb; ++j) se += expf(__ldg(l + off + j) - mx);
    const float lse = logf(se) + mx;
    int best = 0;
    float bv = -INFINITY;
    if (deterministic) {
        for (int j = 0; j < nb; ++j) { const float v = __ldg(l + off + j); if (v > bv) { bv = v; best = j; } }
    } else {
        uint32_t c0, c1;
        threefry_split_at(policy_key[0], policy_key[1], (uint32_t)i, (uint32_t)L.A, part, c0, c1);   // sample_keys[i]
        const uint64_t size = (uint64_t)rows * (uint64_t)nb;
        for (int j = 0; j < nb; ++j) {
            const uint32_t bits = threefry_bits_at(c0, c1, (uint64_t)row * nb + j, size, part);
            const float g = -logf(-logf(uniform_from_bits(bits)));
            const float v = g + __ldg(l + off + j);
            if (v > bv) { bv = v; best = j; }
        }
    }
    actions[row * L.A + i] = best;
    if (log_probs) log_probs[row * L.A + i] = __ldg(l + off + best) - lse;
    if (values && i == 0) values[row] = cb.V == 1 ? __ldg(l + vcol) : twohot_mean(l + vcol, cb);
}

// XLA's fp32 erf_inv (the polynomial of M. Giles, "Approximating the erfinv function"), which
// jax.random.normal evaluates on a uniform in (-1, 1): normal = sqrt(2) * erf_inv(u)
__device__ __forceinline__ float erfinv_xla(float x) {
    float w = -log1pf(-x * x), p;
    if (w < 5.f) {
        w -= 2.5f;
        p = 2.81022636e-08f;
        p = fmaf(p, w, 3.43273939e-07f); p = fmaf(p, w, -3.5233877e-06f); p = fmaf(p, w, -4.39150654e-06f);
        p = fmaf(p, w, 0.00021858087f); p = fmaf(p, w, -0.00125372503f); p = fmaf(p, w, -0.00417768164f);
        p = fmaf(p, w, 0.246640727f); p = fmaf(p, w, 1.50140941f);
    } else {
        w = sqrtf(w) - 3.f;
        p = -0.000200214257f;
        p = fmaf(p, w, 0.000100950558f); p = fmaf(p, w, 0.00134934322f); p = fmaf(p, w, -0.00367342844f);
        p = fmaf(p, w, 0.00573950773f); p = fmaf(p, w, -0.0076224613f); p = fmaf(p, w, 0.00943887047f);
        p = fmaf(p, w, 1.00167406f); p = fmaf(p, w, 2.83297682f);
    }
    return fabsf(x) == 1.f ? copysignf(INFINITY, x) : p * x;
}

// ContinuousActionDistributions.sample / .best (ml/dists.py:216-258): one thread per (row, dimension).
// actions are stored as the fp32 bit pattern in the int32 action buffer.
__global__ void __launch_bounds__(128)
sample_continuous_kernel(const float* __restrict__ head, int ld, const uint32_t* __restrict__ policy_key,
                         int A, float lo, float hi, long long rows, int part, int deterministic,
                         int32_t* __restrict__ actions, float* __restrict__ log_probs,
                         float* __restrict__ values, int vcol, CriticBins cb) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * A) return;
    const long long row = t / A;
    const int i = (int)(t - row * A);
    const float* l = head + row * ld;
    const float mu = tanhf(__ldg(l + i));
    const float sd = (hi - lo) * sigmoid_f(__ldg(l + A + i) + 2.0f) + lo;
    float a = mu;
    if (!deterministic) {
        uint32_t c0, c1;
        threefry_split_at(policy_key[0], policy_key[1], 0u, 1u, part, c0, c1);      // sample_keys = split(prng_key, 1)
        const uint32_t bits = threefry_bits_at(c0, c1, (uint64_t)t, (uint64_t)rows * (uint64_t)A, part);
        // jax.random.normal: uniform(minval = nextafter(-1, 0), maxval = 1), then sqrt(2) * erf_inv
        const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
        const float mn = -0.99999994f;
        const float u = fmaxf(mn, f * (1.0f - mn) + mn);
        a = fmaf(1.41421356237f * erfinv_xla(u), sd, mu);
    }
    actions[row * A + i] = __float_as_int(a);
    if (log_probs) log_probs[row * A + i] = normal_logpdf(a, mu, sd);
    if (values && i == 0) values[row] = cb.V == 1 ? __ldg(l + vcol) : twohot_mean(l + vcol, cb);
}

struct LossPartial {
    double obj, vl, ent;           // weighted, scaled sums that make up the loss
    double s[4], ss[4];            // metric streams: action obj, value loss, |value err|, entropy
    float mn[4], mx[4];
};

// Fused PPO loss + head gradients.  Thread layout: a block owns RB rows; thread t plays role
// t / RB for row t % RB -- roles 0..A-1 are the action components, role A is the critic.  RB is a
// multiple of 32, so a warp holds 32 consecutive rows of ONE role: no divergence between lanes,
// conflict-free tile rows (odd stride), and (A+1)x the parallelism of a thread-per-row layout.
struct WarpPartial {
    float p0, p1;                  // loss contributions (obj | vl, entropy | -)
    float s0, ss0, mn0, mx0;       // metric stream 0 of the role (action obj | value loss)
    float s1, ss1, mn1, mx1;       // metric stream 1 of the role (entropy | |value err|)
};

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One action component of one row, bucket count known at compile time (warp-uniform dispatch):
// fully unrolled, logits / exponentials in registers, no predication.
struct CompOut { float obj, H; };
template <int NB>
__device__ __forceinline__ CompOut loss_component(float* __restrict__ l, int act, float olp, float a, float w,
                                                  float clip, float obj_scale, float ent_scale) {
    float lg[NB], e[NB];
    float mxl = -INFINITY;
#pragma unroll
    for (int j = 0; j < NB; ++j) { lg[j] = l[j]; mxl = fmaxf(mxl, lg[j]); }
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) { e[j] = expf(lg[j] - mxl); se += e[j]; }
    const float lse = logf(se) + mxl;
    const float inv_se = __frcp_rn(se);
    float H = 0.f, lp_new = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        lg[j] -= lse;                                   // log_softmax
        e[j] *= inv_se;                                 // softmax
        H = fmaf(-e[j], lg[j], H);
        lp_new = (j == act) ? lg[j] : lp_new;
    }
    const float ratio = expf(lp_new - olp);                         // ml/ppo.py:146-147
    const float surr1 = a * ratio;
    const float surr2 = a * fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
    const float obj = fminf(surr1, surr2);                          // :155-162
    const bool inside = (ratio >= 1.f - clip) && (ratio <= 1.f + clip);
    const float dobj = (surr1 <= surr2 || inside) ? a : 0.f;
    const float dlp = -(w * dobj * ratio) * obj_scale;              // d loss / d lp_new
    const float dH = -(w * ent_scale);                              // d loss / d H
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        float g = -e[j] * dlp + dH * (-e[j] * (lg[j] + H));
        if (j == act) g += dlp;
        l[j] = g;                                                   // overwrite logits with grads
    }
    return CompOut{obj, H};
}

// Final combine over the per-block partials, run by the LAST block of the loss kernel to finish
// (ticket counter in the workspace): one warp per reduced quantity -- 0-10 the fp64 sums (obj, vl,
// ent, s[4], ss[4]), 11-14 the minima, 15-18 the maxima -- then thread 0 assembles mlb_ppo_stats.
__device__ void loss_finalize(const LossPartial* __restrict__ part, int nparts, double rows, int A,
                              float vcoef, mlb_ppo_stats* __restrict__ out) {
    __shared__ double sd[11];
    __shared__ float sf[8];
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int q = threadIdx.x >> 5; q < 19; q += nw) {
        // independent loads first (the partials sit in L2: one round trip per 8 blocks, not per block)
        if (q < 11) {
            double a = 0.0;
            for (int b0 = lane; b0 < nparts; b0 += 32 * 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int b = b0 + 32 * u;
                    if (b < nparts) {
                        const LossPartial& P = part[b];
                        v[u] = q == 0 ? P.obj : q == 1 ? P.vl : q == 2 ? P.ent : q < 7 ? P.s[q - 3] : P.ss[q - 7];
                    } else v[u] = 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) a += v[u];
            }
            a = warp_sum(a);
            if (lane == 0) sd[q] = a;
        } else {
            const bool is_min = q < 15;
            const int k = is_min ? q - 11 : q - 15;
            float m = is_min ? INFINITY : -INFINITY;
            for (int b0 = lane; b0 < nparts; b0 += 32 * 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int b = b0 + 32 * u;
                    v[u] = b < nparts ? (is_min ? part[b].mn[k] : part[b].mx[k]) : (is_min ? INFINITY : -INFINITY);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) m = is_min ? fminf(m, v[u]) : fmaxf(m, v[u]);
            }
            m = is_min ? warp_min_f(m) : warp_max_f(m);
            if (lane == 0) sf[is_min ? k : 4 + k] = m;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double obj = sd[0], vl = sd[1], ent = sd[2];
        const float loss = (float)(-obj + (double)vcoef * vl - ent);
        out->loss = loss; out->action_obj = (float)obj; out->value_loss = (float)((double)vcoef * vl);
        out->entropy = (float)ent;
        mlb_metric* m = &out->metrics[0];          // 'Loss': a scalar record (ml/ppo.py:352)
        m->mean = loss; m->m2 = 0.f; m->min = loss; m->max = loss; m->count = 1;
    }
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;
        const double s = sd[3 + k], ss = sd[7 + k];
        const double cnt = (k == 0 || k == 3) ? rows * A : rows;
        const double mean = s / cnt;
        double m2 = ss - s * mean;
        if (m2 < 0) m2 = 0;
        mlb_metric* m = &out->metrics[k + 1];
        m->mean = (float)mean; m->m2 = (float)m2; m->min = sf[k]; m->max = sf[4 + k]; m->count = (int32_t)cnt;
    }
}

// MAXT / MINB: the common layouts (<= 7 roles x 64 rows = 448 threads) are compiled for 3 resident blocks per SM
// (48 registers): the kernel is a latency chain (tile load -> barrier -> per-row math -> barrier -> store) that
// only other resident blocks can hide (r2 ncu: barrier 7.9 + long_scoreboard 3.9 stall cycles per issue at 2
// blocks per SM).
template <int RB, int MAXT = 1024, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB)
ppo_loss_kernel(const float* __restrict__ head, int ld, const int32_t* __restrict__ actions,
                const float* __restrict__ old_lp, const float* __restrict__ adv,
                const float* __restrict__ ret, const float* __restrict__ old_v,
                const float* __restrict__ mb_w, const float* __restrict__ adv_mr,
                const float* __restrict__ vn, Layout L, long long rows, long long M,
                float clip, float vcoef, int flags, int vcol, void* __restrict__ dhead,
                float* __restrict__ dbias, LossPartial* __restrict__ part, CriticBins cb, float inv_rows,
                unsigned int* __restrict__ ticket, mlb_ppo_stats* __restrict__ stats) {
    extern __shared__ float tile[];
    __shared__ WarpPartial wpart[32];
    pdl_launch_dependents();
    pdl_wait();
    const int ncols = vcol + cb.V;
    const int ts = ncols | 1;                 // odd tile stride
    float* colsum = tile + RB * ts;           // [ncols]
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) colsum[c] = 0.f;
    const int role = threadIdx.x / RB, r = threadIdx.x - role * RB;
    // A block walks row tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...: the launch is one resident wave
    // (host: grid = min(tiles, resident blocks)), so the block-level tail -- the fp64 fold of the warp partials,
    // the bias-gradient flush, the ticket -- is paid once per block instead of once per 64 rows (a third of a
    // one-tile block's life was spent at those barriers).  Per-thread accumulators carry the loss / metric
    // partials across the tiles; the bias-gradient column sums accumulate in shared memory.
    float ap0 = 0.f, ap1 = 0.f, as0 = 0.f, ass0 = 0.f, as1 = 0.f, ass1 = 0.f;
    float amn0 = INFINITY, amx0 = -INFINITY, amn1 = INFINITY, amx1 = -INFINITY;
    const long long ntiles = (rows + RB - 1) / RB;
    bool first = true;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long row0 = t * RB;
    if (!first) __syncthreads();              // the previous tile's stage_out / column sums have read `tile`
    first = false;
    stage_in(tile, ts, head, row0, rows, ld, ncols, RB);
    __syncthreads();
    const long long row = row0 + r;
    const bool valid = row < rows;
    float p0 = 0.f, p1 = 0.f, x0 = 0.f, x1 = 0.f;
    if (valid) {
        float* l = tile + r * ts;
        const float w = mb_w ? mb_w[row % M] : 1.f;
        if (role < L.A && L.continuous) {
            // continuous action dimension `role`: Normal(tanh(m), std(s)) (ml/dists.py:260-284), the clipped
            // surrogate and the entropy bonus of ml/ppo.py:146-165 on its log-density
            const int i = role;
            float a = adv[row];
            if (adv_mr) a = (a - adv_mr[0]) * adv_mr[1];
            const float x = __int_as_float(actions[row * L.A + i]);
            const float olp = old_lp[row * L.A + i];
            const float os = L.obj_scale[i], es = L.ent_scale[i];
            const float mu = tanhf(l[i]);
            const float sg = sigmoid_f(l[L.A + i] + 2.0f);
            const float sd = (L.std_hi - L.std_lo) * sg + L.std_lo;
            const float lp = normal_logpdf(x, mu, sd);
            const float H = 0.5f * logf(6.28318530717958648f * sd * sd) + 0.5f;
            const float ratio = expf(lp - olp);
            const float surr1 = a * ratio;
            const float surr2 = a * fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
            const bool inside = (ratio >= 1.f - clip) && (ratio <= 1.f + clip);
            const float dobj = (surr1 <= surr2 || inside) ? a : 0.f;
            const float dlp = -(w * dobj * ratio) * os;
            const float dH = -(w * es);
            const float d = x - mu;
            const float dmu = dlp * d / (sd * sd);                           // d lp / d mu
            const float dsd = dlp * (d * d / (sd * sd * sd) - 1.f / sd) + dH / sd;
            l[i] = dmu * (1.f - mu * mu);                                    // tanh'
            l[L.A + i] = dsd * (L.std_hi - L.std_lo) * sg * (1.f - sg);      // sigmoid'
            p0 = (w * fminf(surr1, surr2)) * os;
            p1 = (w * H) * es;
            x0 = fminf(surr1, surr2);
            x1 = H;
        } else if (role < L.A) {
            const int i = role;
            float a = adv[row];
            if (adv_mr) a = (a - adv_mr[0]) * adv_mr[1];                // zscore_data, per minibatch
            const int off = L.off[i], nb = L.nb[i];
            const int act = min(max(actions[row * L.A + i], 0), nb - 1);   // never index outside the bucket
            const float olp = old_lp[row * L.A + i];
            const float os = L.obj_scale[i], es = L.ent_scale[i];
            float* lo = l + off;
            CompOut o;
            switch (nb) {                                               // warp-uniform
                case 1: o = loss_component<1>(lo, act, olp, a, w, clip, os, es); break;
                case 2: o = loss_component<2>(lo, act, olp, a, w, clip, os, es); break;
                case 3: o = loss_component<3>(lo, act, olp, a, w, clip, os, es); break;
                case 4: o = loss_component<4>(lo, act, olp, a, w, clip, os, es); break;
                case 5: o = loss_component<5>(lo, act, olp, a, w, clip, os, es); break;
                case 6: o = loss_component<6>(lo, act, olp, a, w, clip, os, es); break;
                case 7: o = loss_component<7>(lo, act, olp, a, w, clip, os, es); break;
                case 8: o = loss_component<8>(lo, act, olp, a, w, clip, os, es); break;
                default: {
                    float mxl = -INFINITY;
                    for (int j = 0; j < nb; ++j) mxl = fmaxf(mxl, lo[j]);
                    float se = 0.f;
                    for (int j = 0; j < nb; ++j) se += expf(lo[j] - mxl);
                    const float lse = logf(se) + mxl;
                    const float inv_se = __frcp_rn(se);
                    float H = 0.f;
                    for (int j = 0; j < nb; ++j) H -= (expf(lo[j] - mxl) * inv_se) * (lo[j] - lse);
                    const float ratio = expf((lo[act] - lse) - olp);
                    const float surr1 = a * ratio;
                    const float surr2 = a * fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
                    const bool inside = (ratio >= 1.f - clip) && (ratio <= 1.f + clip);
                    const float dobj = (surr1 <= surr2 || inside) ? a : 0.f;
                    const float dlp = -(w * dobj * ratio) * os;
                    const float dH = -(w * es);
                    for (int j = 0; j < nb; ++j) {
                        const float lpj = lo[j] - lse;
                        const float pj = expf(lo[j] - mxl) * inv_se;
                        float g = -pj * dlp + dH * (-pj * (lpj + H));
                        if (j == act) g += dlp;
                        lo[j] = g;
                    }
                    o = CompOut{fminf(surr1, surr2), H};
                }
            }
            p0 = (w * o.obj) * os;
            p1 = (w * o.H) * es;
            x0 = o.obj;
            x1 = o.H;
        } else if (cb.V > 1 && cb.kind == 1) {
            // HL-Gauss critic (ml/ppo.py:178-185, HLGaussDist.loss ml/models.py:212-250): cross-entropy against
            // the histogram of a Gaussian centred on the (clipped) return, sigma = smoothness * bin width
            const int V = cb.V;
            float* lc = l + vcol;
            const float rr = ret[row];
            const float vmean = twohot_mean(lc, cb);
            const float t = fminf(fmaxf(rr, cb.bins[0]), cb.bins[V - 1]);
            int nle = 0;
            for (int k = 0; k <= V; ++k) nle += (cb.bounds[k] <= t);
            const int lo = min(max(nle - 1, 0), V - 1), hi = min(max(nle, 1), V);
            const float den = 1.41421356237f * (cb.smooth * (cb.bounds[hi] - cb.bounds[lo]));
            const float cdf0 = erff((cb.bounds[0] - t) / den);
            const float zinv = 1.f / (erff((cb.bounds[V] - t) / den) - cdf0);
            float mxc = -INFINITY;
            for (int k = 0; k < V; ++k) mxc = fmaxf(mxc, lc[k]);
            float sec = 0.f;
            for (int k = 0; k < V; ++k) sec += expf(lc[k] - mxc);
            const float lsec = logf(sec) + mxc;
            const float gsc = vcoef * w * inv_rows;
            float vl = 0.f, prev = cdf0;
            for (int k = 0; k < V; ++k) {
                const float cur = erff((cb.bounds[k + 1] - t) / den);
                const float ck = zinv * (cur - prev);
                prev = cur;
                const float lp = lc[k] - lsec;
                vl -= ck * lp;
                lc[k] = gsc * (expf(lp) - ck);                          // d CE / d logit = softmax - target
            }
            p0 = (w * vl) * inv_rows;
            x0 = vl;
            x1 = fabsf(vmean - rr);
        } else if (cb.V > 1) {
            // distributional critic: two-hot cross-entropy (ml/ppo.py:169-177, ml/dists.py:172-208)
            const int V = cb.V;
            float* lc = l + vcol;
            const float rr = ret[row];
            const float vmean = twohot_mean(lc, cb);
            int nle = 0, ngt = 0;
            for (int k = 0; k < V; ++k) { nle += (cb.bins[k] <= rr); ngt += (cb.bins[k] > rr); }
            const int lo = min(max(nle - 1, 0), V - 1), hi = min(max(V - ngt, 0), V - 1);
            const bool same = lo == hi;
            const float dl = same ? 1.f : fabsf(cb.bins[lo] - rr);
            const float du = same ? 1.f : fabsf(cb.bins[hi] - rr);
            const float wl = dl / (dl + du), wu = du / (dl + du);     // (sic) the reference's weights
            float mxc = -INFINITY;
            for (int k = 0; k < V; ++k) mxc = fmaxf(mxc, lc[k]);
            float sec = 0.f;
            for (int k = 0; k < V; ++k) sec += expf(lc[k] - mxc);
            const float lsec = logf(sec) + mxc;
            const float vl = -(wl * (lc[lo] - lsec) + wu * (lc[hi] - lsec));
            p0 = (w * vl) * inv_rows;
            x0 = vl;
            x1 = fabsf(vmean - rr);
            const float gsc = vcoef * w * inv_rows;
            for (int k = 0; k < V; ++k) {
                float tk = 0.f;
                if (k == lo) tk += wl;
                if (k == hi) tk += wu;
                lc[k] = gsc * (expf(lc[k] - lsec) - tk);                // d CE / d logit = softmax - target
            }
        } else {
            // critic (plain, V = 1): ml/ppo.py:186-218
            const float v = l[vcol];
            const float rr = ret[row];
            float verr, rn;
            if (vn) { verr = (v * vn[1] + vn[0]) - rr; rn = (rr - vn[2]) * vn[3]; }
            else { verr = v - rr; rn = rr; }
            float vu = v, vmask = 1.f;
            if (flags & MLB_PPO_CLIP_VALUE_LOSS) {
                const float lo = old_v[row] - clip, hi = old_v[row] + clip;
                vu = fminf(fmaxf(v, lo), hi);
                vmask = (v >= lo && v <= hi) ? 1.f : 0.f;
            }
            const float d = vu - rn;
            float vl, dvl;
            if (flags & MLB_PPO_HUBER_VALUE_LOSS) {
                const float q = fminf(fabsf(d), 1.f);
                vl = 0.5f * q * q + (fabsf(d) - q);
                dvl = fminf(fmaxf(d, -1.f), 1.f);
            } else { vl = 0.5f * d * d; dvl = d; }
            p0 = (w * vl) * inv_rows;
            x0 = vl;
            x1 = fabsf(verr);
            l[vcol] = vcoef * w * dvl * vmask * inv_rows;
        }
    }
    ap0 += p0; ap1 += p1;
    as0 += x0; ass0 = fmaf(x0, x0, ass0);
    as1 += x1; ass1 = fmaf(x1, x1, ass1);
    if (valid) { amn0 = fminf(amn0, x0); amx0 = fmaxf(amx0, x0); amn1 = fminf(amn1, x1); amx1 = fmaxf(amx1, x1); }
    __syncthreads();                          // every role has written its gradients into the tile
    if (flags & MLB_PPO_DHEAD_BF16) stage_out(tile, ts, reinterpret_cast<__nv_bfloat16*>(dhead), row0, rows, ld, ncols, RB);
    else stage_out(tile, ts, reinterpret_cast<float*>(dhead), row0, rows, ld, ncols, RB);
    if (dbias) {                              // bias gradients of the heads: column sums of this tile
        const int nrow = (int)min((long long)RB, rows - row0);
        const int c = threadIdx.x & 63, seg = threadIdx.x >> 6, nseg = blockDim.x >> 6;
        for (int cc = c; cc < ncols; cc += 64) {
            float acc = 0.f;
            for (int rr = seg; rr < nrow; rr += nseg) acc += tile[rr * ts + cc];
            atomicAdd(&colsum[cc], acc);
        }
    }
    }   // tile loop
    // warp-level partials (fp32 over the rows a warp visited), combined across warps / blocks in fp64 below
    {
        WarpPartial P;
        P.p0 = warp_sum(ap0); P.p1 = warp_sum(ap1);
        P.s0 = warp_sum(as0); P.ss0 = warp_sum(ass0);
        P.s1 = warp_sum(as1); P.ss1 = warp_sum(ass1);
        P.mn0 = warp_min_f(amn0); P.mx0 = warp_max_f(amx0);
        P.mn1 = warp_min_f(amn1); P.mx1 = warp_max_f(amx1);
        if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = P;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // one warp folds the per-warp partials of the block: lane = warp index
        const int nw = blockDim.x >> 5, wpr = RB >> 5;       // warps per role
        const int wi = threadIdx.x;
        const bool on = wi < nw;
        const bool comp = on && (wi / wpr) < L.A, crit = on && !comp;
        const WarpPartial P = wpart[on ? wi : 0];
        LossPartial out;
        out.obj = warp_sum(comp ? (double)P.p0 : 0.0);
        out.ent = warp_sum(comp ? (double)P.p1 : 0.0);
        out.vl = warp_sum(crit ? (double)P.p0 : 0.0);
        out.s[0] = warp_sum(comp ? (double)P.s0 : 0.0);  out.ss[0] = warp_sum(comp ? (double)P.ss0 : 0.0);
        out.s[3] = warp_sum(comp ? (double)P.s1 : 0.0);  out.ss[3] = warp_sum(comp ? (double)P.ss1 : 0.0);
        out.s[1] = warp_sum(crit ? (double)P.s0 : 0.0);  out.ss[1] = warp_sum(crit ? (double)P.ss0 : 0.0);
        out.s[2] = warp_sum(crit ? (double)P.s1 : 0.0);  out.ss[2] = warp_sum(crit ? (double)P.ss1 : 0.0);
        out.mn[0] = warp_min_f(comp ? P.mn0 : INFINITY); out.mx[0] = warp_max_f(comp ? P.mx0 : -INFINITY);
        out.mn[3] = warp_min_f(comp ? P.mn1 : INFINITY); out.mx[3] = warp_max_f(comp ? P.mx1 : -INFINITY);
        out.mn[1] = warp_min_f(crit ? P.mn0 : INFINITY); out.mx[1] = warp_max_f(crit ? P.mx0 : -INFINITY);
        out.mn[2] = warp_min_f(crit ? P.mn1 : INFINITY); out.mx[2] = warp_max_f(crit ? P.mx1 : -INFINITY);
        if (threadIdx.x == 0) part[blockIdx.x] = out;
    }
    if (dbias) {
        __syncthreads();
        for (int c = threadIdx.x; c < ncols; c += blockDim.x) atomicAdd(dbias + c, colsum[c]);
    }
    // the last block to finish folds all per-block partials into mlb_ppo_stats (no second launch)
    __shared__ bool last_block;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last_block = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (last_block) *ticket = 0;              // re-armed for the next launch
    }
    __syncthreads();
    if (last_block) {
        __threadfence();
        loss_finalize(part, (int)gridDim.x, (double)rows, L.A, vcoef, stats);
    }
}

// hlgauss: bins_host holds  centres[V] | bounds[V + 1] | smoothness
int make_bins(CriticBins& cb, const float* bins_host, int num_bins, bool hlgauss = false) {
    cb.kind = 0;
    cb.smooth = 0.f;
    if (num_bins <= 1 || !bins_host) { cb.V = 1; return hlgauss ? MLB_EINVAL : MLB_OK; }
    if (num_bins > MLB_MAX_CRITIC_BINS || num_bins % 2 == 0) return MLB_EINVAL;
    cb.V = num_bins;
    for (int k = 0; k < num_bins; ++k) cb.bins[k] = bins_host[k];
    if (hlgauss) {
        cb.kind = 1;
        for (int k = 0; k <= num_bins; ++k) cb.bounds[k] = bins_host[num_bins + k];
        cb.smooth = bins_host[2 * num_bins + 1];
        if (!(cb.smooth > 0.f)) return MLB_EINVAL;
    }
    return MLB_OK;
}

int make_layout(Layout& L, const int32_t* buckets, int A, const float* obj_scale,
                const float* ent_scale, int ld, int extra_cols, bool continuous = false) {
    L.continuous = 0;
    L.std_lo = L.std_hi = 0.f;
    if (continuous) {
        // A action dimensions: raw means in columns [0, A), raw stds in [A, 2A); ent_scale_host carries the two
        // extra floats stddev_min, stddev_max behind its A entries
        if (A <= 0 || A > MAXC || !ent_scale || !obj_scale) return MLB_EINVAL;
        L.A = A;
        L.continuous = 1;
        L.std_lo = ent_scale[A];
        L.std_hi = ent_scale[A + 1];
        if (!(L.std_hi >= L.std_lo) || !(L.std_lo >= 0.f)) return MLB_EINVAL;
        for (int i = 0; i < A; ++i) {
            L.off[i] = i; L.nb[i] = 0;
            L.obj_scale[i] = obj_scale[i];
            L.ent_scale[i] = ent_scale[i];
        }
        if (2 * A + extra_cols > ld) return MLB_EINVAL;
        return 2 * A;
    }
    if (A <= 0 || A > MAXC || !buckets) return MLB_EINVAL;
    L.A = A;
    int off = 0;
    for (int i = 0; i < A; ++i) {
        if (buckets[i] <= 0) return MLB_EINVAL;
        L.off[i] = off; L.nb[i] = buckets[i];
        L.obj_scale[i] = obj_scale ? obj_scale[i] : 0.f;
        L.ent_scale[i] = ent_scale ? ent_scale[i] : 0.f;
        off += buckets[i];
    }
    if (off + extra_cols > ld) return MLB_EINVAL;
    return off;
}

}  // namespace

MLB_API int mlb_rollout_keys(void* stream, uint32_t* prng_key, uint32_t* policy_key,
                             int partitionable) {
    MLB_REQUIRE(prng_key && policy_key);
    rollout_keys_kernel<<<1, 32, 0, mlb_stream(stream)>>>(prng_key, policy_key, partitionable);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_sample_discrete_f32(void* stream, const float* head, int ld,
                                    const uint32_t* policy_key, const int32_t* buckets_host,
                                    int num_components, long long rows, int partitionable,
                                    int deterministic, int32_t* actions, float* log_probs,
                                    float* values, const float* critic_bins_host,
                                    int num_critic_bins) {
    MLB_REQUIRE(head && actions && rows >= 0 && ld > 0 && (deterministic || policy_key));
    if (rows == 0) return MLB_OK;
    Layout L;
    CriticBins cb;
    if (make_bins(cb, critic_bins_host, num_critic_bins)) return MLB_EINVAL;
    const int vcol = make_layout(L, buckets_host, num_components, nullptr, nullptr, ld, values ? cb.V : 0);
    if (vcol < 0) return vcol;
    sample_kernel<<<mlb_cdiv(rows * num_components, 128), 128, 0, mlb_stream(stream)>>>(
        head, ld, policy_key, L, rows, partitionable, deterministic, actions, log_probs, values, vcol, cb);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_sample_continuous_f32(void* stream, const float* head, int ld, const uint32_t* policy_key,
                                      int num_dims, float stddev_min, float stddev_max, long long rows,
                                      int partitionable, int deterministic, int32_t* actions, float* log_probs,
                                      float* values, const float* critic_bins_host, int num_critic_bins) {
    MLB_REQUIRE(head && actions && rows >= 0 && ld > 0 && num_dims > 0 && (deterministic || policy_key));
    MLB_REQUIRE(stddev_max >= stddev_min && stddev_min >= 0.f);
    if (rows == 0) return MLB_OK;
    CriticBins cb;
    if (make_bins(cb, critic_bins_host, num_critic_bins)) return MLB_EINVAL;
    const int vcol = 2 * num_dims;
    MLB_REQUIRE(vcol + (values ? cb.V : 0) <= ld);
    sample_continuous_kernel<<<mlb_cdiv(rows * num_dims, 128), 128, 0, mlb_stream(stream)>>>(
        head, ld, policy_key, num_dims, stddev_min, stddev_max, rows, partitionable, deterministic, actions,
        log_probs, values, vcol, cb);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API size_t mlb_ppo_loss_workspace(long long rows) {
    return (size_t)mlb_cdiv(rows, LOSS_RB_MIN) * sizeof(LossPartial) + 16;      // + the ticket counter
}

MLB_API int mlb_ppo_loss_f32(void* stream, const float* head, int ld, const int32_t* actions,
                             const float* old_log_probs, const float* advantages,
                             const float* returns, const float* old_values,
                             const float* mb_weights, const float* adv_mean_rstd,
                             const float* vn_params, const int32_t* buckets_host,
                             const float* obj_scale_host, const float* ent_scale_host,
                             int num_components, long long rows, long long M, float clip_coef,
                             float value_loss_coef, int flags, void* d_head, float* d_bias,
                             mlb_ppo_stats* stats, void* ws, size_t ws_bytes,
                             const float* critic_bins_host, int num_critic_bins) {
    MLB_REQUIRE(head && actions && old_log_probs && advantages && returns && d_head && stats);
    MLB_REQUIRE(rows > 0 && M > 0 && ld > 0 && ld % 4 == 0 && obj_scale_host && ent_scale_host);
    MLB_REQUIRE(mlb_aligned16(head) && (reinterpret_cast<uintptr_t>(d_head) & 15) == 0);
    MLB_REQUIRE(!(flags & MLB_PPO_CLIP_VALUE_LOSS) || old_values);
    Layout L;
    CriticBins cb;
    if (make_bins(cb, critic_bins_host, num_critic_bins, (flags & MLB_PPO_HLGAUSS_CRITIC) != 0)) return MLB_EINVAL;
    MLB_REQUIRE(cb.V == 1 || !(vn_params || (flags & (MLB_PPO_CLIP_VALUE_LOSS | MLB_PPO_HUBER_VALUE_LOSS))));
    const int vcol = make_layout(L, buckets_host, num_components, obj_scale_host, ent_scale_host, ld, cb.V,
                                 (flags & MLB_PPO_CONTINUOUS_ACTIONS) != 0);
    if (vcol < 0) return vcol;
    const int RB = (num_components + 1) * 64 <= 1024 ? 64 : LOSS_RB_MIN;
    unsigned g = mlb_cdiv(rows, RB);
    if (!ws || ws_bytes < mlb_ppo_loss_workspace(rows)) return MLB_EWS;
    const int ncols = vcol + cb.V;
    const size_t smem = ((size_t)RB * (ncols | 1) + ncols) * sizeof(float);
    const int nthreads = (num_components + 1) * RB;
    auto kern = RB == 64 ? (nthreads <= 448 ? ppo_loss_kernel<64, 448, 3> : ppo_loss_kernel<64>) : ppo_loss_kernel<LOSS_RB_MIN>;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    {   // one resident wave: a block walks several row tiles (MLB_LOSS_PERSIST=0: one tile per block)
        static const bool persist = [] { const char* v = getenv("MLB_LOSS_PERSIST"); return !(v && v[0] == '0'); }();
        int per = 0;
        if (persist && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, nthreads, smem) == cudaSuccess &&
            per > 0 && g > (unsigned)(per * MLB_NUM_SMS))
            g = (unsigned)(per * MLB_NUM_SMS);
    }
    cudaStream_t s = mlb_stream(stream);
    LossPartial* part = reinterpret_cast<LossPartial*>(ws);
    cudaError_t le = launch_pdl(kern, dim3(g), dim3((num_components + 1) * RB), smem, s,
        head, ld, actions, old_log_probs, advantages,
        returns, old_values, mb_weights, adv_mean_rstd, vn_params, L, rows, M, clip_coef,
        value_loss_coef, flags, vcol, d_head, d_bias, part, cb, 1.f / (float)rows,
        reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + mlb_ppo_loss_workspace(rows) - 16), stats);
    if (le != cudaSuccess) return (int)le;
    return MLB_OK;
}
