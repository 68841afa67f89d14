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

constexpr int ROWS_PER_BLOCK = 128;
constexpr int MAXC = MLB_MAX_ACTION_COMPONENTS;

struct Layout {
    int A;
    int off[MAXC];
    int nb[MAXC];
    float obj_scale[MAXC];
    float ent_scale[MAXC];
};

// Distributional critic (DreamerV3Critic, ml/models.py:157-174 -> SymExpTwoHotDistribution,
// ml/dists.py:119-208): V logits over fixed symexp-spaced bins.  V == 1 is the plain critic.
struct CriticBins {
    int V;
    float bins[MLB_MAX_CRITIC_BINS];
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

// Division-free staging of the first `ncols` columns of a [128 x ld] tile of head rows: one warp
// per row, lanes over columns (shared tile stride `ts` is odd -> conflict-free row access).
__device__ __forceinline__ void stage_in(float* tile, int ts, const float* __restrict__ g, long long row0,
                                         long long rows, int ld, int ncols) {
    const int nrow = (int)min((long long)ROWS_PER_BLOCK, rows - row0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < nrow; r += nw) {
        const float* src = g + (row0 + r) * ld;
        for (int c = lane; c < ncols; c += 32) tile[r * ts + c] = __ldg(src + c);
    }
}

// Gradients back to global: fp32 or bf16 rows of width ld; columns >= ncols are written as zeros.
template <typename T>
__device__ __forceinline__ void stage_out(const float* tile, int ts, T* __restrict__ g, long long row0,
                                          long long rows, int ld, int ncols) {
    const int nrow = (int)min((long long)ROWS_PER_BLOCK, rows - row0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < nrow; r += nw) {
        T* dst = g + (row0 + r) * ld;
        for (int c = lane; c < ld; c += 32) dst[c] = (T)(c < ncols ? tile[r * ts + c] : 0.f);
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

struct LossPartial {
    double obj, vl, ent;           // weighted, scaled sums that make up the loss
    double s[4], ss[4];            // metric streams: action obj, value loss, |value err|, entropy
    float mn[4], mx[4];
};

__global__ void __launch_bounds__(ROWS_PER_BLOCK)
ppo_loss_kernel(const float* __restrict__ head, int ld, const int32_t* __restrict__ actions,
                const float* __restrict__ old_lp, const float* __restrict__ adv,
                const float* __restrict__ ret, const float* __restrict__ old_v,
                const float* __restrict__ mb_w, const float* __restrict__ adv_mr,
                const float* __restrict__ vn, Layout L, long long rows, long long M,
                float clip, float vcoef, int flags, int vcol, void* __restrict__ dhead,
                float* __restrict__ dbias, LossPartial* __restrict__ part, CriticBins cb) {
    extern __shared__ float tile[];
    const long long row0 = (long long)blockIdx.x * ROWS_PER_BLOCK;
    const int ncols = vcol + cb.V;
    const int ts = ncols | 1;                 // odd tile stride
    stage_in(tile, ts, head, row0, rows, ld, ncols);
    __syncthreads();
    const long long row = row0 + threadIdx.x;
    double p_obj = 0.0, p_vl = 0.0, p_ent = 0.0;
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    auto acc = [&](int k, float x) { s[k] += (double)x; ss[k] += (double)x * (double)x;
                                     mn[k] = fminf(mn[k], x); mx[k] = fmaxf(mx[k], x); };
    if (row < rows) {
        float* l = tile + threadIdx.x * ts;
        const float w = mb_w ? mb_w[row % M] : 1.f;
        float a = adv[row];
        if (adv_mr) a = (a - adv_mr[0]) * adv_mr[1];                    // zscore_data, per minibatch
        const float inv_rows = 1.f / (float)rows;
        for (int i = 0; i < L.A; ++i) {
            const int off = L.off[i], nb = L.nb[i];
            float mxl = -INFINITY;
            for (int j = 0; j < nb; ++j) mxl = fmaxf(mxl, l[off + j]);
            float se = 0.f;
            for (int j = 0; j < nb; ++j) se += expf(l[off + j] - mxl);
            const float lse = logf(se) + mxl;
            const float inv_se = 1.f / se;
            float H = 0.f;
            for (int j = 0; j < nb; ++j) {
                const float lp = l[off + j] - lse;
                const float p = expf(l[off + j] - mxl) * inv_se;        // jax.nn.softmax
                H -= p * lp;
            }
            const int act = min(max(actions[row * L.A + i], 0), nb - 1);   // never index outside the bucket
            const float lp_new = l[off + act] - lse;
            const float ratio = expf(lp_new - old_lp[row * L.A + i]);   // ml/ppo.py:146-147
            const float surr1 = a * ratio;
            const float cr = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
            const float surr2 = a * cr;
            const float obj = fminf(surr1, surr2);                      // :155-162
            const bool inside = (ratio >= 1.f - clip) && (ratio <= 1.f + clip);
            const float dobj = (surr1 <= surr2 || inside) ? a : 0.f;
            const float dlp = -(w * dobj * ratio) * L.obj_scale[i];     // d loss / d lp_new
            const float dH = -(w * L.ent_scale[i]);                     // d loss / d H
            p_obj += (double)(w * obj) * (double)L.obj_scale[i];
            p_ent += (double)(w * H) * (double)L.ent_scale[i];
            acc(0, obj);
            acc(3, H);
            for (int j = 0; j < nb; ++j) {
                const float lp = l[off + j] - lse;
                const float p = expf(l[off + j] - mxl) * inv_se;
                float g = -p * dlp + dH * (-p * (lp + H));
                if (j == act) g += dlp;
                l[off + j] = g;                                         // overwrite logits with grads
            }
        }
        if (cb.V > 1) {
            // distributional critic: two-hot cross-entropy (ml/ppo.py:169-177, ml/dists.py:172-208)
            const int V = cb.V;
            float* lc = l + vcol;
            const float r = ret[row];
            const float vmean = twohot_mean(lc, cb);
            int nle = 0, ngt = 0;
            for (int k = 0; k < V; ++k) { nle += (cb.bins[k] <= r); ngt += (cb.bins[k] > r); }
            const int lo = min(max(nle - 1, 0), V - 1), hi = min(max(V - ngt, 0), V - 1);
            const bool same = lo == hi;
            const float dl = same ? 1.f : fabsf(cb.bins[lo] - r);
            const float du = same ? 1.f : fabsf(cb.bins[hi] - r);
            const float wl = dl / (dl + du), wu = du / (dl + du);     // (sic) the reference's weights
            float mxc = -INFINITY;
            for (int k = 0; k < V; ++k) mxc = fmaxf(mxc, lc[k]);
            float sec = 0.f;
            for (int k = 0; k < V; ++k) sec += expf(lc[k] - mxc);
            const float lsec = logf(sec) + mxc;
            const float vl = -(wl * (lc[lo] - lsec) + wu * (lc[hi] - lsec));
            p_vl += (double)(w * vl) * (double)inv_rows;
            acc(1, vl);
            acc(2, fabsf(vmean - r));
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
        const float r = ret[row];
        float verr, rn;
        if (vn) { verr = (v * vn[1] + vn[0]) - r; rn = (r - vn[2]) * vn[3]; }
        else { verr = v - r; rn = r; }
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
        p_vl += (double)(w * vl) * (double)inv_rows;
        acc(1, vl);
        acc(2, fabsf(verr));
        l[vcol] = vcoef * w * dvl * vmask * inv_rows;
        }
    }
    __syncthreads();
    if (flags & MLB_PPO_DHEAD_BF16) stage_out(tile, ts, reinterpret_cast<__nv_bfloat16*>(dhead), row0, rows, ld, ncols);
    else stage_out(tile, ts, reinterpret_cast<float*>(dhead), row0, rows, ld, ncols);
    if (dbias) {                              // bias gradients of the heads: column sums of this tile
        const int nrow = (int)min((long long)ROWS_PER_BLOCK, rows - row0);
        for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
            float acc = 0.f;
            for (int r = 0; r < nrow; ++r) acc += tile[r * ts + c];
            atomicAdd(dbias + c, acc);
        }
    }

    __shared__ double smd[32];
    __shared__ float smf[32];
    LossPartial P;
    P.obj = block_sum_d(p_obj, smd);
    P.vl = block_sum_d(p_vl, smd);
    P.ent = block_sum_d(p_ent, smd);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        P.s[k] = block_sum_d(s[k], smd);
        P.ss[k] = block_sum_d(ss[k], smd);
        P.mn[k] = block_min_f(mn[k], smf);
        P.mx[k] = block_max_f(mx[k], smf);
    }
    if (threadIdx.x == 0) part[blockIdx.x] = P;
}

__global__ void __launch_bounds__(256)
ppo_loss_final_kernel(const LossPartial* __restrict__ part, int nparts, double rows, int A,
                      float vcoef, mlb_ppo_stats* __restrict__ out) {
    __shared__ double smd[32];
    __shared__ float smf[32];
    double obj = 0, vl = 0, ent = 0;
    for (int b = threadIdx.x; b < nparts; b += blockDim.x) { obj += part[b].obj; vl += part[b].vl; ent += part[b].ent; }
    obj = block_sum_d(obj, smd);
    vl = block_sum_d(vl, smd);
    ent = block_sum_d(ent, smd);
    if (threadIdx.x == 0) {
        const float loss = (float)(-obj + (double)vcoef * vl - ent);
        out->loss = loss; out->action_obj = (float)obj; out->value_loss = (float)((double)vcoef * vl);
        out->entropy = (float)ent;
        mlb_metric* m = &out->metrics[0];          // 'Loss': a scalar record (ml/ppo.py:352)
        m->mean = loss; m->m2 = 0.f; m->min = loss; m->max = loss; m->count = 1;
    }
    for (int k = 0; k < 4; ++k) {
        double s = 0, ss = 0;
        float mn = INFINITY, mx = -INFINITY;
        for (int b = threadIdx.x; b < nparts; b += blockDim.x) {
            s += part[b].s[k]; ss += part[b].ss[k];
            mn = fminf(mn, part[b].mn[k]); mx = fmaxf(mx, part[b].mx[k]);
        }
        s = block_sum_d(s, smd);
        ss = block_sum_d(ss, smd);
        mn = block_min_f(mn, smf);
        mx = block_max_f(mx, smf);
        if (threadIdx.x == 0) {
            const double cnt = (k == 0 || k == 3) ? rows * A : rows;
            const double mean = s / cnt;
            double m2 = ss - s * mean;
            if (m2 < 0) m2 = 0;
            mlb_metric* m = &out->metrics[k + 1];
            m->mean = (float)mean; m->m2 = (float)m2; m->min = mn; m->max = mx; m->count = (int32_t)cnt;
        }
    }
}

int make_bins(CriticBins& cb, const float* bins_host, int num_bins) {
    if (num_bins <= 1 || !bins_host) { cb.V = 1; return MLB_OK; }
    if (num_bins > MLB_MAX_CRITIC_BINS || num_bins % 2 == 0) return MLB_EINVAL;
    cb.V = num_bins;
    for (int k = 0; k < num_bins; ++k) cb.bins[k] = bins_host[k];
    return MLB_OK;
}

int make_layout(Layout& L, const int32_t* buckets, int A, const float* obj_scale,
                const float* ent_scale, int ld, int extra_cols) {
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

MLB_API size_t mlb_ppo_loss_workspace(long long rows) {
    return (size_t)mlb_cdiv(rows, ROWS_PER_BLOCK) * sizeof(LossPartial);
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
    MLB_REQUIRE(rows > 0 && M > 0 && ld > 0 && obj_scale_host && ent_scale_host);
    MLB_REQUIRE(!(flags & MLB_PPO_CLIP_VALUE_LOSS) || old_values);
    Layout L;
    CriticBins cb;
    if (make_bins(cb, critic_bins_host, num_critic_bins)) return MLB_EINVAL;
    MLB_REQUIRE(cb.V == 1 || !(vn_params || (flags & (MLB_PPO_CLIP_VALUE_LOSS | MLB_PPO_HUBER_VALUE_LOSS))));
    const int vcol = make_layout(L, buckets_host, num_components, obj_scale_host, ent_scale_host, ld, cb.V);
    if (vcol < 0) return vcol;
    const unsigned g = mlb_cdiv(rows, ROWS_PER_BLOCK);
    if (!ws || ws_bytes < (size_t)g * sizeof(LossPartial)) return MLB_EWS;
    const size_t smem = (size_t)ROWS_PER_BLOCK * (ld + 1) * sizeof(float);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(ppo_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaStream_t s = mlb_stream(stream);
    LossPartial* part = reinterpret_cast<LossPartial*>(ws);
    ppo_loss_kernel<<<g, ROWS_PER_BLOCK, smem, s>>>(head, ld, actions, old_log_probs, advantages,
        returns, old_values, mb_weights, adv_mean_rstd, vn_params, L, rows, M, clip_coef,
        value_loss_coef, flags, vcol, d_head, d_bias, part, cb);
    MLB_CHECK_LAUNCH();
    ppo_loss_final_kernel<<<1, 256, 0, s>>>(part, (int)g, (double)rows, num_components,
                                            value_loss_coef, stats);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
