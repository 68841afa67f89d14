// Observation statistics (ObservationsEMANormalizer), the scalar EMA estimate, and the alternate
// minibatch selections of _ppo:
//
//   mlb_obs_moments_f32 / mlb_obs_stats_merge_f32   per-step batch moments of the raw observations and
//       the equal-weight Chan merge over the steps of an update (ml/rollouts.py:670-678 ->
//       EMANormalizer.update_input_stats ml/moving_avg.py:103-129); the EMA itself is mlb_ema_update_f32
//   mlb_ema_estimate_update_f32     EMAEstimate.update_estimates (ml/moving_avg.py:22-44)
//   mlb_filter_adv_keys / mlb_filter_adv_select      filter_advantages (ml/ppo.py:374-405): sort keys of
//       abs(advantage) in flattened-time order, running max, threshold count, valid_inds
//   mlb_partition_valid             stable partition of the -1 entries to the tail (ml/ppo.py:453-458)
//   mlb_flat_time_index             RolloutData.flatten_time (ml/rollouts.py:331-334) as an index map
//   mlb_traj_scores_f32 / mlb_softmax_weights_f32 / mlb_gumbel_topk_keys / mlb_take_sorted_indices
//       importance_sample_trajectories (ml/ppo.py:407-435): trajectory scores, softmax, 1/(J p) weights,
//       Gumbel top-k (jax.random.choice(replace=False, p=...))
// Sorting itself is mlb_sort_u64 (prng.cu: the bitonic network that also drives the permutations).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// observation moments: raw[d] = {sum_n x[n, d], sum_n x[n, d]^2}
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
obs_moments_kernel(const float* __restrict__ x, long long N, int D, double* __restrict__ raw) {
    // block = 32 feature columns x 32 row lanes; coalesced 128-byte row segments
    const int col = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0, ss = 0.0;
    if (col < D)
        for (long long r = threadIdx.y; r < N; r += blockDim.y) {
            const double v = (double)x[r * D + col];
            s += v; ss += v * v;
        }
    __shared__ double sm[2][32][33];
    sm[0][threadIdx.y][threadIdx.x] = s;
    sm[1][threadIdx.y][threadIdx.x] = ss;
    __syncthreads();
    if (threadIdx.y == 0 && col < D) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 32; ++i) { a += sm[0][i][threadIdx.x]; b += sm[1][i][threadIdx.x]; }
        raw[2 * col] = a;
        raw[2 * col + 1] = b;
    }
}

// raw [T][D][2] -> (mean, var)[D]: per step b_mean, b_var (population), then the reference's running
// equal-weight merge in float32, step by step (n_a = number of previous steps)
__global__ void obs_stats_merge_kernel(const double* __restrict__ raw, int T, double count, int D,
                                       float* __restrict__ mean_out, float* __restrict__ var_out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float a_mean = 0.f, a_var = 0.f;
    for (int t = 0; t < T; ++t) {
        const double s = raw[((long long)t * D + d) * 2], ss = raw[((long long)t * D + d) * 2 + 1];
        const double m = s / count;
        double v = ss / count - m * m;
        if (v < 0.0) v = 0.0;
        const float b_mean = (float)m, b_var = (float)v;
        const float delta = b_mean - a_mean;
        const float b_w = 1.f / (float)(t + 1);
        const float a_w = 1.f - b_w;
        a_var = a_w * a_var + b_w * b_var + (delta * delta) * a_w * b_w;
        a_mean = a_mean + delta * b_w;
    }
    mean_out[d] = a_mean;
    var_out[d] = a_var;
}

// state = {mu, mu_biased, N (int32 bits)}
__global__ void ema_estimate_kernel(float* __restrict__ st, const float* __restrict__ x, float decay) {
    const float one_minus_alpha = decay, alpha = 1.f - decay;
    const int newN = reinterpret_cast<int*>(st)[2] + 1;
    const float mu_b = one_minus_alpha * st[1] + alpha * x[0];
    const float corr = -1.f / expm1f((float)newN * logf(one_minus_alpha));
    st[0] = mu_b * corr;
    st[1] = mu_b;
    reinterpret_cast<int*>(st)[2] = newN;
}

// ---------------------------------------------------------------------------------------
// filter_advantages
// ---------------------------------------------------------------------------------------
// flattened-time element f = j*Tp + s, trajectory j = c*B + b  <->  store[c, s, b]
__device__ __forceinline__ long long flat_to_store(long long f, int Tp, long long B) {
    const long long j = f / Tp, s = f - j * Tp;
    const long long c = j / B, b = j - c * B;
    return (c * Tp + s) * B + b;
}

// comp[f] = (~bits(abs(adv)) << 32) | f : ascending sort = descending abs(adv), ties by ascending f
// (jnp.argsort(descending=True) is stable); pads sort last.  max_bits: running max of abs(adv) bits.
__global__ void __launch_bounds__(256)
filter_keys_kernel(const float* __restrict__ adv, long long n, long long npad, int Tp, long long B,
                   unsigned long long* __restrict__ comp, unsigned int* __restrict__ max_bits) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int bits = 0u;
    if (f < n) {
        bits = __float_as_uint(fabsf(adv[flat_to_store(f, Tp, B)]));
        comp[f] = ((unsigned long long)(~bits) << 32) | (unsigned long long)(unsigned int)f;
    } else if (f < npad) {
        comp[f] = ~0ull;
    }
    unsigned int m = bits;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(max_bits, m);      // non-negative floats order like their bits
}

// one block: counts the elements with abs(adv) >= 0.01 * est (a prefix of the sorted keys), derives
// num_minibatches (ml/ppo.py:399-402) and writes valid_inds
__global__ void __launch_bounds__(1024)
filter_select_kernel(const unsigned long long* __restrict__ comp, long long n, long long M,
                     const float* __restrict__ est_mu, int32_t* __restrict__ valid, int32_t* __restrict__ counts) {
    __shared__ long long sm_cnt[32];
    __shared__ long long total;
    const float thr = 0.01f * est_mu[0];
    long long c = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = __uint_as_float(~(unsigned int)(comp[i] >> 32));
        c += (a >= thr) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sm_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm_cnt[w];
        long long nmb = (t + (M - 1)) / M;
        const long long cap = n / M;
        if (nmb > cap) nmb = cap;
        total = nmb * M;
        counts[0] = (int32_t)nmb;
        counts[1] = (int32_t)t;
    }
    __syncthreads();
    const long long ndp = total;
    for (long long i = threadIdx.x; i < n; i += blockDim.x)
        valid[i] = i < ndp ? (int32_t)(unsigned int)comp[i] : -1;
}

// stable partition of one row: entries >= 0 first (order kept), -1 after (ml/ppo.py:453-458)
__global__ void __launch_bounds__(1024)
partition_valid_kernel(const int32_t* __restrict__ x, int32_t* __restrict__ out, long long J) {
    const int32_t* row = x + (long long)blockIdx.x * J;
    int32_t* orow = out + (long long)blockIdx.x * J;
    __shared__ long long offs[1025];
    const long long seg = (J + blockDim.x - 1) / blockDim.x;
    const long long lo = threadIdx.x * seg, hi = min(J, lo + seg);
    long long c = 0;
    for (long long i = lo; i < hi; ++i) c += row[i] >= 0;
    offs[threadIdx.x + 1] = c;
    if (threadIdx.x == 0) offs[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 1; i <= (int)blockDim.x; ++i) offs[i] += offs[i - 1];
    __syncthreads();
    long long w = offs[threadIdx.x];
    for (long long i = lo; i < hi; ++i)
        if (row[i] >= 0) orow[w++] = row[i];
    const long long total = offs[blockDim.x];
    for (long long i = total + threadIdx.x; i < J; i += blockDim.x) orow[i] = -1;
}

__global__ void __launch_bounds__(256)
flat_time_index_kernel(const int32_t* __restrict__ in, long long n, int Tp, long long B, int32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t f = in[i];
    out[i] = f < 0 ? -1 : (int32_t)flat_to_store(f, Tp, B);
}

// ---------------------------------------------------------------------------------------
// importance_sample_trajectories
// ---------------------------------------------------------------------------------------
// scores[j] = mean_s abs(adv) + mean_s abs(values - returns), trajectory j = c*B + b (P = 1)
__global__ void __launch_bounds__(256)
traj_scores_kernel(const float* __restrict__ adv, const float* __restrict__ val, const float* __restrict__ ret,
                   int Tp, long long B, long long J, float* __restrict__ scores) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    const long long c = j / B, b = j - c * B;
    float sa = 0.f, se = 0.f;
    for (int s = 0; s < Tp; ++s) {
        const long long o = (c * Tp + s) * B + b;
        sa += fabsf(adv[o]);
        se += fabsf(val[o] - ret[o]);
    }
    scores[j] = sa / (float)Tp + se / (float)Tp;
}

// probs = softmax(scores), weights = (1 / J) / probs   (one block; J <= a few million)
__global__ void __launch_bounds__(1024)
softmax_weights_kernel(const float* __restrict__ scores, long long J, float* __restrict__ probs,
                       float* __restrict__ weights) {
    __shared__ float smf[32];
    __shared__ double smd[32];
    __shared__ float bmax;
    __shared__ double bsum;
    float mx = -INFINITY;
    for (long long i = threadIdx.x; i < J; i += blockDim.x) mx = fmaxf(mx, scores[i]);
    mx = block_max_f(mx, smf);
    if (threadIdx.x == 0) bmax = mx;
    __syncthreads();
    mx = bmax;
    double s = 0.0;
    for (long long i = threadIdx.x; i < J; i += blockDim.x) s += (double)expf(scores[i] - mx);
    s = block_sum_d(s, smd);
    if (threadIdx.x == 0) bsum = s;
    __syncthreads();
    const float inv = (float)(1.0 / bsum), invJ = 1.f / (float)J;
    for (long long i = threadIdx.x; i < J; i += blockDim.x) {
        const float p = expf(scores[i] - mx) * inv;
        probs[i] = p;
        weights[i] = invJ / p;
    }
}

// g = gumbel(key, (J,)) + log(p); comp[j] = (~ordered(g) << 32) | j : ascending sort = descending g
__global__ void __launch_bounds__(256)
gumbel_keys_kernel(const uint32_t* __restrict__ key, const float* __restrict__ probs, long long J, long long Jpad,
                   int partitionable, unsigned long long* __restrict__ comp) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Jpad) return;
    if (j >= J) { comp[j] = ~0ull; return; }
    const uint32_t bits = threefry_bits_at(key[0], key[1], (uint64_t)j, (uint64_t)J, partitionable);
    // jax.random.uniform(minval=tiny, maxval=1): mantissa bits -> [1, 2) - 1, rescaled, clamped below
    const float tiny = 1.17549435e-38f;
    float u = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
    u = fmaxf(tiny, u * (1.0f - tiny) + tiny);
    const float g = -logf(-logf(u)) + logf(probs[j]);
    uint32_t o = __float_as_uint(g);
    o ^= (o >> 31) ? 0xFFFFFFFFu : 0x80000000u;                       // total order of floats as unsigned
    comp[j] = ((unsigned long long)(~o) << 32) | (unsigned long long)(unsigned int)j;
}

__global__ void __launch_bounds__(256)
take_sorted_kernel(const unsigned long long* __restrict__ comp, long long k, int32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) out[i] = (int32_t)(unsigned int)comp[i];
}

__global__ void __launch_bounds__(256)
gather_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = idx[i] >= 0 ? src[idx[i]] : 0.f;
}

}  // namespace

MLB_API int mlb_obs_moments_f32(void* stream, const float* obs, long long N, int D, double* raw) {
    MLB_REQUIRE(obs && raw && N > 0 && D > 0);
    obs_moments_kernel<<<mlb_cdiv(D, 32), dim3(32, 32), 0, mlb_stream(stream)>>>(obs, N, D, raw);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_obs_stats_merge_f32(void* stream, const double* raw, int T, double count, int D,
                                    float* mean_out, float* var_out) {
    MLB_REQUIRE(raw && mean_out && var_out && T > 0 && D > 0 && count > 0);
    obs_stats_merge_kernel<<<mlb_cdiv(D, 128), 128, 0, mlb_stream(stream)>>>(raw, T, count, D, mean_out, var_out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_ema_estimate_update_f32(void* stream, float* state, const float* x, float decay) {
    MLB_REQUIRE(state && x);
    ema_estimate_kernel<<<1, 1, 0, mlb_stream(stream)>>>(state, x, decay);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_filter_adv_keys(void* stream, const float* advantages, int C, int Tp, long long B,
                                long long n_pad, unsigned long long* comp, float* max_abs) {
    MLB_REQUIRE(advantages && comp && max_abs && C > 0 && Tp > 0 && B > 0);
    const long long n = (long long)C * Tp * B;
    MLB_REQUIRE(n_pad >= n && n < (1ll << 31));
    cudaStream_t s = mlb_stream(stream);
    cudaError_t e = cudaMemsetAsync(max_abs, 0, sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
    filter_keys_kernel<<<mlb_cdiv(n_pad, 256), 256, 0, s>>>(advantages, n, n_pad, Tp, B, comp,
                                                            reinterpret_cast<unsigned int*>(max_abs));
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_filter_adv_select(void* stream, const unsigned long long* comp_sorted, long long n, long long M,
                                  const float* max_adv_est_mu, int32_t* valid_inds, int32_t* counts) {
    MLB_REQUIRE(comp_sorted && max_adv_est_mu && valid_inds && counts && n > 0 && M > 0);
    filter_select_kernel<<<1, 1024, 0, mlb_stream(stream)>>>(comp_sorted, n, M, max_adv_est_mu, valid_inds, counts);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_partition_valid(void* stream, const int32_t* x, int32_t* out, int E, long long J) {
    MLB_REQUIRE(x && out && x != out && E > 0 && J > 0);
    partition_valid_kernel<<<E, 1024, 0, mlb_stream(stream)>>>(x, out, J);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_flat_time_index(void* stream, const int32_t* idx, long long n, int Tp, long long B, int32_t* out) {
    MLB_REQUIRE(idx && out && n >= 0 && Tp > 0 && B > 0);
    if (n == 0) return MLB_OK;
    flat_time_index_kernel<<<mlb_cdiv(n, 256), 256, 0, mlb_stream(stream)>>>(idx, n, Tp, B, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_traj_scores_f32(void* stream, const float* advantages, const float* values, const float* returns,
                                int C, int Tp, long long B, float* scores) {
    MLB_REQUIRE(advantages && values && returns && scores && C > 0 && Tp > 0 && B > 0);
    const long long J = (long long)C * B;
    traj_scores_kernel<<<mlb_cdiv(J, 256), 256, 0, mlb_stream(stream)>>>(advantages, values, returns, Tp, B, J, scores);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_softmax_weights_f32(void* stream, const float* scores, long long J, float* probs, float* weights) {
    MLB_REQUIRE(scores && probs && weights && J > 0);
    softmax_weights_kernel<<<1, 1024, 0, mlb_stream(stream)>>>(scores, J, probs, weights);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_gumbel_topk_keys(void* stream, const uint32_t* key, const float* probs, long long J, long long J_pad,
                                 int partitionable, unsigned long long* comp) {
    MLB_REQUIRE(key && probs && comp && J > 0 && J_pad >= J && J < (1ll << 31));
    gumbel_keys_kernel<<<mlb_cdiv(J_pad, 256), 256, 0, mlb_stream(stream)>>>(key, probs, J, J_pad, partitionable, comp);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_take_sorted_indices(void* stream, const unsigned long long* comp_sorted, long long k, int32_t* out) {
    MLB_REQUIRE(comp_sorted && out && k > 0);
    take_sorted_kernel<<<mlb_cdiv(k, 256), 256, 0, mlb_stream(stream)>>>(comp_sorted, k, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_gather_f32(void* stream, const float* src, const int32_t* idx, long long n, float* out) {
    MLB_REQUIRE(src && idx && out && n >= 0);
    if (n == 0) return MLB_OK;
    gather_f32_kernel<<<mlb_cdiv(n, 256), 256, 0, mlb_stream(stream)>>>(src, idx, n, out);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}
