/* libmlb200 -- C ABI of the B200-native madrona-learn learner hot path.
 *
 * This is the drop-in boundary (DESIGN.md "Boundary").  The reference has no native layer:
 * every function below replaces an XLA-compiled jnp expression inside the reference's Python
 * (cited per function as ml/<file>:<lines>, ml = /root/reference/src/madrona_learn).  The
 * signatures are what a jax.ffi custom-call (or the ctypes binding shipped in
 * madrona-learn_b200/_lib.py) binds:
 *
 *   int mlb_<op>(void* stream, <const T* in...>, <T* out...>, <dims...>, <scalars...>);
 *
 *   - `stream` is a cudaStream_t.  Every call only ENQUEUES work on it: no allocation, no
 *     synchronisation, no host<->device copy, safe to capture in a CUDA graph.
 *   - All pointers are DEVICE pointers owned by the caller; outputs are pre-allocated.
 *     Scratch is caller-provided (`ws`, sized by the matching mlb_*_workspace()).
 *   - Returns 0 on success, a positive cudaError_t if the launch failed, or a negative
 *     MLB_E* code for invalid arguments.  Never throws, never prints.
 *   - Re-entrant; no mutable global state (the MLB_* environment switches documented in
 *     README.md are read once, on first use, and are immutable afterwards); callable from
 *     any host thread.
 *   - Built for sm_100a only (nvcc -gencode arch=compute_100a,code=sm_100a).
 */
#ifndef MLB200_H_
#define MLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLB_OK       0
#define MLB_EINVAL  (-1)   /* bad dimension / null pointer / unsupported combination */
#define MLB_EALIGN  (-2)   /* pointer not aligned as required */
#define MLB_EWS     (-3)   /* workspace too small */

/* ABI version; bumped on any signature change. */
int mlb_abi_version(void);

/* ------------------------------------------------------------------------------------ */
/* Metric records (ml/metrics.py:12-98): {mean, m2, min, max} float32 + count int32.      */
/* ------------------------------------------------------------------------------------ */
typedef struct mlb_metric {
    float mean, m2, min, max;
    int32_t count;
} mlb_metric;

/* ------------------------------------------------------------------------------------ */
/* K1: GAE(lambda) + returns, reverse-time column-parallel scan over [T, N] buffers.      */
/* Replaces compute_advantages (ml/algo_common.py:84-130), `returns = advantages +       */
/* values` (ml/rollouts.py:769), the value-normaliser invert on load (ml/rollouts.py:     */
/* 726-741) and, when `metrics` != NULL, the full-buffer Metric.init_from_data reductions  */
/* for Rewards / Values / Est Returns / Advantages (ml/rollouts.py:806-816).              */
/*   rewards, values : f32 [T, N]     dones : u8 [T, N] (0/1)     bootstrap : f32 [N]     */
/*   advantages, returns : f32 [T, N] out (returns may be NULL)                           */
/*   gamma_lambda = (float)(gamma * gae_lambda) formed in double (ml/algo_common.py:120)   */
/*   vn_mu_sigma : NULL, or the device EMA-normaliser state of dim 1 (layout below: mu at  */
/*                 [0], sigma at [2]): values/bootstrap are stored normalised and are      */
/*                 inverted (v*sigma+mu) on load; the un-normalised values are what        */
/*                 `returns` and the Values metric see.                                    */
/*   metrics : NULL, or device mlb_metric[4] out (rewards, values, returns, advantages);   */
/*             needs ws of mlb_gae_workspace(T, N) bytes.                                  */
/* Algorithmic bytes: 17*T*N + 4*N.                                                       */
/* ------------------------------------------------------------------------------------ */
size_t mlb_gae_workspace(int T, long long N);
int mlb_gae_f32(void* stream, const float* rewards, const float* values, const uint8_t* dones,
                const float* bootstrap, float* advantages, float* returns,
                int T, long long N, float gamma, float gamma_lambda,
                const float* vn_mu_sigma, mlb_metric* metrics, void* ws, size_t ws_bytes);

/* Discounted returns only: compute_returns (ml/algo_common.py:45-81). 9*T*N + 4*N bytes. */
int mlb_returns_f32(void* stream, const float* rewards, const uint8_t* dones,
                    const float* bootstrap, float* returns, int T, long long N, float gamma);

/* ------------------------------------------------------------------------------------ */
/* K2: moments / z-score (ml/algo_common.py:133-140).                                     */
/* mlb_moments_f32 -> out f32[4] = {mean, rstd = rsqrt(max(var, var_floor)), var, n}.     */
/* mlb_zscore_apply_f32: out = (x - mean) * rstd, one pass (8 B/elem).                    */
/* mlb_zscore_f32 = moments + apply (x is read twice).                                    */
/* ------------------------------------------------------------------------------------ */
size_t mlb_moments_workspace(long long n);
int mlb_moments_f32(void* stream, const float* x, long long n, float var_floor,
                    float* out4, void* ws, size_t ws_bytes);
int mlb_zscore_apply_f32(void* stream, const float* x, float* out, long long n,
                         const float* mean_rstd);
int mlb_zscore_f32(void* stream, const float* x, float* out, long long n,
                   float* out4, void* ws, size_t ws_bytes);

/* Full-buffer Metric.init_from_data (ml/metrics.py:31-48) of a contiguous f32 array. */
int mlb_metric_f32(void* stream, const float* x, long long n, mlb_metric* out,
                   void* ws, size_t ws_bytes);

/* Per-trajectory first/second moments of a [T, N] buffer split in C BPTT chunks:        */
/* out f64 [C*N][2] = {sum, sum of squares} of x[c*T'+s, n] over s, trajectory j=c*N+n    */
/* (ml/rollouts.py:788-804 layout, P=1).                                                  */
int mlb_traj_moments_f32(void* stream, const float* x, int T, long long N, int C, double* out);

/* Per-minibatch moments for ALL (epoch, minibatch) pairs of an update at once (SURVEY    */
/* App. C.1): perm i32 [E, J] trajectory ids; minibatch k of epoch e = perm[e, k*M:(k+1)*M]*/
/* out f32 [E*J/M][4] = {mean, rstd(var floor), var, n} with n = M*T'.                    */
/* raw_out (may be NULL): f64 [E*J/M][2] = {sum, sumsq} -- what a data-parallel run        */
/* all-reduces (SUM) across ranks before mlb_moments_finalize_f32 (count = global n).       */
int mlb_mb_moments_f32(void* stream, const double* traj_moments, const int32_t* perm,
                       int E, long long J, long long M, int Tp, float var_floor, float* out,
                       double* raw_out);
int mlb_moments_finalize_f32(void* stream, const double* raw, int K, double count,
                             float var_floor, float* out);

/* ------------------------------------------------------------------------------------ */
/* K3: EMA normaliser (ml/moving_avg.py:48-198).  State: f32 [5][dim] rows = mu,           */
/* inv_sigma, sigma, mu_biased, sigma_sq_biased, followed by one int32 N (stored as the   */
/* first word after the 5*dim floats).                                                     */
/* ------------------------------------------------------------------------------------ */
int mlb_ema_update_f32(void* stream, float* state, int dim, const float* batch_mean,
                       const float* batch_var, float decay, float eps);
/* Value-normaliser recurrence over all minibatches of an update (ml/ppo.py:205-211 inside */
/* the loop of :460-482): consumes mb moments [K][4] (mean at 0, var at 2) and emits       */
/* per-minibatch f32 [K][4] = {mu_old, sigma_old, mu_new, inv_sigma_new}; state advanced K  */
/* times.  dim must be 1.                                                                  */
int mlb_ema_scan_f32(void* stream, float* state, const float* mb_moments, int K,
                     float decay, float eps, float* out);
int mlb_ema_normalize_f32(void* stream, const float* state, int dim, const float* x,
                          float* out, long long rows);
int mlb_ema_invert_f32(void* stream, const float* state, int dim, const float* x,
                       float* out, long long rows);

/* ObservationsEMANormalizer statistics (ml/observations.py:71-132, ml/rollouts.py:670-678):      */
/* mlb_obs_moments_f32: raw f64 [D][2] = per-feature {sum, sum of squares} over the N rows of one  */
/* step's raw observations (what a data-parallel run SUM-all-reduces);                            */
/* mlb_obs_stats_merge_f32: raw [T][D][2] of the T steps of an update -> (mean, var) f32 [D] by     */
/* the reference's running equal-weight Chan merge (EMANormalizer.update_input_stats,             */
/* ml/moving_avg.py:103-129; count = rows per step); feed the result to mlb_ema_update_f32.        */
int mlb_obs_moments_f32(void* stream, const float* obs, long long N, int D, double* raw);
int mlb_obs_stats_merge_f32(void* stream, const double* raw, int T, double count, int D,
                            float* mean_out, float* var_out);
/* EMAEstimate.update_estimates (ml/moving_avg.py:22-44).  state: f32 {mu, mu_biased} followed by   */
/* the int32 counter N; x: device f32 scalar (the value whose EMA is tracked).                    */
int mlb_ema_estimate_update_f32(void* stream, float* state, const float* x, float decay);

/* Per-step episodic-return bookkeeping of the rollout loop (ml/rollouts.py:938-939,971-973): */
/* er = r + gamma*er; trace[n] = er (may be NULL; feeds the 'Env Returns' metric);           */
/* er = done ? 0 : er.                                                                       */
int mlb_env_returns_f32(void* stream, const float* rewards, const uint8_t* dones,
                        float* env_returns, float* trace, long long N, float gamma);
/* "Post Step Rollout Store" (ml/rollouts.py:946-978) in one launch: copies the simulator's   */
/* rewards (f32 [N]) and dones (1 byte [N]) into the store slabs and advances the discounted   */
/* env-return trace exactly like mlb_env_returns_f32.                                          */
int mlb_post_step_store_f32(void* stream, const float* rewards, const uint8_t* dones,
                            float* reward_slab, uint8_t* done_slab, float* env_returns,
                            float* trace, long long N, float gamma);

/* ------------------------------------------------------------------------------------ */
/* PRNG: JAX threefry2x32 (bit-exact).  keys are uint32[2].  partitionable selects the     */
/* jax>=0.5 counter layout.                                                                */
/* ------------------------------------------------------------------------------------ */
int mlb_threefry_split(void* stream, const uint32_t* key, uint32_t* out, int num,
                       int partitionable);
int mlb_threefry_bits(void* stream, const uint32_t* key, uint32_t* out, long long n,
                      int partitionable);
/* The PPO minibatch permutations of one update (ml/ppo.py:445-458, ml/train_state.py:     */
/* 134-136): for e in [0,E): (rnd, key) = split(key); perm[e] = permutation(rnd, arange(J)) */
/* `key` (device uint32[2]) is advanced in place.  ws from mlb_ppo_permutations_workspace.  */
size_t mlb_ppo_permutations_workspace(int E, long long J);
int mlb_ppo_permutations(void* stream, uint32_t* key, int32_t* perm, int E, long long J,
                         int partitionable, void* ws, size_t ws_bytes);
/* The same shuffle applied to an arbitrary int32 array `values` [J] (device; shared by all E     */
/* epochs): perm[e] = random.permutation(rnd_e, values) -- the filter_advantages /               */
/* importance-sampling branches of _ppo permute `valid_inds` (ml/ppo.py:445-451).                */
int mlb_ppo_permutations_of(void* stream, uint32_t* key, const int32_t* values, int32_t* perm,
                            int E, long long J, int partitionable, void* ws, size_t ws_bytes);
/* Ascending in-place sort of n_pad (a power of two, mlb_sort_pad(n)) u64 keys: the bitonic       */
/* network behind the permutations, exposed for the argsort / top-k selections below (a key is    */
/* (order-preserving 32-bit value << 32) | index, pads are ~0).                                   */
long long mlb_sort_pad(long long n);
int mlb_sort_u64(void* stream, unsigned long long* keys, long long n_pad);

/* Index-exact data-parallel mode: this rank's M trajectory ids of every GLOBAL minibatch                 */
/* perm[e, k*world*M : (k+1)*world*M] (global ids j = c*(world*B) + r*B + b, perm identical on all ranks): */
/* every rank keeps the ids it owns (permutation order, at most M); the surplus of over-represented ranks */
/* fills the deficits of the others, both in rank order -- the union over ranks is exactly the global     */
/* minibatch, and only the binomial imbalance is fetched from peers.  out [E, nmb, M].                    */
int mlb_dp_assign_minibatches(void* stream, const int32_t* perm, long long perm_ld, int E, int nmb,
                              int world, int rank, long long B, long long M, int32_t* out);
/* ------------------------------------------------------------------------------------ */
/* K5: minibatch gather straight from the [C, T', P=1, B, *leaf] store                      */
/* (RolloutData.minibatch ml/rollouts.py:319-329 composed with the relayout :788-804):      */
/* out[s, m, :] = store[j/B, s, j%B, :], j = idx[m];  row_bytes = bytes of one leaf row.    */
/* ------------------------------------------------------------------------------------ */
int mlb_mb_gather(void* stream, const void* store, const int32_t* idx, void* out,
                  int C, int Tp, long long B, long long M, long long row_bytes);
/* rnn_start_states [C, B, row] -> [M, row] (ml/rollouts.py:800-804 + :321-323) */
int mlb_mb_gather_rnn(void* stream, const void* store, const int32_t* idx, void* out,
                      int C, long long B, long long M, long long row_bytes);
/* ------------------------------------------------------------------------------------ */
/* Alternate minibatch selections of _ppo (ml/ppo.py:374-435).  advantages / values / returns are  */
/* the [C, T', B] rollout-store leaves (P = 1); flattened-time element f = j*T' + s of trajectory  */
/* j = c*B + b (RolloutData.flatten_time, ml/rollouts.py:331-334).                                 */
/* filter_advantages: mlb_filter_adv_keys builds the sort keys of abs(advantage) (descending,      */
/*   stable) for all n = C*T'*B elements (+ pads up to n_pad) and max_abs = max abs(advantage);    */
/*   after mlb_sort_u64 and the EMAEstimate update, mlb_filter_adv_select counts the elements with */
/*   abs(adv) >= 0.01 * est_mu, sets counts = {num_minibatches, num_above_threshold} (:392-402)    */
/*   and valid_inds[i] = sorted index i < num_minibatches*M else -1 (:403-405).                    */
/* mlb_partition_valid: per row of x [E][J], entries >= 0 first in order, then -1 (:453-458).       */
/* mlb_flat_time_index: flattened-time ids -> row ids of the [C*T'*B] store (-1 stays -1), so that  */
/*   mlb_mb_gather_multi(C=1, Tp=1, B=C*T'*B) gathers single-step "trajectories".                  */
/* importance_sample_trajectories: scores[j] = mean_s abs(adv) + mean_s abs(values - returns)      */
/*   (:419-425); probs = softmax(scores), weights = (1/J)/probs (:426-427); mlb_gumbel_topk_keys =  */
/*   the keys of jax.random.choice(key, J, replace=False, p=probs) (Gumbel top-k: g = gumbel(key,   */
/*   (J,)) + log p, descending); after mlb_sort_u64, mlb_take_sorted_indices reads the first k ids. */
/* mlb_gather_f32: out[i] = src[idx[i]] (0 where idx < 0): traj_weights[mb_inds] (:468).            */
/* ------------------------------------------------------------------------------------ */
int mlb_filter_adv_keys(void* stream, const float* advantages, int C, int Tp, long long B,
                        long long n_pad, unsigned long long* keys, float* max_abs);
int mlb_filter_adv_select(void* stream, const unsigned long long* keys_sorted, long long n,
                          long long M, const float* max_adv_est_mu, int32_t* valid_inds,
                          int32_t* counts);
int mlb_partition_valid(void* stream, const int32_t* x, int32_t* out, int E, long long J);
int mlb_flat_time_index(void* stream, const int32_t* idx, long long n, int Tp, long long B,
                        int32_t* out);
int mlb_traj_scores_f32(void* stream, const float* advantages, const float* values,
                        const float* returns, int C, int Tp, long long B, float* scores);
int mlb_softmax_weights_f32(void* stream, const float* scores, long long J, float* probs,
                            float* weights);
int mlb_gumbel_topk_keys(void* stream, const uint32_t* key, const float* probs, long long J,
                         long long J_pad, int partitionable, unsigned long long* keys);
int mlb_take_sorted_indices(void* stream, const unsigned long long* keys_sorted, long long k,
                            int32_t* out);
int mlb_gather_f32(void* stream, const float* src, const int32_t* idx, long long n, float* out);

/* All leaves of one minibatch in a single launch.  Per leaf: store [C, T', B, row], out            */
/* [T', M, row] (may be NULL when out_bf16 is set), out_bf16 (may be NULL): the same rows converted */
/* f32 -> bf16 (row_bytes % 16 == 0) -- the tensor-core forward's A operand.  Up to 8 leaves.       */
typedef struct mlb_gather_leaf {
    const void* store;
    void* out;
    void* out_bf16;
    long long row_bytes;
} mlb_gather_leaf;
int mlb_mb_gather_multi(void* stream, const mlb_gather_leaf* leaves_host, int num_leaves,
                        const int32_t* idx, int C, int Tp, long long B, long long M);
/* Index-exact data-parallel variant (the world-sharded analogue of ml/ppo.py:437-466):  */
/* idx holds GLOBAL trajectory ids j = c*(world*B) + r*B + b; rank r's stores are reachable through */
/* peer_stores_host[r*num_leaves + leaf] (HOST array of device pointers mapped over NVLink, e.g.     */
/* symmetric memory; rank order; includes this rank's own stores).  leaves_host[i].store is ignored. */
/* Up to 8 ranks.  out[s, m, :] = store_owner[c, s, b, :]: bit-identical to mlb_mb_gather_multi on    */
/* the concatenated [C, T', world*B, row] store.                                                     */
int mlb_mb_gather_multi_peer(void* stream, const mlb_gather_leaf* leaves_host, int num_leaves,
                             const void* const* peer_stores_host, int world, const int32_t* idx,
                             int C, int Tp, long long B, long long M);

/* ------------------------------------------------------------------------------------ */
/* K6/K9 (fp32 path): Dense layers and their transposes.                                  */
/* C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N]); op(X) = X or X^T (row-major storage).   */
/* Replaces nn.Dense's dot_general (ml/models.py:110-115,129-135,148-154) and its autodiff  */
/* (ml/ppo.py:276-281).  splitk > 1 reduces K-slices with fp32 atomics into a pre-          */
/* initialised C (requires accumulate != 0).                                               */
/* ------------------------------------------------------------------------------------ */
int mlb_gemm_f32(void* stream, const float* A, const float* B, float* C, const float* bias,
                 int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                 int accumulate, int splitk);
/* The same product on the tensor cores (tcgen05.mma.kind::tf32, fp32 accumulation): what XLA:GPU  */
/* evaluates for an f32 dot_general at its default precision, i.e. the reference's compute_dtype   */
/* default (ml/cfg.py:96).  Same arguments as mlb_gemm_f32; operands are read as fp32 straight from */
/* the activations / master weights (TMA needs lda, ldb, ldc, N multiples of 4 and 16-byte aligned  */
/* bases: mlb_gemm_tf32_ok says whether a problem qualifies).  With accumulate != 0 the reduction is */
/* fp32 red.global.add and the launch widens splitk so that tiles x K-slices fill the SMs.           */
int mlb_gemm_tf32_tc(void* stream, const float* A, const float* B, float* C, const float* bias,
                     int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                     int accumulate, int splitk);
int mlb_gemm_tf32_ok(int M, int N, int K, int lda, int ldb, int ldc, const void* A, const void* B,
                     const void* C);
/* One Dense -> LayerNorm -> ReLU layer of compute_dtype=float32 (ml/models.py:107-117) in ONE launch: */
/* z = x W on tcgen05 kind::tf32 with the LayerNorm statistics, affine and ReLU computed out of TMEM in */
/* the epilogue (same arithmetic as mlb_ln_relu_fwd_f32).  x [rows, K] (row stride ldx), w [K, H] (the   */
/* flax Dense kernel), z [rows, H] or NULL, y [rows, H], stats [rows][2] = {mean, rstd} or NULL;         */
/* H in {64, 128, 256} (the accumulator tile spans the whole feature row).                               */
int mlb_dense_ln_relu_fwd_tf32(void* stream, const float* x, const float* w, const float* scale,
                               const float* bias, float* z, float* y, float* stats, long long rows, int K,
                               int H, int ldx);

/* LayerNorm(eps 1e-6, fast variance) + ReLU (ml/models.py:46-56,116-117).                  */
/* fwd: y = relu(((z-mean)*rstd)*scale+bias); stats (may be NULL) f32 [rows][2]={mean,rstd}  */
/* bwd: dz from dy; dscale/dbias are ACCUMULATED (fp32 atomics) into pre-zeroed buffers.     */
int mlb_ln_relu_fwd_f32(void* stream, const float* z, const float* scale, const float* bias,
                        float* y, float* stats, long long rows, int H);
int mlb_ln_relu_bwd_f32(void* stream, const float* dy, const float* z, const float* stats,
                        const float* scale, const float* bias, float* dz, float* dscale,
                        float* dbias, long long rows, int H);

int mlb_ln_relu_fwd_bf16(void* stream, const float* z, const float* scale, const float* bias,
                         void* y_bf16, float* stats, long long rows, int H);
int mlb_ln_relu_bwd_bf16(void* stream, const void* dy_bf16, const float* z, const float* stats,
                         const float* scale, const float* bias, void* dz_bf16, float* dscale,
                         float* dbias, long long rows, int H);

/* ------------------------------------------------------------------------------------ */
/* K6/K9 (tensor-core path, compute_dtype=bfloat16): tcgen05 / TMEM / TMA GEMM.           */
/* C[M,N] (+)= A * B^T, bf16 operands, fp32 accumulation.                                 */
/*   a_mn == 0: A row-major [M, K] (lda);  a_mn == 1: A stored [K, M] (MN-major operand)   */
/*   b_mn == 0: B row-major [N, K] (ldb);  b_mn == 1: B stored [K, N]                      */
/*   epi 0: fp32 store (+ bias[N]);  1: bf16 store;  2: fp32 atomic add into a             */
/*   pre-initialised C (split-K over `splitk` CTAs; the dW = X^T dZ product).              */
/* lda/ldb multiples of 8, N multiple of 8, 16-byte aligned pointers.                      */
/* ------------------------------------------------------------------------------------ */
int mlb_gemm_bf16_tc(void* stream, const void* A, const void* B, void* C, const float* bias,
                     int M, int N, int K, int lda, int ldb, int ldc, int a_mn, int b_mn, int epi,
                     int splitk);
/* Fused layer kernels (whole 128 x H accumulator tile in TMEM, H <= 256 multiple of 32, or 512): */
/* forward  Y = relu(LN(X Wt^T)*scale+bias): X bf16 [M,K] (ldx), Wt bf16 [HN,K] (ldw) = W^T;      */
/*   Y bf16 [M,HN]; XH (NULL at inference) bf16 [M,HN] = normalised pre-activation; rstd f32 [M]. */
/* backward dZ_out = LN'/ReLU'(DZ_in W^T) for the PREVIOUS layer, whose scale/bias/XH/rstd are    */
/*   given; W bf16 [HN, K] (ldw) is this layer's kernel [in=HN, out=K]; dscale/dbias f32 [HN]     */
/*   accumulated (pre-zeroed).  Z and dY never touch HBM.                                         */
int mlb_dense_ln_relu_fwd_tc(void* stream, const void* X, const void* Wt, const float* scale,
                             const float* bias, void* Y, void* XH, float* rstd, int M, int K,
                             int HN, int ldx, int ldw);
int mlb_dense_dx_lnbwd_tc(void* stream, const void* DZ_in, const void* W, const float* scale,
                          const float* bias, const void* XH, const float* rstd, void* DZ_out,
                          float* dscale, float* dbias, int M, int K, int HN, int lda, int ldw);
int mlb_cast_f32_bf16(void* stream, const float* src, void* dst, long long n);
/* bf16 copies of an fp32 weight matrix W [rows, cols]: dst_t = W^T (ld_t), dst = W (ld_d, may be NULL) */
int mlb_cast_weight_bf16(void* stream, const float* src, void* dst_t, void* dst, int rows,
                         int cols, int ld_src, int ld_t, int ld_d);

/* ------------------------------------------------------------------------------------ */
/* K11: LSTM cell (flax OptimizedLSTMCell semantics, ml/rnn.py:10-111), fp32.               */
/* z f32 [M, 4H] = x W_i + h W_h (gate blocks i|f|g|o, without bias), bias f32 [4H].        */
/* fwd: h_seq = unmasked output of the step; (c_carry, h_carry) = next-step state, zeroed   */
/*      where ends[m] (LSTM.sequence resets AFTER the step, ml/rnn.py:91-96); ends may be   */
/*      NULL (rollout: the reset is mlb_rnn_reset_f32 after the env step); stash (may be    */
/*      NULL) f32 [M, 5H] = i, f, g, o, tanh(c') for the backward.                           */
/* bwd: dh_seq [M, ld_dh] gradient w.r.t. the unmasked output, dh_carry/dc_carry gradients   */
/*      w.r.t. the masked carry (both NULL at the last step) -> dz [M, 4H], dc_prev [M, H]. */
/* ------------------------------------------------------------------------------------ */
int mlb_lstm_cell_fwd_f32(void* stream, const float* z, const float* bias, const float* c_prev,
                          const uint8_t* ends, float* h_seq, float* c_carry, float* h_carry,
                          float* stash, long long M, int H);
int mlb_lstm_cell_bwd_f32(void* stream, const float* dh_seq, int ld_dh, const float* dh_carry,
                          const float* dc_carry, const uint8_t* ends, const float* stash,
                          const float* c_prev, float* dz, float* dc_prev, long long M, int H);
/* Tensor-core path (compute_dtype = bfloat16; ml/rnn.py:10-111 with the recurrent products on tcgen05   */
/* through mlb_gemm_bf16_tc): same cell math, fp32 cell state, bf16 GEMM operands.  h_seq_bf16 [M, H]     */
/* unmasked output; h_carry (f32, may be NULL) / h_carry_bf16 (may be NULL): masked carry; bias may be    */
/* NULL when the GEMM epilogue already added it.  dz_bf16 [M, 4H].                                        */
int mlb_lstm_cell_fwd_tc(void* stream, const float* z, const float* bias, const float* c_prev,
                         const uint8_t* ends, void* h_seq_bf16, float* c_carry, float* h_carry,
                         void* h_carry_bf16, float* stash, long long M, int H);
int mlb_lstm_cell_bwd_tc(void* stream, const float* dh_seq, int ld_dh, const float* dh_carry,
                         const float* dc_carry, const uint8_t* ends, const float* stash,
                         const float* c_prev, void* dz_bf16, float* dc_prev, long long M, int H);
/* Fused step: z = [x_t | h_{t-1}] [W_i | W_h]^T on tcgen05 with the cell math in the epilogue, */
/* gate pre-activations never written to memory.  w_packed bf16 [4H, in + H], bias_packed f32 [4H]: rows     */
/* permuted to  nb*256 + gate*64 + j  <->  (gate, hidden unit nb*64 + j)  by mlb_lstm_pack_weights_bf16 from  */
/* the arena's W_i^T [4H, in], W_h^T [4H, H], b [4H] (after every optimiser step).  x bf16 [M, in] (ldx),    */
/* h_prev bf16 [M, H]; outputs as mlb_lstm_cell_fwd_tc.  in, H multiples of 64.                               */
int mlb_lstm_pack_weights_bf16(void* stream, const float* wi_t, const float* wh_t, const float* bias,
                               void* w_packed, float* bias_packed, int in_dim, int H);
int mlb_lstm_step_tc(void* stream, const void* x, int ldx, const void* h_prev, const void* w_packed,
                     const float* bias_packed, const float* c_prev, const uint8_t* ends, void* h_seq_bf16,
                     float* c_carry, float* h_carry, void* h_carry_bf16, float* stash, long long M,
                     int in_dim, int H);
/* clear_recurrent_state (ml/rnn.py:66-81): state[m, :] = 0 where dones[m] */
int mlb_rnn_reset_f32(void* stream, float* state, const uint8_t* dones, long long M, int H);

/* ------------------------------------------------------------------------------------ */
/* K7: rollout action sampling.  `head` f32 [rows, ld]: columns [0,sumA) logits of the      */
/* concatenated discrete components, column sumA the critic value.                          */
/* mlb_rollout_keys: (prng_key, step_key) = split(prng_key); policy_key =                   */
/*   split(step_key, 1)[0]  (ml/rollouts.py:878-880, one policy chunk).                     */
/* mlb_sample_discrete_f32: DiscreteActionDistributions.sample (ml/dists.py:26-44):         */
/*   keys = split(policy_key, A); a_i = argmax(logits_i + gumbel(keys_i));                  */
/*   log_prob = logit[a] - logsumexp.  deterministic != 0 -> best() (:46-52).               */
/*   buckets_host: HOST int32[A].  values (may be NULL) <- head[:, sumA].                   */
/* ------------------------------------------------------------------------------------ */
#define MLB_MAX_ACTION_COMPONENTS 16
#define MLB_MAX_CRITIC_BINS 127
int mlb_rollout_keys(void* stream, uint32_t* prng_key, uint32_t* policy_key, int partitionable);
int mlb_sample_discrete_f32(void* stream, const float* head, int ld, const uint32_t* policy_key,
                            const int32_t* buckets_host, int num_components, long long rows,
                            int partitionable, int deterministic, int32_t* actions,
                            float* log_probs, float* values, const float* critic_bins_host,
                            int num_critic_bins);
/* ContinuousActionDistributions.sample / .best (ml/dists.py:216-258): mean = tanh(raw), std = (max - min) *  */
/* sigmoid(raw + 2) + min, action = mean + std * normal(split(policy_key, 1)[0]) (XLA's erf_inv polynomial on   */
/* the threefry uniform; deterministic: the mean), log_probs = Normal log-density.  actions int32 [rows,       */
/* num_dims] receive the fp32 BIT PATTERNS.                                                                    */
int mlb_sample_continuous_f32(void* stream, const float* head, int ld, const uint32_t* policy_key,
                              int num_dims, float stddev_min, float stddev_max, long long rows,
                              int partitionable, int deterministic, int32_t* actions, float* log_probs,
                              float* values, const float* critic_bins_host, int num_critic_bins);
/* critic_bins_host / num_critic_bins: NULL / <=1 for the plain critic (one value column); else  */
/* the HOST array of the DreamerV3 critic's bin centres (odd count, ml/dists.py:127-141): the    */
/* head then carries num_critic_bins critic logits and `values` receives the two-hot mean.       */

/* ------------------------------------------------------------------------------------ */
/* mlb_policy_rollout_tc: one launch for a whole rollout policy step (compute_dtype bf16):  */
/* the PRNG key chain of ml/rollouts.py:878-880 (key_in -> key_out, policy key internal),   */
/* the observation copy into the rollout store (:637-668; obs_store may be NULL),           */
/* ActorCritic.rollout (ml/actor_critic.py:74-96) = num_layers x [Dense, LayerNorm, ReLU]   */
/* + heads, and DiscreteActionDistributions.sample (ml/dists.py:26-44).  A CTA owns 128     */
/* agents; activations stay in shared memory / TMEM between layers.                         */
/* Weights are the bf16 transposed copies W_l^T [hidden, d_l] (d_0 = obs_dim) and           */
/* Wh^T [head_width, hidden] with fp32 LayerNorm scale/bias and head bias; key_in != key_out. */
/* head_out (may be NULL) [rows, head_width] fp32 receives the raw head outputs.            */
/* ------------------------------------------------------------------------------------ */
#define MLB_MLP_TC_MAX_LAYERS 4
typedef struct mlb_mlp_tc_desc {
    int num_layers, obs_dim, hidden, head_width;
    const void* w_t[MLB_MLP_TC_MAX_LAYERS];
    const float* scale[MLB_MLP_TC_MAX_LAYERS];
    const float* bias[MLB_MLP_TC_MAX_LAYERS];
    const void* wh_t;
    const float* head_bias;
} mlb_mlp_tc_desc;
int mlb_policy_rollout_tc(void* stream, const mlb_mlp_tc_desc* desc_host, const float* obs,
                          float* obs_store, long long rows, const uint32_t* key_in,
                          uint32_t* key_out, const int32_t* buckets_host, int num_components,
                          int partitionable, int deterministic, int32_t* actions,
                          float* log_probs, float* values, const float* critic_bins_host,
                          int num_critic_bins, float* head_out);
/* The same launch also doing the PREVIOUS step's "Post Step Rollout Store" (ml/rollouts.py:946-978, */
/* what mlb_post_step_store_f32 does) on the rows each CTA owns: the simulator's rewards / dones of    */
/* step t-1 reach the store inside the policy launch of step t (the bootstrap launch takes the last    */
/* step's), one launch fewer per rollout step.  post_step_host may be NULL (= mlb_policy_rollout_tc).   */
typedef struct mlb_post_step {
    const float* rewards;
    const uint8_t* dones;
    float* reward_slab;
    uint8_t* done_slab;
    float* env_returns;
    float* trace;              /* may be NULL */
    float gamma;
} mlb_post_step;
int mlb_policy_rollout_ps_tc(void* stream, const mlb_mlp_tc_desc* desc_host, const float* obs,
                             float* obs_store, long long rows, const uint32_t* key_in,
                             uint32_t* key_out, const int32_t* buckets_host, int num_components,
                             int partitionable, int deterministic, int32_t* actions,
                             float* log_probs, float* values, const float* critic_bins_host,
                             int num_critic_bins, float* head_out, const mlb_post_step* post_step_host);

/* ------------------------------------------------------------------------------------ */
/* K8: fused PPO loss + gradient w.r.t. the head outputs (ml/ppo.py:129-262 + autodiff).    */
/* rows = T'*M in [T', M] order.  adv_mean_rstd: NULL or device f32[2] (per-minibatch       */
/* z-score, ml/ppo.py:134-143).  vn_params: NULL or device f32[4] = {mu_old, sigma_old,     */
/* mu_new, inv_sigma_new} (value normaliser before / after this minibatch's EMA update,     */
/* ml/ppo.py:190-211).  obj_scale_host[i] = 1/(rows*A_g), ent_scale_host[i] =                */
/* entropy_coef[g]/(rows*A_g) for component i of action group g (HOST arrays).              */
/* d_head [rows, ld] out: f32, or bf16 when flags has MLB_PPO_DHEAD_BF16 (tensor-core path); */
/* padding columns are written as zeros.  d_bias (may be NULL): f32 [sumA+1], the heads' bias  */
/* gradients (column sums of d_head) are ACCUMULATED here.  stats out; ws of                    */
/* mlb_ppo_loss_workspace(rows) bytes, ZERO-INITIALISED once by the caller (it ends with the       */
/* ticket counter of the last-block final reduction).  ld % 4 == 0; head, d_head 16-byte aligned.  */
/* ------------------------------------------------------------------------------------ */
#define MLB_PPO_CLIP_VALUE_LOSS  1
#define MLB_PPO_HUBER_VALUE_LOSS 2
#define MLB_PPO_DHEAD_BF16       4
/* HL-Gauss critic (ml/models.py:177-306): critic_bins_host = centres[V] | bounds[V+1] | smoothness */
#define MLB_PPO_HLGAUSS_CRITIC   8
/* Continuous action group (ContinuousActionDistributions, ml/dists.py:211-284): num_components action   */
/* dimensions, head columns = raw means [0, A) | raw stds [A, 2A) | critic; `actions` holds the fp32 bit  */
/* patterns; buckets_host is ignored; ent_scale_host has A + 2 entries, the last two = stddev_min, _max.  */
#define MLB_PPO_CONTINUOUS_ACTIONS 16
typedef struct mlb_ppo_stats {
    float loss, action_obj, value_loss, entropy;
    mlb_metric metrics[5];   /* Loss, Action Obj, Value Loss, Value Errors (abs), Entropy
                                (the order of PPO.add_metrics, ml/ppo.py:98-104) */
} mlb_ppo_stats;
size_t mlb_ppo_loss_workspace(long long rows);
int mlb_ppo_loss_f32(void* stream, const float* head, int ld, const int32_t* actions,
                     const float* old_log_probs, const float* advantages, const float* returns,
                     const float* old_values, const float* mb_weights,
                     const float* adv_mean_rstd, const float* vn_params,
                     const int32_t* buckets_host, const float* obj_scale_host,
                     const float* ent_scale_host, int num_components, long long rows,
                     long long M, float clip_coef, float value_loss_coef, int flags,
                     void* d_head, float* d_bias, mlb_ppo_stats* stats, void* ws, size_t ws_bytes,
                     const float* critic_bins_host, int num_critic_bins);

/* ------------------------------------------------------------------------------------ */
/* K10: optimiser over a flat fp32 arena (ml/ppo.py:84-90,283-338).                         */
/* ------------------------------------------------------------------------------------ */
typedef struct mlb_segment {
    long long offset, length;   /* in floats, within the parameter arena */
    int32_t kind;               /* 0 none, 1 kernel re-projection, 2 LayerNorm (scale|bias) */
    float target;               /* kind 1: initial L2 norm; kind 2: num_features */
} mlb_segment;
int mlb_fill_zero(void* stream, void* p, size_t bytes);
/* rows x cols block of a row-major fp32 matrix with row stride ld elements (BackboneSeparate: the  */
/* off-diagonal blocks of the fused head gradient, ml/actor_critic.py:247-303)                     */
int mlb_fill_zero_2d(void* stream, float* p, int rows, int cols, int ld);
int mlb_copy_bytes(void* stream, const void* src, void* dst, size_t bytes);
size_t mlb_sumsq_workspace(long long n);
int mlb_sumsq_f32(void* stream, const float* x, long long n, double* out, void* ws, size_t ws_bytes);
/* step: device int32 (number of Adam steps taken so far; incremented by mlb_renorm_segments) */
/* grad_sumsq: device double (sum of squares of `grads`) or NULL; grads are used as         */
/* grad_scale*grads (1/world_size after a sum-allreduce).                                   */
int mlb_adam_step_f32(void* stream, float* params, const float* grads, float* m, float* v,
                      long long n, const int32_t* step, const double* grad_sumsq, float lr,
                      float b1, float b2, float eps, float max_grad_norm, float grad_scale);
/* Optional per-segment bf16 operand copies refreshed in the same pass (tensor-core path):  */
/* the segment is an fp32 matrix [rows, cols]; dst_t = bf16 W^T [cols, rows] (ld_t), dst =   */
/* bf16 W [rows, cols] (ld_d, may be NULL).  dst_t == NULL: no copy for this segment.         */
typedef struct mlb_bf16_copy {
    void* dst_t;
    void* dst;
    int32_t rows, cols, ld_t, ld_d;
} mlb_bf16_copy;
/* One 8-CTA thread-block cluster per segment (DSMEM reduction).  copies_dev may be NULL.    */
int mlb_renorm_segments(void* stream, float* params, const mlb_segment* segments_dev,
                        int num_segments, int32_t* step, const mlb_bf16_copy* copies_dev);
/* The three steps above in ONE launch (clip_by_global_norm + adam + re-projection / LayerNorm   */
/* renorm + bf16 refresh): a co-resident grid with two device-wide barriers.  Segments must tile */
/* the arena exactly in offset order (kind 0 for plain tensors; <= 32).  grad_sumsq: device      */
/* double, computed here unless have_sumsq (the data-parallel all-reduce kernel already wrote     */
/* it).  sync_state: device uint32[2], zero-initialised; ws of mlb_optimizer_fused_workspace().   */
/* zero_after (may be NULL): an n-float arena (the local gradient arena) cleared at the end of    */
/* the step, ready for the next minibatch's accumulation.                                         */
size_t mlb_optimizer_fused_workspace(void);
int mlb_optimizer_step_fused(void* stream, float* params, const float* grads, float* m, float* v,
                             long long n, const mlb_segment* segments_dev, int num_segments,
                             const mlb_bf16_copy* copies_dev, int32_t* step, double* grad_sumsq,
                             int have_sumsq, float lr, float b1, float b2, float eps,
                             float max_grad_norm, float grad_scale, uint32_t* sync_state, void* ws,
                             size_t ws_bytes, float* zero_after);
/* out[c] += sum_r x[r, c] for c < ncols (bias gradients of the heads) */
int mlb_colsum_f32(void* stream, const float* x, long long rows, int ld, int ncols, float* out);
int mlb_colsum_bf16(void* stream, const void* x, long long rows, int ld, int ncols, float* out);

/* ------------------------------------------------------------------------------------ */
/* PBT policy-batch reorder: _compute_reorder_chunks                        */
/* (ml/rollouts.py:1107-1190) as a stable counting sort.  assignments i32 [S] in [0, P);        */
/* to_policy i32 [B*C] (B chunks of C; B*C >= S), to_sim i32 [S].  Integer-only, bit-exact      */
/* against the reference's KAT vectors (tests/test_rollouts.py:58-81).                          */
/* mlb_gather_rows_clip: out[k] = src[clip(idx[k], 0, n_src-1)] -- PolicyBatchReorderState.     */
/* to_policy / to_sim (ml/rollouts.py:143-168).                                                 */
/* ------------------------------------------------------------------------------------ */
size_t mlb_reorder_chunks_workspace(long long S, int P);
int mlb_reorder_chunks(void* stream, const int32_t* assignments, long long S, int P, int C,
                       long long B, int32_t* to_policy, int32_t* to_sim, void* ws, size_t ws_bytes);
int mlb_gather_rows_clip(void* stream, const void* src, const int32_t* idx, void* out,
                         long long n_idx, long long n_src, long long row_bytes);

/* ------------------------------------------------------------------------------------ */
/* Data-parallel gradient exchange: SUM all-reduce of the flat gradient arena     */
/* over NVLink peer memory fused with the clip_by_global_norm reduction (ml/ppo.py:84-90).     */
/* Every rank's arena and a signal region (uint32[2*MLB_MAX_PEERS], zero-initialised) live in  */
/* peer-mapped (symmetric) memory; peers_host lists all ranks' pointers as mapped in THIS      */
/* process.  out <- sum over ranks (rank order: bit-identical on every rank); sumsq_out (may   */
/* be NULL) <- sum(out^2) in fp64; state = device uint32[2] (zero-initialised, private to this */
/* rank); ws of mlb_allreduce_workspace() bytes.  Every rank must enqueue the call the same    */
/* number of times; the arena may be rewritten as soon as the call's kernel completes.          */
/* ------------------------------------------------------------------------------------ */
#define MLB_MAX_PEERS 16
typedef struct mlb_peer_table {
    int rank, world;
    const float* grads[MLB_MAX_PEERS];
    uint32_t* signals[MLB_MAX_PEERS];
} mlb_peer_table;
size_t mlb_allreduce_workspace(void);
/* NVLS variant (NVSwitch in-network reduction): mc_grads / mc_out are the MULTICAST addresses of  */
/* the arena and of a symmetric `reduced` buffer, out_local this rank's unicast view of the latter */
/* (what the optimiser then reads).  Rank r reduces slice r with multimem.ld_reduce and broadcasts  */
/* it with multimem.st.  state = device uint32[4], zero-initialised.                                */
int mlb_allreduce_nvls_f32(void* stream, const mlb_peer_table* peers_host, const float* mc_grads,
                           float* mc_out, const float* out_local, long long n, double* sumsq_out,
                           uint32_t* state, void* ws, size_t ws_bytes);
int mlb_allreduce_sumsq_f32(void* stream, const mlb_peer_table* peers_host, float* out,
                            long long n, double* sumsq_out, uint32_t* state, void* ws,
                            size_t ws_bytes);

/* ------------------------------------------------------------------------------------ */
/* Synthetic vector environment (stand-in for sim_fns['step'], ml/rollouts.py:905-936).     */
/* tcount: device int32[2] = {step counter, block-arrival scratch}; the step kernel advances  */
/* the counter itself (last block to finish), so a step is ONE launch.                        */
/* ------------------------------------------------------------------------------------ */
int mlb_synth_env_init(void* stream, float* obs, long long N, int D, uint32_t seed,
                       int32_t* tcount);
int mlb_synth_env_step(void* stream, const float* obs_in, float* obs_out, const int32_t* actions,
                       int A, float* rewards, uint8_t* dones, int32_t* tcount, long long N,
                       int D, uint32_t seed, float p_done);

#ifdef __cplusplus
}
#endif
#endif /* MLB200_H_ */
