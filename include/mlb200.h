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
 *   - Re-entrant; no mutable global state; callable from any host thread.
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
/*   vn_mu_sigma : NULL, or device f32[2] = {mu, sigma}: values/bootstrap are stored      */
/*                 normalised and are inverted (v*sigma+mu) on load; the un-normalised     */
/*                 values are what `returns` and the Values metric see.                    */
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
int mlb_mb_moments_f32(void* stream, const double* traj_moments, const int32_t* perm,
                       int E, long long J, long long M, int Tp, float var_floor, float* out);

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

#ifdef __cplusplus
}
#endif
#endif /* MLB200_H_ */
