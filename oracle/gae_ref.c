/* Oracle (TEST INFRASTRUCTURE; see oracle/__init__.py): plain-C restatement of the
 * reference's GAE / returns recurrences, for the CPU baseline timing of the GAE sweep and as
 * a second, independent check of oracle/algo_common.py.
 *   compute_advantages  /root/reference/src/madrona_learn/algo_common.py:84-130
 *   compute_returns     /root/reference/src/madrona_learn/algo_common.py:45-81
 * Same structure as the reference: reverse loop over T, vectorised over the N columns
 * (the inner loop).  Each call handles the column range [n_begin, n_end); oracle/cgae.py
 * spreads ranges over the host cores with Python threads (ctypes drops the GIL; libgomp is
 * not in the image).  float32, no FMA contraction (-ffp-contract=off) so results are
 * bit-identical to the NumPy restatement.
 * Build: make -C oracle   (gcc -O3 -ffp-contract=off -shared -fPIC)
 */
#include <stddef.h>
#include <stdint.h>

void oracle_gae_f32(const float* rewards, const float* values, const uint8_t* dones,
                    const float* bootstrap, float* adv, float* ret, int T, long long N,
                    float gamma, float gamma_lambda, long long n_begin, long long n_end) {
    for (long long n0 = n_begin; n0 < n_end; n0 += 1024) {
        const long long n1 = n0 + 1024 < n_end ? n0 + 1024 : n_end;
        float na[1024], nv[1024];
        for (long long n = n0; n < n1; ++n) { na[n - n0] = 0.f; nv[n - n0] = bootstrap[n]; }
        for (int i = T - 1; i >= 0; --i) {
            const float* r = rewards + (size_t)i * N;
            const float* v = values + (size_t)i * N;
            const uint8_t* d = dones + (size_t)i * N;
            float* a = adv + (size_t)i * N;
            for (long long n = n0; n < n1; ++n) {
                const float nvj = d[n] ? 0.f : nv[n - n0];        /* :112 */
                const float naj = d[n] ? 0.f : na[n - n0];        /* :113 */
                const float td = (r[n] + gamma * nvj) - v[n];     /* :116 */
                const float cur = td + gamma_lambda * naj;        /* :120 */
                a[n] = cur;
                if (ret) ret[(size_t)i * N + n] = cur + v[n];     /* ml/rollouts.py:769 */
                na[n - n0] = cur;                                 /* :124 */
                nv[n - n0] = v[n];
            }
        }
    }
}

void oracle_returns_f32(const float* rewards, const uint8_t* dones, const float* bootstrap,
                        float* ret, int T, long long N, float gamma, long long n_begin,
                        long long n_end) {
    for (long long n0 = n_begin; n0 < n_end; n0 += 1024) {
        const long long n1 = n0 + 1024 < n_end ? n0 + 1024 : n_end;
        float nr[1024];
        for (long long n = n0; n < n1; ++n) nr[n - n0] = bootstrap[n];
        for (int i = T - 1; i >= 0; --i) {
            const float* r = rewards + (size_t)i * N;
            const uint8_t* d = dones + (size_t)i * N;
            for (long long n = n0; n < n1; ++n) {
                const float cur = r[n] + gamma * (d[n] ? 0.f : nr[n - n0]);   /* :70-72 */
                ret[(size_t)i * N + n] = cur;
                nr[n - n0] = cur;
            }
        }
    }
}
