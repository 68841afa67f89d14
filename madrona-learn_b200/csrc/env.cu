// Synthetic vector environment (our stand-in for the external simulator behind the
// reference's sim_fns['step'] boundary, ml/rollouts.py:905-936; SURVEY 8d "Synthetic inputs").
//
//   reward = obs[:,0] * (act[:,0] - 1.5) * 0.1                (uses the obs the policy saw)
//   done   = Bernoulli(p_done) from threefry bits, or the deterministic parity variant
//            done = ((t + n) % 61 == 0)  when p_done < 0
//   obs'   = 0.9 * obs + 0.1 * xi      xi ~ U[-sqrt3, sqrt3) (unit variance), or a fresh xi
//            when the episode ended
// Noise is a pure function of (seed, t, n, d) through threefry and an exact int->float
// conversion, and the arithmetic is unfused fp32, so oracle/env.py reproduces it bit for bit.
#include "common.cuh"

namespace {

__host__ __device__ __forceinline__ float env_noise(uint32_t bits) {
    // 24-bit uniform in [0,1) -> [-sqrt3, sqrt3)
    const float u = (float)(bits >> 8) * 5.9604644775390625e-08f;   // 2^-24, exact
#ifdef __CUDA_ARCH__
    return __fadd_rn(__fmul_rn(u, 3.4641016151377544f), -1.7320508075688772f);
#else
    return u * 3.4641016151377544f + -1.7320508075688772f;
#endif
}

// Step-counter tick folded into the step kernel: every block has read tcount[0] before it
// arrives here, so the LAST block to arrive (tcount[1] = arrival counter) advances the clock.
__device__ __forceinline__ void env_tick_last_block(int32_t* tcount, int t) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&tcount[1], 1) == (int)gridDim.x - 1) {
            tcount[1] = 0;
            tcount[0] = t + 1;
        }
    }
}

__global__ void __launch_bounds__(256)
synth_env_step_kernel(const float* obs_in, float* obs_out,   // may alias (in-place step)
                      const int32_t* __restrict__ actions, int A, float* __restrict__ rewards,
                      uint8_t* __restrict__ dones, int32_t* tcount, long long N,
                      int D, uint32_t seed, float p_done) {
    pdl_launch_dependents();
    pdl_wait();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = *tcount;     // L1 is invalidated at launch boundaries; only the last block writes it
    if (e < N * D) {
        const long long n = e / D;
        const int d = (int)(e - n * D);
        bool done;
        if (p_done < 0.f) done = ((t + n) % 61) == 0;
        else {
            uint32_t x0 = (uint32_t)n, x1 = 0x80000000u | (uint32_t)t;
            threefry2x32(seed, 0x9E3779B9u, x0, x1);
            done = (float)(x0 >> 8) * 5.9604644775390625e-08f < p_done;
        }
        uint32_t x0 = (uint32_t)e, x1 = (uint32_t)t;
        threefry2x32(seed, (uint32_t)(e >> 32), x0, x1);
        const float xi = env_noise(x0);
        const float o = obs_in[e];
        if (d == 0) {
            const float a0 = (float)actions[n * A];
            rewards[n] = __fmul_rn(__fmul_rn(o, __fadd_rn(a0, -1.5f)), 0.1f);
            dones[n] = done ? 1 : 0;
        }
        obs_out[e] = done ? xi : __fadd_rn(__fmul_rn(0.9f, o), __fmul_rn(0.1f, xi));
    }
    env_tick_last_block(tcount, t);
}

__global__ void __launch_bounds__(256)
synth_env_init_kernel(float* __restrict__ obs, long long total, uint32_t seed) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    uint32_t x0 = (uint32_t)e, x1 = 0xFFFFFFFFu;
    threefry2x32(seed, (uint32_t)(e >> 32), x0, x1);
    obs[e] = env_noise(x0);
}

}  // namespace

MLB_API int mlb_synth_env_init(void* stream, float* obs, long long N, int D, uint32_t seed,
                               int32_t* tcount) {
    MLB_REQUIRE(obs && tcount && N > 0 && D > 0);
    cudaStream_t s = mlb_stream(stream);
    cudaError_t e = cudaMemsetAsync(tcount, 0, 2 * sizeof(int32_t), s);
    if (e != cudaSuccess) return (int)e;
    synth_env_init_kernel<<<mlb_cdiv(N * D, 256), 256, 0, s>>>(obs, N * D, seed);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_synth_env_step(void* stream, const float* obs_in, float* obs_out,
                               const int32_t* actions, int A, float* rewards, uint8_t* dones,
                               int32_t* tcount, long long N, int D, uint32_t seed, float p_done) {
    MLB_REQUIRE(obs_in && obs_out && actions && rewards && dones && tcount && N > 0 && D > 0 && A > 0);
    cudaStream_t s = mlb_stream(stream);
    cudaError_t e = launch_pdl(synth_env_step_kernel, dim3(mlb_cdiv(N * D, 256)), dim3(256), 0, s, obs_in, obs_out, actions, A,
                               rewards, dones, tcount, N, D, seed, p_done);
    if (e != cudaSuccess) return (int)e;
    return MLB_OK;
}
