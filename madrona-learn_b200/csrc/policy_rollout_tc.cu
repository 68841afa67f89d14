// One-kernel rollout policy step on the tensor cores (compute_dtype=bfloat16).
//
// Replaces, for one rollout step (ml/rollouts.py:877-896 "Policy Inference" + the obs part of
// "Pre Step Rollout Store" :637-668): the PRNG key chain, the copy of the observations into
// the rollout store, ActorCritic.rollout (ml/actor_critic.py:74-96) = L x [Dense -> LayerNorm
// -> ReLU] + actor/critic heads, and DiscreteActionDistributions.sample (ml/dists.py:26-44).
//
// A CTA owns 128 agents for the whole network: the activations never leave the SM.
//   * epilogue warps load the fp32 observation rows, write them to the store slab and, as
//     bf16, into shared memory in the canonical K-major SWIZZLE_128B operand layout;
//   * warp 0 streams every layer's W^T through a 2-stage TMA ring (weights are L2-resident);
//   * warp 1 issues tcgen05.mma with the A operand taken from the activation panels;
//   * the epilogue warps (16 = 4 TMEM lane quadrants x 4 column groups) run LayerNorm + ReLU out
//     of TMEM and write the next layer's A operand straight back into the same panels;
//   * the head GEMM lands in TMEM columns [256, 256+NH), is staged as fp32 in shared memory and
//     sampled by all 512 epilogue threads ((row, component) work items, threefry Gumbel-max).
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int PR_THREADS = 576;          // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue (4 lane quadrants x 4 column groups)
constexpr int PR_EPI = 512;
constexpr int PR_MAX_STAGES = 4;       // weight ring depth is chosen at launch: as deep as shared memory allows
constexpr int MAXL = MLB_MLP_TC_MAX_LAYERS;
constexpr int MAXC = MLB_MAX_ACTION_COMPONENTS;
constexpr float LN_EPS = 1e-6f;

struct PRMaps {
    CUtensorMap w[MAXL];        // W_l^T  [H, d_l]   box {64, H}
    CUtensorMap wh;             // Wh^T   [NH, H]    box {64, NH}
};

struct PRArgs {
    int L, D, H, NH, A, V, vcol;
    int off[MAXC], nb[MAXC];
    const float* scale[MAXL];
    const float* bias[MAXL];
    const float* head_bias;
    float bins[MLB_MAX_CRITIC_BINS];
    // optional: the PREVIOUS step's "Post Step Rollout Store" (ml/rollouts.py:946-978) done by this launch
    const float* ps_r;
    const uint8_t* ps_d;
    float* ps_rs;
    uint8_t* ps_ds;
    float* ps_er;
    float* ps_trace;
    float ps_gamma;
};

__device__ __forceinline__ float uniform_from_bits(uint32_t bits) {
    const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
    const float tiny = 1.17549435e-38f;
    return fmaxf(tiny, f * (1.0f - tiny) + tiny);
}

__global__ void __launch_bounds__(PR_THREADS, 1)
policy_rollout_kernel(const __grid_constant__ PRMaps maps, const __grid_constant__ PRArgs a,
                      const float* __restrict__ obs, float* __restrict__ obs_store, long long rows,
                      const uint32_t* __restrict__ key_in, uint32_t* __restrict__ key_out, int part,
                      int deterministic, int32_t* __restrict__ actions, float* __restrict__ log_probs,
                      float* __restrict__ values, float* __restrict__ head_out, int PR_STAGES) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_launch_dependents();
    uint8_t* smem = align_smem_1024(smem_raw);
    const int H = a.H, NH = a.NH, L = a.L, D = a.D;
    const int act_panels = (H > D ? H : D + 63) / 64;                 // panels of [128 x 64] bf16
    uint8_t* act = smem;                                              // act_panels * 16 KB
    uint8_t* ring = act + act_panels * 16384;                         // PR_STAGES * (H * 128 B)
    const int stage_bytes = H * 128;
    float* head_sm = reinterpret_cast<float*>(ring + PR_STAGES * stage_bytes);   // [128][NH + 1]
    float* fsm = head_sm + 128 * (NH + 1);                            // [2][2H] scale|bias, [4][128][2] partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(fsm + 4 * H + 1024);
    uint64_t* full_bar = bars;                   // [PR_STAGES]
    uint64_t* empty_bar = bars + PR_MAX_STAGES;  // [PR_STAGES]
    uint64_t* acc_bar = bars + 2 * PR_MAX_STAGES;   // accumulator of the current layer complete
    uint64_t* a_bar = acc_bar + 1;               // A operand of the next layer ready (256 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
    uint32_t* keys_sm = tmem_slot + 2;           // policy key (2 words)
    uint32_t* ckeys_sm = keys_sm + 2;            // per-component sample keys (2 * A words)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * BM;
    const int kb_of_layer0 = (D + 63) / 64, kb_h = H / 64;
    // Row pairing: jax's (non-partitionable) random_bits feeds element e and element e + size/2 from
    // ONE threefry block, i.e. rows r and r + rows/2 of the same component share their random
    // blocks.  A CTA therefore owns 64 rows of the first half of the batch and the matching 64
    // rows of the second half, and every threefry evaluation yields two samples.
    const bool paired = (rows % BM == 0) && !part;
    const long long half_rows = rows >> 1;
    auto grow = [&](int r) -> long long {
        return paired ? (r < 64 ? (long long)blockIdx.x * 64 + r : half_rows + (long long)blockIdx.x * 64 + (r - 64))
                      : m0 + r;
    };

    if (warp == 0 && lane == 0) {
        for (int l = 0; l < L; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[l])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.wh)) : "memory");
        for (int s = 0; s < PR_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        mbar_init(a_bar, PR_EPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // PRNG key chain (ml/rollouts.py:878-880): (prng_key, step_key) = split(prng_key);
        // policy_key = split(step_key, 1)[0].  Every CTA derives it; CTA 0 publishes the new key.
        pdl_wait();                              // the key was advanced by the previous step's kernel
        if (!deterministic) {
            const uint32_t k0 = key_in[0], k1 = key_in[1];
            uint32_t n0, n1, s0, s1, p0, p1;
            threefry_split_at(k0, k1, 0, 2, part, n0, n1);
            threefry_split_at(k0, k1, 1, 2, part, s0, s1);
            threefry_split_at(s0, s1, 0, 1, part, p0, p1);
            keys_sm[0] = p0; keys_sm[1] = p1;
            if (blockIdx.x == 0) { key_out[0] = n0; key_out[1] = n1; }
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    pdl_wait();                                  // observations / stores below belong to this step
    const uint32_t tmem_base = *tmem_slot;
    if (a.ps_r != nullptr && threadIdx.x >= PR_THREADS - BM) {
        // the simulator's rewards / dones of the previous step -> rollout store, discounted env-return trace
        // (same arithmetic as post_step_kernel); the last four epilogue warps, idle until the first accumulator
        const long long i = m0 + (threadIdx.x - (PR_THREADS - BM));
        if (i < rows) {
            const float ri = a.ps_r[i];
            const uint8_t di = a.ps_d[i];
            a.ps_rs[i] = ri;
            a.ps_ds[i] = di;
            const float v = __fadd_rn(ri, __fmul_rn(a.ps_gamma, a.ps_er[i]));
            if (a.ps_trace) a.ps_trace[i] = v;
            a.ps_er[i] = di ? 0.f : v;
        }
    }

    if (warp == 0) {
        // ================= weight producer: all layers + heads through the ring =================
        if (lane == 0) {
            int it = 0;
            for (int l = 0; l <= L; ++l) {
                const int nkb = l == 0 ? kb_of_layer0 : kb_h;
                const CUtensorMap* map = l < L ? &maps.w[l] : &maps.wh;
                const uint32_t bytes = (uint32_t)(l < L ? H : NH) * 128u;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % PR_STAGES;
                    mbar_wait(&empty_bar[s], ((it / PR_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[s], bytes);
                    tma_load_2d(map, &full_bar[s], ring + s * stage_bytes, kb * 64, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int it = 0;
            for (int l = 0; l <= L; ++l) {
                const int nkb = l == 0 ? kb_of_layer0 : kb_h;
                const int n = l < L ? H : NH;
                const uint32_t idesc = umma_idesc(false, false, n);
                const uint32_t d_tmem = tmem_base + (l < L ? 0u : 256u);
                mbar_wait(a_bar, (uint32_t)(l & 1));          // A panels of layer l written, TMEM drained
                tcgen05_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % PR_STAGES;
                    mbar_wait(&full_bar[s], (it / PR_STAGES) & 1);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(act + kb * 16384);
                    const uint32_t sb = smem_u32(ring + s * stage_bytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        tcgen05_mma_f16(d_tmem, umma_desc(sa + k * 32, 16, 1024), umma_desc(sb + k * 32, 16, 1024),
                                        idesc, (kb | k) ? 1u : 0u);
                    tcgen05_commit(&empty_bar[s]);
                }
                tcgen05_commit(acc_bar);
            }
        }
    } else {
        // ================= epilogue warps =================
        const int quad = warp & 3, grp = (warp - 2) >> 2;
        const int rt = quad * 32 + lane;
        const int et = threadIdx.x - 64;                 // 0..511
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        float* partials = fsm + 4 * H;
        if (!deterministic && et < a.A) {                // sample_keys = split(policy_key, A)  (ml/dists.py:31)
            uint32_t c0, c1;
            threefry_split_at(keys_sm[0], keys_sm[1], (uint32_t)et, (uint32_t)a.A, part, c0, c1);
            ckeys_sm[2 * et] = c0; ckeys_sm[2 * et + 1] = c1;
        }
        // ---- observations: fp32 rows -> store slab (fp32) + bf16 A panels (SWIZZLE_128B, K-major) ----
        {
            const int chunks = kb_of_layer0 * 8;         // 16-byte bf16 chunks per row (zero padded)
            for (int item = et; item < 128 * chunks; item += PR_EPI) {
                const int r = item / chunks, j = item - r * chunks;
                const long long row = grow(r);
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (row < rows && j * 8 < D) {
                    const float* src = obs + row * D + j * 8;
                    v0 = __ldg(reinterpret_cast<const float4*>(src));
                    v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                    if (obs_store) {
                        float* dst = obs_store + row * D + j * 8;
                        *reinterpret_cast<float4*>(dst) = v0;
                        *(reinterpret_cast<float4*>(dst) + 1) = v1;
                    }
                }
                *reinterpret_cast<uint4*>(act + (j >> 3) * 16384 + sw128(r, j & 7)) =
                    make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
            }
            fence_async_smem();
            mbar_arrive(a_bar);
        }
        const int nchunks = H / 32;
        for (int l = 0; l < L; ++l) {
            float* s = fsm + (l & 1) * 2 * H;
            float* b = s + H;
            for (int i = et; i < H; i += PR_EPI) { s[i] = a.scale[l][i]; b[i] = a.bias[l][i]; }
            mbar_wait(acc_bar, (uint32_t)(l & 1));
            tcgen05_fence_after();
            float sum = 0.f, sq = 0.f;
            for (int ch = grp; ch < nchunks; ch += 4) {
                uint32_t r[32];
                tmem_ld32(taddr + ch * 32, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float z = __uint_as_float(r[j]); sum += z; sq = fmaf(z, z, sq); }
            }
            {
                float2* pp = reinterpret_cast<float2*>(partials);
                pp[grp * 128 + rt] = make_float2(sum, sq);
                named_bar_sync(3, PR_EPI);             // also orders the s/b loads above
                const float2 p0 = pp[rt], p1 = pp[128 + rt], p2 = pp[256 + rt], p3 = pp[384 + rt];
                sum = (p0.x + p1.x) + (p2.x + p3.x);
                sq = (p0.y + p1.y) + (p2.y + p3.y);
            }
            const float invH = 1.f / (float)H;
            const float mean = sum * invH;
            const float rstd = rsqrtf(fmaxf(0.f, sq * invH - mean * mean) + LN_EPS);
            for (int ch = grp; ch < nchunks; ch += 4) {
                const int c = ch * 32;
                uint32_t r[32];
                tmem_ld32(taddr + c, r);
                uint8_t* pan = act + (c >> 6) * 16384;
                const int hf = ch & 1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float y[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = 8 * q + e;
                        y[e] = fmaxf(0.f, fmaf((__uint_as_float(r[j]) - mean) * rstd, s[c + j], b[c + j]));
                    }
                    *reinterpret_cast<uint4*>(pan + sw128(rt, hf * 4 + q)) =
                        make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                }
            }
            tcgen05_fence_before();
            fence_async_smem();
            mbar_arrive(a_bar);                           // next layer's A operand is in place
        }
        // ---- heads: TMEM columns [256, 256+NH) -> fp32 (+bias) tile in shared memory ----
        mbar_wait(acc_bar, (uint32_t)(L & 1));
        tcgen05_fence_after();
        {
            const int hch = NH / 32;                      // NH is a multiple of 64
            for (int ch = grp; ch < hch; ch += 4) {
                uint32_t r[32];
                tmem_ld32(taddr + 256 + ch * 32, r);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    head_sm[rt * (NH + 1) + ch * 32 + j] = __uint_as_float(r[j]) + a.head_bias[ch * 32 + j];
            }
        }
        named_bar_sync(3, PR_EPI);
        if (head_out) {
            for (int e = et; e < 128 * NH; e += PR_EPI) {
                const int r = e / NH, c = e - r * NH;
                if (grow(r) < rows) head_out[grow(r) * NH + c] = head_sm[r * (NH + 1) + c];
            }
        }
        // ---- sampling: (row, component) work items over the 512 epilogue threads ----
        auto lse_of = [&](const float* l, int off, int nb) {
            float mx = -INFINITY;
            for (int j = 0; j < nb; ++j) mx = fmaxf(mx, l[off + j]);
            float se = 0.f;
            for (int j = 0; j < nb; ++j) se += expf(l[off + j] - mx);
            return logf(se) + mx;
        };
        auto emit = [&](long long row, const float* l, int i, int off, int best, float lse) {
            actions[row * a.A + i] = best;
            if (log_probs) log_probs[row * a.A + i] = l[off + best] - lse;
            if (values && i == 0) {
                if (a.V == 1) values[row] = l[a.vcol];
                else {        // two-hot mean (ml/dists.py:143-170)
                    const float* lc = l + a.vcol;
                    const int V = a.V, mid = (V - 1) / 2;
                    float m2 = -INFINITY;
                    for (int k = 0; k < V; ++k) m2 = fmaxf(m2, lc[k]);
                    float s2 = 0.f;
                    for (int k = 0; k < V; ++k) s2 += expf(lc[k] - m2);
                    const float inv = 1.f / s2;
                    float acc = 0.f;
                    for (int k = 0; k < mid; ++k)
                        acc += expf(lc[mid - 1 - k] - m2) * inv * a.bins[mid - 1 - k] + expf(lc[mid + 1 + k] - m2) * inv * a.bins[mid + 1 + k];
                    values[row] = expf(lc[mid] - m2) * inv * a.bins[mid] + acc;
                }
            }
        };
        if (paired && !deterministic) {
            for (int item = et; item < 64 * a.A; item += PR_EPI) {
                const int r = item / a.A, i = item - r * a.A;
                const long long rowA = grow(r), rowB = rowA + half_rows;
                const float* lA = head_sm + r * (NH + 1);
                const float* lB = head_sm + (r + 64) * (NH + 1);
                const int off = a.off[i], nb = a.nb[i];
                const float lseA = lse_of(lA, off, nb), lseB = lse_of(lB, off, nb);
                const uint32_t c0 = ckeys_sm[2 * i], c1 = ckeys_sm[2 * i + 1];
                const uint64_t half = (uint64_t)half_rows * (uint64_t)nb;      // = size / 2 (size is even)
                int bestA = 0, bestB = 0;
                float bvA = -INFINITY, bvB = -INFINITY;
                for (int j = 0; j < nb; ++j) {
                    const uint64_t idx = (uint64_t)rowA * nb + j;
                    uint32_t x0 = (uint32_t)idx, x1 = (uint32_t)(idx + half);
                    threefry2x32(c0, c1, x0, x1);               // one block: x0 -> row A, x1 -> row A + rows/2
                    const float vA = -logf(-logf(uniform_from_bits(x0))) + lA[off + j];
                    const float vB = -logf(-logf(uniform_from_bits(x1))) + lB[off + j];
                    if (vA > bvA) { bvA = vA; bestA = j; }
                    if (vB > bvB) { bvB = vB; bestB = j; }
                }
                emit(rowA, lA, i, off, bestA, lseA);
                emit(rowB, lB, i, off, bestB, lseB);
            }
        } else {
            for (int item = et; item < 128 * a.A; item += PR_EPI) {
                const int r = item / a.A, i = item - r * a.A;
                const long long row = grow(r);
                if (row >= rows) continue;
                const float* l = head_sm + r * (NH + 1);
                const int off = a.off[i], nb = a.nb[i];
                const float lse = lse_of(l, off, nb);
                int best = 0;
                float bv = -INFINITY;
                if (deterministic) {
                    for (int j = 0; j < nb; ++j) if (l[off + j] > bv) { bv = l[off + j]; best = j; }
                } else {
                    const uint32_t c0 = ckeys_sm[2 * i], c1 = ckeys_sm[2 * i + 1];
                    const uint64_t size = (uint64_t)rows * (uint64_t)nb;
                    for (int j = 0; j < nb; ++j) {
                        const uint32_t bits = threefry_bits_at(c0, c1, (uint64_t)row * nb + j, size, part);
                        const float v = -logf(-logf(uniform_from_bits(bits))) + l[off + j];
                        if (v > bv) { bv = v; best = j; }
                    }
                }
                emit(row, l, i, off, best, lse);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

MLB_API int mlb_policy_rollout_tc(void* stream, const mlb_mlp_tc_desc* d, const float* obs,
                                  float* obs_store, long long rows, const uint32_t* key_in,
                                  uint32_t* key_out, const int32_t* buckets_host, int num_components,
                                  int partitionable, int deterministic, int32_t* actions,
                                  float* log_probs, float* values, const float* critic_bins_host,
                                  int num_critic_bins, float* head_out) {
    return mlb_policy_rollout_ps_tc(stream, d, obs, obs_store, rows, key_in, key_out, buckets_host, num_components,
                                    partitionable, deterministic, actions, log_probs, values, critic_bins_host,
                                    num_critic_bins, head_out, nullptr);
}

MLB_API int mlb_policy_rollout_ps_tc(void* stream, const mlb_mlp_tc_desc* d, const float* obs,
                                     float* obs_store, long long rows, const uint32_t* key_in,
                                     uint32_t* key_out, const int32_t* buckets_host, int num_components,
                                     int partitionable, int deterministic, int32_t* actions,
                                     float* log_probs, float* values, const float* critic_bins_host,
                                     int num_critic_bins, float* head_out, const mlb_post_step* ps) {
    MLB_REQUIRE(d && obs && actions && rows > 0 && buckets_host);
    MLB_REQUIRE(!ps || (ps->rewards && ps->dones && ps->reward_slab && ps->done_slab && ps->env_returns));
    MLB_REQUIRE(deterministic || (key_in && key_out));
    MLB_REQUIRE(d->num_layers >= 1 && d->num_layers <= MAXL && d->hidden >= 64 && d->hidden <= 256 &&
                d->hidden % 64 == 0 && d->obs_dim % 8 == 0 && d->obs_dim >= 8 && d->obs_dim <= 256 &&
                d->head_width % 64 == 0 && d->head_width <= 256);
    MLB_REQUIRE(num_components > 0 && num_components <= MAXC && mlb_aligned16(obs) &&
                (!obs_store || mlb_aligned16(obs_store)));
    PRMaps maps;
    PRArgs a;
    a.L = d->num_layers; a.D = d->obs_dim; a.H = d->hidden; a.NH = d->head_width; a.A = num_components;
    int off = 0;
    for (int i = 0; i < num_components; ++i) { a.off[i] = off; a.nb[i] = buckets_host[i]; off += buckets_host[i]; }
    a.V = (num_critic_bins > 1 && critic_bins_host) ? num_critic_bins : 1;
    if (a.V > MLB_MAX_CRITIC_BINS || off + a.V > a.NH) return MLB_EINVAL;
    for (int k = 0; k < a.V && a.V > 1; ++k) a.bins[k] = critic_bins_host[k];
    a.vcol = off;
    a.ps_r = ps ? ps->rewards : nullptr;
    a.ps_d = ps ? ps->dones : nullptr;
    a.ps_rs = ps ? ps->reward_slab : nullptr;
    a.ps_ds = ps ? ps->done_slab : nullptr;
    a.ps_er = ps ? ps->env_returns : nullptr;
    a.ps_trace = ps ? ps->trace : nullptr;
    a.ps_gamma = ps ? ps->gamma : 0.f;
    int rc;
    for (int l = 0; l < a.L; ++l) {
        const int dl = l == 0 ? a.D : a.H;
        if ((rc = make_map(&maps.w[l], d->w_t[l], dl, a.H, dl, 64, a.H))) return rc;
        a.scale[l] = d->scale[l]; a.bias[l] = d->bias[l];
    }
    if ((rc = make_map(&maps.wh, d->wh_t, a.H, a.NH, a.H, 64, a.NH))) return rc;
    a.head_bias = d->head_bias;
    const int act_panels = (a.H > a.D ? a.H : a.D + 63) / 64;
    const size_t fixed = (size_t)act_panels * 16384 + (size_t)128 * (a.NH + 1) * 4 + (size_t)(4 * a.H + 1024) * 4 +
                         16 * 8 + 64 + 128 + 1024;
    // measured: a deeper ring (3-4 stages) does not help -- the per-layer chain is MMA -> two-pass
    // epilogue, not weight-load latency (21.6 us/step with 2 stages, 22.6 with 4)
    static const int want = [] { const char* v = getenv("MLB_ROLLOUT_STAGES"); return v ? atoi(v) : 2; }();
    int stages = 2;
    while (stages < want && stages < PR_MAX_STAGES && fixed + (size_t)(stages + 1) * a.H * 128 <= 227 * 1024) ++stages;
    const size_t smem = fixed + (size_t)stages * a.H * 128;
    if (smem > 227 * 1024) return MLB_EINVAL;
    cudaError_t e = cudaFuncSetAttribute(policy_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(policy_rollout_kernel, dim3(mlb_cdiv(rows, BM)), dim3(PR_THREADS), smem, mlb_stream(stream),
        maps, a, obs, obs_store, rows, key_in, key_out, partitionable, deterministic, actions, log_probs,
        values, head_out, stages);
    if (e != cudaSuccess) return (int)e;
    return MLB_OK;
}
