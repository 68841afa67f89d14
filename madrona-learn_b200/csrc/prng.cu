// JAX-compatible Threefry PRNG on device: split / random bits / the PPO minibatch
// permutations (bit-exact with jax.random.permutation's sort-based shuffle).
//
// Replaces random.split / random.permutation at ml/ppo.py:445-458 and
// ml/train_state.py:134-136.  jax's _shuffle = ceil(3 ln n / ln(2^32-1)) rounds of
// "stable sort by fresh random 32-bit keys"; a stable sort by key equals an ordinary sort of
// the 64-bit composite (key << 32 | position), which is what the batched bitonic network
// below sorts (all E epochs of an update in one grid; the key chain is data-independent).
#include "common.cuh"

namespace {

constexpr int SORT_TILE = 4096;       // u64 elements per shared-memory tile (32 KB)
constexpr int SORT_THREADS = 512;
constexpr int MAX_ROUNDS = 4;

__global__ void split_kernel(const uint32_t* __restrict__ key, uint32_t* __restrict__ out,
                             int num, int part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num) return;
    uint32_t o0, o1;
    threefry_split_at(key[0], key[1], (uint32_t)i, (uint32_t)num, part, o0, o1);
    out[2 * i] = o0; out[2 * i + 1] = o1;
}

__global__ void bits_kernel(const uint32_t* __restrict__ key, uint32_t* __restrict__ out,
                            long long n, int part) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = threefry_bits_at(key[0], key[1], (uint64_t)i, (uint64_t)n, part);
}

// Key chain of one update: per epoch (rnd, key) = split(key); inside permutation(rnd):
// per round (k, sub) = split(k).  subkeys: [E][MAX_ROUNDS][2].
__global__ void perm_keys_kernel(uint32_t* __restrict__ key, uint32_t* __restrict__ subkeys,
                                 int E, int rounds, int part) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t k0 = key[0], k1 = key[1];
    for (int e = 0; e < E; ++e) {
        uint32_t r0, r1, n0, n1;
        threefry_split_at(k0, k1, 0, 2, part, r0, r1);
        threefry_split_at(k0, k1, 1, 2, part, n0, n1);
        k0 = n0; k1 = n1;
        uint32_t c0 = r0, c1 = r1;
        for (int r = 0; r < rounds; ++r) {
            uint32_t a0, a1, s0, s1;
            threefry_split_at(c0, c1, 0, 2, part, a0, a1);
            threefry_split_at(c0, c1, 1, 2, part, s0, s1);
            c0 = a0; c1 = a1;
            subkeys[(e * MAX_ROUNDS + r) * 2 + 0] = s0;
            subkeys[(e * MAX_ROUNDS + r) * 2 + 1] = s1;
        }
    }
    key[0] = k0; key[1] = k1;
}

__global__ void perm_fill_kernel(const uint32_t* __restrict__ subkeys, int round,
                                 unsigned long long* __restrict__ comp, long long J,
                                 long long Jpad, int part) {
    const int e = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Jpad) return;
    unsigned long long v = ~0ull;
    if (i < J) {
        const uint32_t k0 = subkeys[(e * MAX_ROUNDS + round) * 2 + 0];
        const uint32_t k1 = subkeys[(e * MAX_ROUNDS + round) * 2 + 1];
        const uint32_t b = threefry_bits_at(k0, k1, (uint64_t)i, (uint64_t)J, part);
        v = ((unsigned long long)b << 32) | (unsigned long long)(uint32_t)i;
    }
    comp[(long long)e * Jpad + i] = v;
}

__device__ __forceinline__ void cmp_swap(unsigned long long& a, unsigned long long& b, bool asc) {
    if ((a > b) == asc) { const unsigned long long t = a; a = b; b = t; }
}

// Sort stages k = k_lo .. k_hi (doubling) restricted to strides j < tile, in shared memory.
// For k > tile only the j < tile tail of stage k is done here (k_lo == k_hi == k).
__global__ void __launch_bounds__(SORT_THREADS)
bitonic_tile_kernel(unsigned long long* __restrict__ data, long long Jpad, int tile,
                    long long k_lo, long long k_hi) {
    extern __shared__ unsigned long long sm[];
    unsigned long long* base = data + (long long)blockIdx.y * Jpad + (long long)blockIdx.x * tile;
    const long long g0 = (long long)blockIdx.x * tile;
    for (int i = threadIdx.x; i < tile; i += blockDim.x) sm[i] = base[i];
    __syncthreads();
    for (long long k = k_lo; k <= k_hi; k <<= 1) {
        long long j0 = k >> 1;
        if (j0 >= tile) j0 = tile >> 1;
        for (long long j = j0; j > 0; j >>= 1) {
            for (int p = threadIdx.x; p < tile / 2; p += blockDim.x) {
                // p-th pair: i has bit j clear
                const int i = (int)(((p & ~((int)j - 1)) << 1) | (p & ((int)j - 1)));
                const int l = i | (int)j;
                const bool asc = (((g0 + i) & k) == 0);
                cmp_swap(sm[i], sm[l], asc);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < tile; i += blockDim.x) base[i] = sm[i];
}

// One global compare-exchange step (stride j >= tile) of stage k.
__global__ void __launch_bounds__(256)
bitonic_global_kernel(unsigned long long* __restrict__ data, long long Jpad, long long k,
                      long long j) {
    unsigned long long* d = data + (long long)blockIdx.y * Jpad;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Jpad / 2) return;
    const long long i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
    const long long l = i | j;
    const bool asc = ((i & k) == 0);
    unsigned long long a = d[i], b = d[l];
    if ((a > b) == asc) { d[i] = b; d[l] = a; }
}

__global__ void perm_apply_kernel(const unsigned long long* __restrict__ comp,
                                  const int32_t* __restrict__ in, long long in_stride,
                                  int32_t* __restrict__ out, long long J, long long Jpad) {
    const int e = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J) return;
    const uint32_t pos = comp ? (uint32_t)comp[(long long)e * Jpad + i] : (uint32_t)i;
    out[(long long)e * J + i] = in ? in[(long long)e * in_stride + pos] : (int32_t)pos;
}

long long next_pow2(long long n) { long long p = 1; while (p < n) p <<= 1; return p; }

int shuffle_rounds(long long n) {
    if (n <= 1) return 0;
    return (int)ceil(3.0 * log((double)n) / log(4294967295.0));
}

// ascending sort of E independent rows of Jpad (a power of two) u64 keys: shared-memory tiles for
// strides < tile, one global compare-exchange pass per larger stride
int sort_rows_u64(cudaStream_t s, unsigned long long* comp, int E, long long Jpad) {
    const int tile = (int)(Jpad < SORT_TILE ? Jpad : SORT_TILE);
    if (tile < 2) return MLB_OK;
    const size_t smem = (size_t)tile * sizeof(unsigned long long);
    const dim3 tgrid((unsigned)(Jpad / tile), (unsigned)E);
    const int tthreads = tile / 2 < SORT_THREADS ? (tile / 2 < 32 ? 32 : tile / 2) : SORT_THREADS;
    bitonic_tile_kernel<<<tgrid, tthreads, smem, s>>>(comp, Jpad, tile, 2, tile);
    MLB_CHECK_LAUNCH();
    for (long long k = 2ll * tile; k <= Jpad; k <<= 1) {
        for (long long j = k >> 1; j >= tile; j >>= 1) {
            bitonic_global_kernel<<<dim3(mlb_cdiv(Jpad / 2, 256), E), 256, 0, s>>>(comp, Jpad, k, j);
            MLB_CHECK_LAUNCH();
        }
        bitonic_tile_kernel<<<tgrid, tthreads, smem, s>>>(comp, Jpad, tile, k, k);
        MLB_CHECK_LAUNCH();
    }
    return MLB_OK;
}

}  // namespace

MLB_API int mlb_threefry_split(void* stream, const uint32_t* key, uint32_t* out, int num,
                               int partitionable) {
    MLB_REQUIRE(key && out && num > 0);
    split_kernel<<<mlb_cdiv(num, 128), 128, 0, mlb_stream(stream)>>>(key, out, num, partitionable);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_threefry_bits(void* stream, const uint32_t* key, uint32_t* out, long long n,
                              int partitionable) {
    MLB_REQUIRE(key && out && n > 0);
    bits_kernel<<<mlb_cdiv(n, 256), 256, 0, mlb_stream(stream)>>>(key, out, n, partitionable);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API size_t mlb_ppo_permutations_workspace(int E, long long J) {
    const long long Jpad = next_pow2(J);
    return (size_t)E * MAX_ROUNDS * 2 * sizeof(uint32_t) + 256 +
           (size_t)E * Jpad * sizeof(unsigned long long) + (size_t)E * J * sizeof(int32_t);
}

static int permutations_impl(void* stream, uint32_t* key, const int32_t* values, int32_t* perm, int E,
                             long long J, int partitionable, void* ws, size_t ws_bytes) {
    MLB_REQUIRE(key && perm && E > 0 && J > 0 && J < (1ll << 31));
    if (!ws || ws_bytes < mlb_ppo_permutations_workspace(E, J)) return MLB_EWS;
    const int rounds = shuffle_rounds(J);
    MLB_REQUIRE(rounds <= MAX_ROUNDS);
    cudaStream_t s = mlb_stream(stream);
    const long long Jpad = next_pow2(J);
    uint8_t* w = reinterpret_cast<uint8_t*>(ws);
    uint32_t* subkeys = reinterpret_cast<uint32_t*>(w);
    w += ((size_t)E * MAX_ROUNDS * 2 * sizeof(uint32_t) + 255) / 256 * 256;
    unsigned long long* comp = reinterpret_cast<unsigned long long*>(w);
    w += (size_t)E * Jpad * sizeof(unsigned long long);
    int32_t* tmp = reinterpret_cast<int32_t*>(w);

    perm_keys_kernel<<<1, 32, 0, s>>>(key, subkeys, E, rounds, partitionable);
    MLB_CHECK_LAUNCH();
    if (rounds == 0) {
        perm_apply_kernel<<<dim3(mlb_cdiv(J, 256), E), 256, 0, s>>>(nullptr, values, 0, perm, J, Jpad);
        MLB_CHECK_LAUNCH();
        return MLB_OK;
    }
    const int tile = (int)(Jpad < SORT_TILE ? Jpad : SORT_TILE);
    // ping-pong so the last round lands in `perm`
    int32_t* bufs[2] = {perm, tmp};
    int cur = (rounds % 2 == 1) ? 0 : 1;   // buffer written by round 0
    const int32_t* prev = values;        // round 0 permutes `values` (NULL: arange(J)), shared by all epochs
    const long long in_stride = 0;
    for (int r = 0; r < rounds; ++r) {
        perm_fill_kernel<<<dim3(mlb_cdiv(Jpad, 256), E), 256, 0, s>>>(subkeys, r, comp, J, Jpad, partitionable);
        MLB_CHECK_LAUNCH();
        if (tile >= 2) {
            const int rc = sort_rows_u64(s, comp, E, Jpad);
            if (rc != MLB_OK) return rc;
        }
        perm_apply_kernel<<<dim3(mlb_cdiv(J, 256), E), 256, 0, s>>>(comp, prev, r == 0 ? in_stride : J, bufs[cur], J, Jpad);
        MLB_CHECK_LAUNCH();
        prev = bufs[cur];
        cur ^= 1;
    }
    return MLB_OK;
}

MLB_API int mlb_ppo_permutations(void* stream, uint32_t* key, int32_t* perm, int E, long long J,
                                 int partitionable, void* ws, size_t ws_bytes) {
    return permutations_impl(stream, key, nullptr, perm, E, J, partitionable, ws, ws_bytes);
}

MLB_API int mlb_ppo_permutations_of(void* stream, uint32_t* key, const int32_t* values, int32_t* perm, int E,
                                    long long J, int partitionable, void* ws, size_t ws_bytes) {
    MLB_REQUIRE(values);
    return permutations_impl(stream, key, values, perm, E, J, partitionable, ws, ws_bytes);
}

MLB_API long long mlb_sort_pad(long long n) { return next_pow2(n); }

MLB_API int mlb_sort_u64(void* stream, unsigned long long* keys, long long n_pad) {
    MLB_REQUIRE(keys && n_pad > 0 && (n_pad & (n_pad - 1)) == 0);
    return sort_rows_u64(mlb_stream(stream), keys, 1, n_pad);
}
