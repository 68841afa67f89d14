// Probe of the tcgen05.ld/st .16x256b register <-> (lane, column) mapping assumed by the fused layer
// epilogues: write a [128 lanes x 64 cols] pattern with .32x32b (lane = row, register = column), read
// it back with .16x256b.x8 and check   reg[4k + 2h + e] == (row = 16*half + lane/4 + 8h, col = 8k + 2(lane%4) + e).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/tmem_probe tools/probe/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t* out, int* bad) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // write: lane = row (warp*32 + lane), reg j = col j  -> value row*1000 + col
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) v[j] = (warp * 32 + lane) * 1000 + c0 + j;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(base + lane_base + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                       "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncwarp();
    for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        const uint32_t ta = base + lane_base + ((uint32_t)(half * 16) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 8; ++k)
            for (int h = 0; h < 2; ++h)
                for (int e = 0; e < 2; ++e) {
                    const int row = warp * 32 + half * 16 + lane / 4 + 8 * h, col = 8 * k + 2 * (lane % 4) + e;
                    const uint32_t got = r[4 * k + 2 * h + e];
                    if (got != (uint32_t)(row * 1000 + col)) atomicAdd(bad, 1);
                    if (warp == 0 && half == 0) out[lane * 32 + 4 * k + 2 * h + e] = got;
                }
        // round trip: store the same registers back with the same shape, shifted by +1, and re-check through 32x32b
        for (int j = 0; j < 32; ++j) r[j] += 1;
        asm volatile(
            "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
              "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
              "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
              "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncwarp();
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(base + lane_base + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j)
            if (v[j] != (uint32_t)((warp * 32 + lane) * 1000 + c0 + j + 1)) atomicAdd(bad + 1, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
    }
}

int main() {
    uint32_t* out; int* bad;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    probe<<<1, 128>>>(out, bad);
    cudaError_t e = cudaDeviceSynchronize();
    int hb[2]; uint32_t ho[1024];
    cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 4096, cudaMemcpyDeviceToHost);
    printf("TMEM_PROBE err=%d mismatches_16x256b=%d roundtrip_mismatches=%d\n", (int)e, hb[0], hb[1]);
    printf("lane0 regs:"); for (int j = 0; j < 8; ++j) printf(" %u", ho[j]); printf("\nlane1 regs:");
    for (int j = 0; j < 8; ++j) printf(" %u", ho[32 + j]); printf("\nlane4 regs:");
    for (int j = 0; j < 8; ++j) printf(" %u", ho[128 + j]); printf("\n");
    return (e == cudaSuccess && hb[0] == 0 && hb[1] == 0) ? 0 : 1;
}
