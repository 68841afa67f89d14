// Phase timeline of the persistent fused-layer kernels (second-generation epilogues): includes the
// product kernels with MLB_PHASE_PROFILE so that two epilogue warps (first and last) stamp %globaltimer
// at: 0 tile start, 1 accumulator ready, 2/6 xhat panel of chunk 0/1 ready (backward), 3 pass 1 done,
// 4 row totals exchanged, 5 pass 2 + stores done.  Prints mean phase durations per tile slot.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/phase_profile \
//        tools/probe/phase_profile.cu -lcuda
#define MLB_PHASE_PROFILE 1
#include "../../madrona-learn_b200/csrc/mlp_tc_persist.cu"
#include <cstdio>
#include <vector>

static void report(const char* name, const std::vector<unsigned long long>& h, int ctas, int tiles, bool bwd) {
    printf("== %s (ns, mean over %d CTAs; warp A = first epilogue warp, warp B = last)\n", name, ctas);
    for (int w = 0; w < 2; ++w)
        for (int it = 0; it < tiles; ++it) {
            double d[8] = {0}; int n = 0;
            double t0m = 0;
            for (int b = 0; b < ctas; ++b) {
                const unsigned long long* r = &h[(((size_t)b * 8 + it) * 2 + w) * 8];
                if (!r[0] || !r[5]) continue;
                ++n;
                d[0] += (double)(r[1] - r[0]);                       // wait for the accumulator
                d[1] += bwd ? (double)(r[2] - r[1]) : 0.0;           // wait for xhat panel (chunk 0)
                d[2] += (double)(r[3] - (bwd ? r[2] : r[1]));        // pass 1 (incl. second panel wait)
                d[3] += (double)(r[4] - r[3]);                       // row totals exchange
                d[4] += (double)(r[5] - r[4]);                       // pass 2 + stores
                d[5] += (double)(r[5] - r[0]);
                d[6] += bwd && r[6] ? (double)(r[6] - r[2]) : 0.0;   // chunk-1 panel ready, relative to chunk 0
                t0m += (double)(r[0] - h[(((size_t)b * 8 + 0) * 2 + w) * 8]);
            }
            if (!n) continue;
            printf("warp %c tile %d: start@%7.0f accwait %6.0f xhwait %6.0f pass1 %6.0f exch %6.0f pass2 %6.0f total %6.0f (panel1 ready +%5.0f)  n=%d\n",
                   w ? 'B' : 'A', it, t0m / n, d[0] / n, d[1] / n, d[2] / n, d[3] / n, d[4] / n, d[5] / n, d[6] / n, n);
        }
}

int main() {
    const int M = 65536, K = 256, HN = 256;
    __nv_bfloat16 *X, *W, *XH, *Y, *DZ;
    float *scale, *bias, *rstd, *ds, *db;
    cudaMalloc(&X, (size_t)M * K * 2); cudaMalloc(&W, (size_t)K * HN * 2); cudaMalloc(&XH, (size_t)M * HN * 2);
    cudaMalloc(&Y, (size_t)M * HN * 2); cudaMalloc(&DZ, (size_t)M * HN * 2);
    cudaMalloc(&scale, HN * 4); cudaMalloc(&bias, HN * 4); cudaMalloc(&rstd, (size_t)M * 4);
    cudaMalloc(&ds, HN * 4); cudaMalloc(&db, HN * 4);
    cudaMemset(X, 0x11, (size_t)M * K * 2); cudaMemset(W, 0x22, (size_t)K * HN * 2); cudaMemset(XH, 0x33, (size_t)M * HN * 2);
    cudaMemset(scale, 0, HN * 4); cudaMemset(bias, 0, HN * 4); cudaMemset(rstd, 0, (size_t)M * 4);
    cudaMemset(ds, 0, HN * 4); cudaMemset(db, 0, HN * 4);
    const int ctas = 148;
    const size_t n = (size_t)ctas * 8 * 2 * 8;
    unsigned long long* dprof;
    cudaMalloc(&dprof, n * 8);
    cudaMemcpyToSymbol(g_prof, &dprof, sizeof(dprof));
    std::vector<unsigned long long> h(n);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(dprof, 0, n * 8);
        int rc = tcp::launch_fwd_persist(0, X, W, scale, bias, Y, DZ, rstd, M, K, HN, K, K);
        cudaError_t e = cudaDeviceSynchronize();
        if (rc || e) { printf("fwd failed rc=%d err=%d\n", rc, (int)e); return 1; }
    }
    cudaMemcpy(h.data(), dprof, n * 8, cudaMemcpyDeviceToHost);
    report("fwd_persist2", h, ctas, 4, false);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(dprof, 0, n * 8);
        int rc = tcp::launch_dx_persist(0, X, W, scale, bias, XH, rstd, DZ, ds, db, M, K, HN, K, HN);
        cudaError_t e = cudaDeviceSynchronize();
        if (rc || e) { printf("dx failed rc=%d err=%d\n", rc, (int)e); return 1; }
    }
    cudaMemcpy(h.data(), dprof, n * 8, cudaMemcpyDeviceToHost);
    report("dx_persist2", h, ctas, 4, true);
    // decomposition of the backward kernel: device time with parts of the epilogue switched off
    unsigned long long* nullp = nullptr;
    cudaMemcpyToSymbol(g_prof, &nullp, sizeof(nullp));
    const int modes[] = {0, 8, 4 | 1 | 2, 4 | 1, 4, 1, 2, 1 | 2};
    const char* names[] = {"full", "null epilogue (TMA + MMA only)", "TMEM passes only (no xhat, no math, no stores)",
                           "no xhat, no pass-1 math", "no xhat panels", "no pass-1 math", "no stores", "no pass-1 math, no stores"};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mi = 0; mi < 8; ++mi) {
        cudaMemcpyToSymbol(g_dbg, &modes[mi], sizeof(int));
        for (int w = 0; w < 3; ++w) tcp::launch_dx_persist(0, X, W, scale, bias, XH, rstd, DZ, ds, db, M, K, HN, K, HN);
        cudaEventRecord(e0);
        for (int w = 0; w < 10; ++w) tcp::launch_dx_persist(0, X, W, scale, bias, XH, rstd, DZ, ds, db, M, K, HN, K, HN);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("dx2 dbg=%2d %-48s %7.2f us/launch (warm L2, 10 back-to-back)  err=%d\n", modes[mi], names[mi], ms * 100.f, (int)e);
    }
    return 0;
}
