// Stand-alone check of the tcgen05 3xTF32 GEMM tiles (csrc/fwgpu_umma.cuh) against a double-precision CPU product.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_gemm_test tools/umma_gemm_test.cu
//   run  : timeout 120 tools/umma_gemm_test
#include "../fwumious_wabbit_b200/csrc/fwgpu_umma.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace fwgpu;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

static uint32_t rng_state = 12345u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.0f; }

template <bool A_T, bool B_T, int EPI> static void launch(const HeadGemmParams &p, uint32_t splits)
{
    auto kern = k_umma_gemm<A_T, B_T, EPI>;
    static bool configured = false;
    if (!configured) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UMMA_SMEM_BYTES)); configured = true; }
    dim3 grid((p.N + (EPI == HEAD_EPI_SUMS ? 1 : 0) + UMMA_BN - 1) / UMMA_BN, (p.M + UMMA_BM - 1) / UMMA_BM, splits);
    kern<<<grid, UMMA_THREADS, UMMA_SMEM_BYTES>>>(p);
}

template <bool A_T, bool B_T> static int run_case(uint32_t M, uint32_t N, uint32_t K, bool sums, bool timeit)
{
    const uint32_t lda = A_T ? M : K, ldb = B_T ? N : K;
    std::vector<float> A((size_t)M * K), B((size_t)N * K), C((size_t)M * N), C2((size_t)M * N);
    for (auto &x : A) x = frand();
    for (auto &x : B) x = frand() * 0.1f;
    std::vector<float> bias(N), G1b(M), G2b(M);
    for (auto &x : bias) x = frand();
    float *dA, *dB, *dC, *dC2, *dbias, *dG1b, *dG2b;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMalloc(&dC2, C.size() * 4));
    CK(cudaMalloc(&dbias, N * 4)); CK(cudaMalloc(&dG1b, M * 4)); CK(cudaMalloc(&dG2b, M * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, C.size() * 4)); CK(cudaMemset(dC2, 0, C.size() * 4)); CK(cudaMemset(dG1b, 0, M * 4)); CK(cudaMemset(dG2b, 0, M * 4));
    HeadGemmParams p{};
    p.A = dA; p.lda = lda; p.B = dB; p.ldb = ldb; p.C = dC; p.ldc = N; p.M = M; p.N = N; p.K = K; p.G1 = dC; p.G2 = dC2; p.G1_bias = dG1b; p.G2_bias = dG2b;
    p.bias = dbias; p.relu = 0;
    uint32_t splits = 1;
    if (sums) { splits = 10; p.k_split = ((K + splits - 1) / splits + 15) / 16 * 16; splits = (K + p.k_split - 1) / p.k_split; }
    if (sums) launch<A_T, B_T, HEAD_EPI_SUMS>(p, splits); else launch<A_T, B_T, HEAD_EPI_BIAS_ACT>(p, 1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(C2.data(), dC2, C.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(G1b.data(), dG1b, M * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(G2b.data(), dG2b, M * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0, max_err2 = 0, max_ref2 = 0;
    const uint32_t step_m = M > 512 ? 37 : 1, step_n = N > 512 ? 11 : 1;
    for (uint32_t m = 0; m < M; m += step_m) {
        for (uint32_t n = 0; n < N; n += step_n) {
            double s = sums ? 0.0 : (double)bias[n], s2 = 0;
            for (uint32_t k = 0; k < K; k++) {
                const double a = A_T ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k], b = B_T ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k];
                s += a * b; s2 += (double)((float)a * (float)a) * (double)((float)b * (float)b);
            }
            max_err = fmax(max_err, fabs(s - C[(size_t)m * N + n])); max_ref = fmax(max_ref, fabs(s));
            max_err2 = fmax(max_err2, fabs(s2 - C2[(size_t)m * N + n])); max_ref2 = fmax(max_ref2, fabs(s2));
        }
        if (sums) { // bias column: sums of A(m, .) and of its squares
            double s = 0, s2 = 0;
            for (uint32_t k = 0; k < K; k++) { const double a = A_T ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]; s += a; s2 += (double)((float)a * (float)a); }
            max_err = fmax(max_err, fabs(s - G1b[m])); max_ref = fmax(max_ref, fabs(s)); max_err2 = fmax(max_err2, fabs(s2 - G2b[m])); max_ref2 = fmax(max_ref2, fabs(s2));
        }
    }
    const bool ok = max_err <= 2e-5 * fmax(max_ref, 1.0) && (!sums || max_err2 <= 2e-5 * fmax(max_ref2, 1.0)); // 3xTF32 + tensor-core fp32 accumulation
    printf("%s A_T=%d B_T=%d %s M=%u N=%u K=%u  max|err| %.3e (max|c| %.3e)", ok ? "ok  " : "FAIL", (int)A_T, (int)B_T, sums ? "SUMS " : "BIAS ", M, N, K, max_err, max_ref);
    if (sums) printf("  squares: max|err| %.3e (max %.3e)", max_err2, max_ref2);
    if (timeit) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int reps = 50;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) { if (sums) launch<A_T, B_T, HEAD_EPI_SUMS>(p, splits); else launch<A_T, B_T, HEAD_EPI_BIAS_ACT>(p, 1); }
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  %.1f us/launch, %.1f TFLOP/s (fp32-equivalent)", ms * 1e3 / reps, 2.0 * M * N * K * (sums ? 2 : 1) / (ms * 1e-3 / reps) * 1e-12);
    }
#ifdef UMMA_DEBUG_TIMING
    if (timeit) {
        unsigned long long h[16];
        CK(cudaMemcpyFromSymbol(h, umma_dbg, sizeof(h)));
        printf("\n      block 0, thread 0 (ns): alloc %llu, k-loop %llu, wait for the last MMAs %llu, epilogue TMEM->smem %llu, smem->global %llu", h[1] - h[0], h[3] - h[1], h[4] - h[3], h[5] - h[4], h[6] - h[5]);
    }
#endif
    printf("\n");
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dC2); cudaFree(dbias); cudaFree(dG1b); cudaFree(dG2b);
    return ok ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += run_case<false, false>(128, 128, 32, false, false);   // one tile, one k-block
    bad += run_case<false, false>(128, 128, 256, false, false);
    bad += run_case<false, false>(300, 256, 820, false, false);  // forward layer shape, ragged M
    bad += run_case<false, true>(300, 820, 256, false, false);   // backward: dX = dZ W
    bad += run_case<true, true>(256, 820, 1000, true, false);    // update: G1 / G2 over the sub-batch, split K
    bad += run_case<false, false>(4096, 256, 820, false, true);
    bad += run_case<false, true>(4096, 820, 256, false, true);
    bad += run_case<true, true>(256, 820, 4096, true, true);
    bad += run_case<false, false>(4096, 256, 32, false, true);
    bad += run_case<false, false>(4096, 256, 256, false, true);
    bad += run_case<false, false>(4096, 256, 1664, false, true);
    bad += run_case<false, false>(4096, 128, 820, false, true);
    bad += run_case<false, false>(128, 128, 820, false, true);
    bad += run_case<false, false>(16384, 256, 820, false, true);
    bad += run_case<false, false>(37888, 256, 832, false, true);  // 296 row tiles x 2 = 4 full waves
    printf(bad ? "FAILED (%d cases)\n" : "all cases ok\n", bad);
    return bad ? 1 : 0;
}
