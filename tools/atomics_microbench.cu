// Microbenchmark that sizes the design: throughput of the access patterns the learn kernel can use
// for its gather and scatter on random rows of an HBM/L2-resident table.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atomics_microbench tools/atomics_microbench.cu
//   run  : tools/atomics_microbench  (prints one line per pattern: payload GB/s)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

enum Mode { LD4 = 0, RED4 = 1, ATOM4 = 2, RED1 = 3, ATOM1 = 4, LDRED4 = 5, ATOMRED4 = 6, BULKRED = 7, LD_ATOMRED4 = 8 };

// each warp-iteration touches one random row of row_floats floats (row_floats % 4 == 0), lanes stride over float4 chunks
template <int MODE>
__global__ void k_rows(float *tab, float *tab2, uint32_t n_rows_mask, uint32_t row_floats, uint32_t align_floats, uint32_t iters, float *sink)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    const uint32_t chunks = row_floats / 4;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t r = mix(warp + it * nwarps + 12345u) & n_rows_mask;
        size_t base = (size_t)r * align_floats;
        for (uint32_t c = lane; c < chunks; c += 32) {
            float4 *p = reinterpret_cast<float4 *>(tab + base) + c;
            float4 *p2 = reinterpret_cast<float4 *>(tab2 + base) + c;
            float4 v = make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f);
            if (MODE == LD4) { float4 x = __ldcg(p); acc += x.x + x.y + x.z + x.w; }
            if (MODE == RED4) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
            if (MODE == ATOM4) { float4 o = atomicAdd(p, v); acc += o.x; }
            if (MODE == RED1) { float *q = reinterpret_cast<float *>(p); for (int j = 0; j < 4; j++) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q + j), "f"(v.x) : "memory"); }
            if (MODE == ATOM1) { float *q = reinterpret_cast<float *>(p); for (int j = 0; j < 4; j++) acc += atomicAdd(q + j, v.x); }
            if (MODE == LDRED4) { float4 x = __ldcg(p); float4 y = __ldcg(p2); v.x = x.x * 1e-9f + y.x * 1e-9f;
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p2), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
            if (MODE == ATOMRED4) { float4 o = atomicAdd(p2, v); v.x = o.x * 1e-9f;
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
            if (MODE == LD_ATOMRED4) { float4 x = __ldcg(p); float4 o = atomicAdd(p2, v); v.x = o.x * 1e-9f + x.x * 1e-9f;
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
        }
    }
    if (acc == 123.456f) *sink = acc;
}

// TMA bulk reduce: one elected thread per warp issues cp.reduce.async.bulk of a whole row from smem
__global__ void k_bulkred(float *tab, uint32_t n_rows_mask, uint32_t row_floats, uint32_t align_floats, uint32_t iters)
{
    extern __shared__ __align__(128) float sm[];
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    float *mine = sm + (size_t)wib * row_floats;
    for (uint32_t c = lane; c < row_floats; c += 32) mine[c] = 1e-9f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const uint32_t bytes = row_floats * 4;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t r = mix(warp + it * nwarps + 12345u) & n_rows_mask;
        float *dst = tab + (size_t)r * align_floats;
        if (lane == 0) {
            uint32_t saddr = (uint32_t)__cvta_generic_to_shared(mine);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if ((it & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE> float run(float *tab, float *tab2, uint32_t mask, uint32_t row_floats, uint32_t align_floats, uint32_t iters, int blocks, float *sink)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_rows<MODE><<<blocks, 256>>>(tab, tab2, mask, row_floats, align_floats, iters / 4 + 1, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k_rows<MODE><<<blocks, 256>>>(tab, tab2, mask, row_floats, align_floats, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
    const int blocks = prop.multiProcessorCount * 8;
    const uint64_t nwarps = (uint64_t)blocks * 8;
    float *sink; CK(cudaMalloc(&sink, 4));
    struct Shape { const char *name; uint32_t row_floats, align_floats; } shapes[] = { {"c2 row 128B", 32, 32}, {"c3 row 1248B", 312, 320} };
    uint64_t table_bytes[] = { 8ull << 20, 128ull << 20, 2048ull << 20 };
    const char *mode_names[] = { "ld.cg.v4 (gather)", "red.v4", "atom.v4 (return)", "red.f32 x4", "atom.f32 x4", "ld w+acc, red w+acc (snapshot)", "atom acc -> red w", "TMA bulk reduce", "ld w, atom acc -> red w (k_learn)" };
    for (auto &sh : shapes) {
        for (uint64_t tb : table_bytes) {
            float *tab, *tab2;
            CK(cudaMalloc(&tab, tb + 4096)); CK(cudaMalloc(&tab2, tb + 4096));
            CK(cudaMemset(tab, 0, tb)); CK(cudaMemset(tab2, 0, tb));
            uint32_t n_rows = (uint32_t)(tb / (sh.align_floats * 4));
            uint32_t pow2 = 1; while (pow2 * 2 <= n_rows) pow2 *= 2;
            uint32_t mask = pow2 - 1;
            uint32_t iters = sh.row_floats == 32 ? 512 : 64;
            double rows = (double)nwarps * iters;
            double payload = rows * sh.row_floats * 4;
            printf("== %s, table %llu MB (x2 arrays), %u rows ==\n", sh.name, (unsigned long long)(tb >> 20), pow2);
            float ms;
            ms = run<LD4>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s  (%.3f ms, %.1f Mrows/s)\n", mode_names[0], payload / ms * 1e-6, ms, rows / ms * 1e-3);
            ms = run<RED4>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s  (%.3f ms, %.1f Mrows/s)\n", mode_names[1], payload / ms * 1e-6, ms, rows / ms * 1e-3);
            ms = run<ATOM4>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s  (%.3f ms, %.1f Mrows/s)\n", mode_names[2], payload / ms * 1e-6, ms, rows / ms * 1e-3);
            ms = run<RED1>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s  (%.3f ms, %.1f Mrows/s)\n", mode_names[3], payload / ms * 1e-6, ms, rows / ms * 1e-3);
            ms = run<ATOM1>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s  (%.3f ms, %.1f Mrows/s)\n", mode_names[4], payload / ms * 1e-6, ms, rows / ms * 1e-3);
            ms = run<LDRED4>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s algorithmic(16B/slot)  (%.3f ms, %.1f Mrows/s)\n", mode_names[5], 4 * payload / ms * 1e-6, ms, rows / ms * 1e-3);
            ms = run<ATOMRED4>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s algorithmic(12B/slot)  (%.3f ms, %.1f Mrows/s)\n", mode_names[6], 3 * payload / ms * 1e-6, ms, rows / ms * 1e-3);
            ms = run<LD_ATOMRED4>(tab, tab2, mask, sh.row_floats, sh.align_floats, iters, blocks, sink); printf("  %-36s %8.1f GB/s algorithmic(16B/slot)  (%.3f ms, %.1f Mrows/s)\n", mode_names[8], 4 * payload / ms * 1e-6, ms, rows / ms * 1e-3);
            {
                size_t smem = (size_t)8 * sh.row_floats * 4;
                cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
                k_bulkred<<<blocks, 256, smem>>>(tab, mask, sh.row_floats, sh.align_floats, iters / 4 + 1);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a));
                k_bulkred<<<blocks, 256, smem>>>(tab, mask, sh.row_floats, sh.align_floats, iters);
                CK(cudaEventRecord(b));
                CK(cudaDeviceSynchronize());
                CK(cudaEventElapsedTime(&ms, a, b));
                printf("  %-36s %8.1f GB/s  (%.3f ms, %.1f Mrows/s)\n", mode_names[7], payload / ms * 1e-6, ms, rows / ms * 1e-3);
            }
            CK(cudaFree(tab)); CK(cudaFree(tab2));
        }
    }
    return 0;
}
