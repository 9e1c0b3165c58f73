// Same-address contention of the learn kernel's update pattern: rows drawn uniformly vs log-uniformly
// (Zipf ~ 1, as the synthetic CTR stream does) from n_rows rows of 128 B.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hotrow_microbench tools/hotrow_microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
enum { LD = 0, ATOM = 1, RED = 2, LEARN = 3, LEARN_NORET = 4 };
template <int MODE>
__global__ void k(float *w, float *acc, uint32_t n_rows, int zipf, uint32_t iters, float *sink)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const float logn = logf((float)n_rows + 1.0f);
    float s = 0.f;
    for (uint32_t it = 0; it < iters; it++) {
        // a warp handles 4 rows per iteration (8 lanes x float4 = 128 B each), like 4 of an example's 8 rows
        uint32_t h = mix(warp * 4 + (lane >> 3) + it * nwarps * 4 + 99u);
        uint32_t r;
        if (zipf) { float u = (float)(h >> 8) * (1.0f / 16777216.0f); r = (uint32_t)(__expf(u * logn) - 1.0f); if (r >= n_rows) r = n_rows - 1; }
        else r = h % n_rows;
        size_t off = (size_t)r * 32 + (lane & 7) * 4;
        float4 *pw = reinterpret_cast<float4 *>(w + off), *pa = reinterpret_cast<float4 *>(acc + off);
        float4 v = make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f);
        if (MODE == LD) { float4 x = __ldcg(pw); s += x.x; }
        if (MODE == ATOM) { float4 o = atomicAdd(pa, v); s += o.x; }
        if (MODE == RED) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(pw), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
        if (MODE == LEARN) { float4 x = __ldcg(pw); float4 o = atomicAdd(pa, v); v.x = (o.x + x.x) * 1e-9f;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(pw), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
        if (MODE == LEARN_NORET) { float4 x = __ldcg(pw); float4 a = __ldcg(pa); v.x = (a.x + x.x) * 1e-9f;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(pa), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(pw), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
    }
    if (s == 123.456f) *sink = s;
}
template <int MODE> void run(const char *name, float *w, float *acc, uint32_t n_rows, int zipf, int blocks, float *sink)
{
    const uint32_t iters = 256;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k<MODE><<<blocks, 256>>>(w, acc, n_rows, zipf, iters / 4, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k<MODE><<<blocks, 256>>>(w, acc, n_rows, zipf, iters, sink);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    double rows = (double)blocks * 8 * 4 * iters;
    printf("  %-34s %s  %9.1f Mrows/s  (= %7.1f M examples/s at 8 rows/example)\n", name, zipf ? "zipf   " : "uniform", rows / ms * 1e-3, rows / 8 / ms * 1e-3);
}
int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    float *sink; CK(cudaMalloc(&sink, 4));
    for (uint32_t n_rows : {100000u, 800000u}) {
        float *w, *acc; size_t bytes = (size_t)n_rows * 128;
        CK(cudaMalloc(&w, bytes)); CK(cudaMalloc(&acc, bytes)); CK(cudaMemset(w, 0, bytes)); CK(cudaMemset(acc, 0, bytes));
        for (int occ : {4, 8}) {
            int blocks = prop.multiProcessorCount * occ;
            printf("== %u rows of 128 B, %d blocks/SM x 256 threads ==\n", n_rows, occ);
            for (int z = 0; z < 2; z++) {
                run<LD>("ld.cg.v4", w, acc, n_rows, z, blocks, sink);
                run<ATOM>("atom.v4 (return)", w, acc, n_rows, z, blocks, sink);
                run<RED>("red.v4", w, acc, n_rows, z, blocks, sink);
                run<LEARN>("ld w, atom acc -> red w", w, acc, n_rows, z, blocks, sink);
                run<LEARN_NORET>("ld w, ld acc, red acc, red w", w, acc, n_rows, z, blocks, sink);
            }
        }
        CK(cudaFree(w)); CK(cudaFree(acc));
    }
    return 0;
}
