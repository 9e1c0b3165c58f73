// Speed of light of k_learn_fixed's memory skeleton on c2: what the memory system sustains when a warp does NOTHING but the
// kernel's global-memory operations for two records per round -- 16-byte-aligned 128-byte rows drawn like the bench's ids
// (log-uniform ranks over 1e5 ids per field = csrc/host/synth.cpp zipf_id, or uniform), 7 of 8 chunks per row:
//   ld.cg.v4 w  ->  atom.add.v4 acc (return used)  ->  red.add.v4 w          per chunk, plus the 44-byte record read and the
//   4-byte prediction write.  No translate, no dot products, no sigmoid, no LUT, no LR cells.
// The table geometry is c2's: 2^20 floats + tail per array, rows at (hash & ~3) so that a row usually straddles two lines.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/skeleton_microbench tools/skeleton_microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __host__ inline uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// recs[n][8]: row base (float index, multiple of 4) of each of the 8 fields
template <bool UPDATE, bool RETURN>
__global__ void __launch_bounds__(512, 2) k_skeleton(float *w, float *acc, const uint32_t *recs, uint32_t n, float *preds)
{
    const uint32_t lane = threadIdx.x & 31, sl = lane & 15, sg = lane >> 4;
    const uint32_t n_groups = gridDim.x * 32, g = (blockIdx.x * 16 + (threadIdx.x >> 5)) * 2 + sg;
    for (uint32_t ex = g; ex < n; ex += n_groups) {
        float s = 0.f;
        float4 v[4]; uint32_t at[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {                  // chunk j = sl + 16 t of 64: row e = j / 8, chunk c = j % 8; own-field chunk skipped
            const uint32_t j = sl + 16 * t, e = j >> 3, c = j & 7;
            at[t] = __ldg(recs + (size_t)ex * 8 + e) + 4 * c;
            v[t] = (e != c) ? __ldcg(reinterpret_cast<const float4 *>(w + at[t])) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < 4; t++) s += v[t].x;
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (sl == 0) preds[ex] = s;
        if (UPDATE) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint32_t j = sl + 16 * t, e = j >> 3, c = j & 7;
                if (e == c) continue;
                float4 g4 = make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f), u = g4;
                if (RETURN) { const float4 o = atomicAdd(reinterpret_cast<float4 *>(acc + at[t]), g4); u.x = o.x * 1e-9f; }
                else asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(acc + at[t]), "f"(g4.x), "f"(g4.y), "f"(g4.z), "f"(g4.w) : "memory");
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(w + at[t]), "f"(u.x), "f"(u.y), "f"(u.z), "f"(u.w) : "memory");
            }
        }
    }
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const uint32_t n = 4000000, L = (1u << 20) + 32;
    float *w, *acc, *preds; uint32_t *recs;
    CK(cudaMalloc(&w, L * 4)); CK(cudaMalloc(&acc, L * 4)); CK(cudaMalloc(&preds, n * 4)); CK(cudaMalloc(&recs, (size_t)n * 32));
    CK(cudaMemset(w, 0, L * 4)); CK(cudaMemset(acc, 0, L * 4));
    std::vector<uint32_t> h((size_t)n * 8);
    const uint32_t V = 100000;
    for (int zipf = 1; zipf >= 0; zipf--) {
        uint32_t st = 12345;
        for (size_t i = 0; i < (size_t)n * 8; i++) {
            st = st * 1664525u + 1013904223u;
            uint32_t id;
            if (zipf) { const double u = (st >> 8) * (1.0 / 16777216.0); id = (uint32_t)(exp(u * log(V + 1.0)) - 1.0); if (id >= V) id = V - 1; } // P(rank r) ~ 1/(r+1)
            else id = mix(st) % V;
            h[i] = mix(id * 8 + (uint32_t)(i & 7) + 77u) & ((1u << 20) - 1) & ~3u; // the field's hash of that id, masked like feature_buffer.rs:142-148 (k = 4)
        }
        CK(cudaMemcpy(recs, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        const int blocks = prop.multiProcessorCount * 2;
        auto time = [&](auto kern, const char *name) {
            cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
            kern<<<blocks, 512>>>(w, acc, recs, n / 4, preds); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a)); kern<<<blocks, 512>>>(w, acc, recs, n, preds); CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            printf("  %-11s %-46s %8.1f M records/s\n", zipf ? "log-uniform" : "uniform", name, n / ms * 1e-3);
        };
        time(k_skeleton<false, false>, "gather only (predict)");
        time(k_skeleton<true, true>, "gather, atom acc (return) -> red w  [the kernel]");
        time(k_skeleton<true, false>, "gather, red acc, red w");
    }
    return 0;
}
