// Sizes the sharded-table design: what one B200 sustains over NVLink 5 / NVSwitch against a PEER's HBM for
// the row-granular access pattern of the wide FFM kernel -- 39 random rows of 1248 B per record -- with
//   bulk loads   (cp.async.bulk global -> shared, mbarrier complete_tx)            "pull the rows"
//   bulk reduces (cp.reduce.async.bulk shared -> global .add.f32)                  "push the gradients into the owner's L2"
//   bulk stores  (cp.async.bulk shared -> global)                                  "push the gradients into a staging ring"
//   16-byte ld.v4 / red.v4 / atom.v4                                               what round 1's sharded kernel issued
// against the local table (same GPU) and against the peer's, one direction and both directions at once.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/nvlink_bulk_microbench tools/nvlink_bulk_microbench.cu
//   run  : tools/nvlink_bulk_microbench   (2 GPUs with peer access: the NVLink table; 1 GPU: the local skeleton of the wide kernel)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

constexpr int ROWS = 39, ROW_FLOATS = 312, ROW_BYTES = ROW_FLOATS * 4, ALIGN_FLOATS = 320;
enum Mode { LOAD_BULK = 0, RED_BULK = 1, STORE_BULK = 2, LD16 = 3, RED16 = 4, ATOM16 = 5, RECORD_BULK = 6 };

__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem, const void *g, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g), "r"(bytes),
                 "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void bulk_red(void *g, const void *smem, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g), "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store(void *g, const void *smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(bytes) : "memory");
}

// one "record" per CTA and iteration: ROWS random rows of the table (and of tab2 for RECORD_BULK)
template <int MODE>
__global__ void __launch_bounds__(256) k_rows(float *tab, float *tab2, uint32_t row_mask, uint32_t iters, float *sink)
{
    extern __shared__ __align__(128) float sm[];
    float *W = sm, *A = sm + ROWS * ROW_FLOATS;
    __shared__ uint64_t bar;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) mbar_init(&bar, ROWS);
    for (uint32_t i = tid; i < 2 * ROWS * ROW_FLOATS; i += 256) sm[i] = 1e-9f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    float acc = 0.f;
    for (uint32_t it = 0; it < iters; it++) {
        const uint32_t rec = blockIdx.x + it * gridDim.x;
        if (MODE == LOAD_BULK || MODE == RECORD_BULK) {
            if (tid < ROWS) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
            if (tid < ROWS) {
                const size_t base = (size_t)(mix(rec * ROWS + tid + 777u) & row_mask) * ALIGN_FLOATS;
                mbar_expect(&bar, MODE == RECORD_BULK ? 2 * ROW_BYTES : ROW_BYTES);
                bulk_load(W + tid * ROW_FLOATS, tab + base, ROW_BYTES, &bar);
                if (MODE == RECORD_BULK) bulk_load(A + tid * ROW_FLOATS, tab2 + base, ROW_BYTES, &bar);
            }
            mbar_wait(&bar, it & 1);
            if (MODE == RECORD_BULK) {
                __syncthreads();
                if (tid < ROWS) {
                    const size_t base = (size_t)(mix(rec * ROWS + tid + 777u) & row_mask) * ALIGN_FLOATS;
                    bulk_red(tab + base, W + tid * ROW_FLOATS, ROW_BYTES);
                    bulk_red(tab2 + base, A + tid * ROW_FLOATS, ROW_BYTES);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else if (MODE == RED_BULK || MODE == STORE_BULK) {
            if (tid < ROWS) {
                const size_t base = (size_t)(mix(rec * ROWS + tid + 777u) & row_mask) * ALIGN_FLOATS;
                if (MODE == RED_BULK) bulk_red(tab + base, W + tid * ROW_FLOATS, ROW_BYTES);
                else bulk_store(tab + base, W + tid * ROW_FLOATS, ROW_BYTES);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
        } else {
            for (uint32_t idx = tid; idx < ROWS * (ROW_FLOATS / 4); idx += 256) {
                const uint32_t e = idx / (ROW_FLOATS / 4), c = idx - e * (ROW_FLOATS / 4);
                const size_t base = (size_t)(mix(rec * ROWS + e + 777u) & row_mask) * ALIGN_FLOATS;
                float4 *p = reinterpret_cast<float4 *>(tab + base) + c;
                if (MODE == LD16) { float4 x = __ldcg(p); acc += x.x; }
                if (MODE == RED16) asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1e-9f), "f"(1e-9f), "f"(1e-9f), "f"(1e-9f) : "memory");
                if (MODE == ATOM16) { float4 o = atomicAdd(p, make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f)); acc += o.x; }
            }
        }
    }
    if (tid < ROWS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123.456f) *sink = acc;
}

struct Dev { int id; float *tab, *tab2, *sink; cudaStream_t s; cudaEvent_t a, b; };

template <int MODE> static void launch(Dev &d, float *tab, float *tab2, uint32_t mask, uint32_t iters, int blocks)
{
    CK(cudaSetDevice(d.id));
    const size_t smem = (size_t)2 * ROWS * ROW_BYTES;
    CK(cudaFuncSetAttribute(k_rows<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rows<MODE><<<blocks, 256, smem, d.s>>>(tab, tab2, mask, iters, d.sink);
}

// runs MODE on dev a against `ta` (and, when both, on dev b against `tb` at the same time); returns ms of the slower one
template <int MODE> static float timed(Dev &a, float *ta, float *ta2, Dev *b, float *tb, float *tb2, uint32_t mask, uint32_t iters, int blocks)
{
    launch<MODE>(a, ta, ta2, mask, iters / 4 + 1, blocks);
    if (b) launch<MODE>(*b, tb, tb2, mask, iters / 4 + 1, blocks);
    CK(cudaSetDevice(a.id)); CK(cudaStreamSynchronize(a.s));
    if (b) { CK(cudaSetDevice(b->id)); CK(cudaStreamSynchronize(b->s)); }
    CK(cudaSetDevice(a.id)); CK(cudaEventRecord(a.a, a.s));
    if (b) { CK(cudaSetDevice(b->id)); CK(cudaEventRecord(b->a, b->s)); }
    launch<MODE>(a, ta, ta2, mask, iters, blocks);
    if (b) launch<MODE>(*b, tb, tb2, mask, iters, blocks);
    CK(cudaSetDevice(a.id)); CK(cudaEventRecord(a.b, a.s)); CK(cudaStreamSynchronize(a.s));
    float ms = 0, ms2 = 0;
    CK(cudaEventElapsedTime(&ms, a.a, a.b));
    if (b) { CK(cudaSetDevice(b->id)); CK(cudaEventRecord(b->b, b->s)); CK(cudaStreamSynchronize(b->s)); CK(cudaEventElapsedTime(&ms2, b->a, b->b)); }
    return ms > ms2 ? ms : ms2;
}

int main()
{
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) {
        // one GPU: the local skeleton of the wide kernel's copy-engine traffic for an L2-sized and an HBM-sized table
        // (what bounds k_learn_rows: bytes or the rate of row-sized bulk operations?)
        Dev d0; d0.id = 0;
        CK(cudaSetDevice(0));
        CK(cudaMalloc(&d0.sink, 4)); CK(cudaStreamCreate(&d0.s)); CK(cudaEventCreate(&d0.a)); CK(cudaEventCreate(&d0.b));
        cudaDeviceProp prop0; CK(cudaGetDeviceProperties(&prop0, 0));
        for (uint64_t tb : {64ull << 20, 1024ull << 20}) {
            CK(cudaMalloc(&d0.tab, tb + 4096)); CK(cudaMalloc(&d0.tab2, tb + 4096));
            CK(cudaMemset(d0.tab, 0, tb)); CK(cudaMemset(d0.tab2, 0, tb));
            uint32_t nr = (uint32_t)(tb / (ALIGN_FLOATS * 4)), p2 = 1; while (p2 * 2 <= nr) p2 *= 2;
            for (int per_sm : {2, 4}) {
                const int blocks0 = prop0.multiProcessorCount * per_sm; const uint32_t it0 = 400;
                const double recs0 = (double)blocks0 * it0;
                float l = timed<LOAD_BULK>(d0, d0.tab, d0.tab2, nullptr, nullptr, nullptr, p2 - 1, it0, blocks0);
                float r = timed<RED_BULK>(d0, d0.tab, d0.tab2, nullptr, nullptr, nullptr, p2 - 1, it0, blocks0);
                float q = timed<RECORD_BULK>(d0, d0.tab, d0.tab2, nullptr, nullptr, nullptr, p2 - 1, it0, blocks0);
                printf("local, table 2 x %4llu MB, %d CTAs/SM: 39 bulk loads %6.2f M records/s | 39 bulk reductions %6.2f M records/s | record (78 loads + 78 reductions) %6.2f M records/s\n",
                       (unsigned long long)(tb >> 20), per_sm, recs0 / l * 1e-3, recs0 / r * 1e-3, recs0 / q * 1e-3);
            }
            CK(cudaFree(d0.tab)); CK(cudaFree(d0.tab2));
        }
        return 0;
    }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (!can) { printf("no peer access between GPU 0 and 1\n"); return 0; }
    const uint64_t table_bytes = 1024ull << 20; // per array
    const uint32_t n_rows = (uint32_t)(table_bytes / (ALIGN_FLOATS * 4));
    uint32_t pow2 = 1; while (pow2 * 2 <= n_rows) pow2 *= 2;
    const uint32_t mask = pow2 - 1;
    Dev d[2];
    for (int i = 0; i < 2; i++) {
        d[i].id = i;
        CK(cudaSetDevice(i));
        CK(cudaDeviceEnablePeerAccess(1 - i, 0));
        CK(cudaMalloc(&d[i].tab, table_bytes + 4096)); CK(cudaMalloc(&d[i].tab2, table_bytes + 4096)); CK(cudaMalloc(&d[i].sink, 4));
        CK(cudaMemset(d[i].tab, 0, table_bytes)); CK(cudaMemset(d[i].tab2, 0, table_bytes));
        CK(cudaStreamCreate(&d[i].s)); CK(cudaEventCreate(&d[i].a)); CK(cudaEventCreate(&d[i].b));
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int blocks = prop.multiProcessorCount * 2; // two CTAs of 97 KB per SM, like the learn kernel
    const uint32_t iters = 400;
    const double recs = (double)blocks * iters, row_payload = recs * ROWS * ROW_BYTES;
    printf("device %s x2, %d CTAs, %d rows of %d B per record, table 2 x %llu MB per GPU\n", prop.name, blocks, ROWS, ROW_BYTES, (unsigned long long)(table_bytes >> 20));
    const char *names[] = {"bulk load (pull rows)", "bulk reduce .add.f32 (push to owner L2)", "bulk store (push to staging)", "ld.v4 16 B", "red.v4 16 B", "atom.v4 16 B (return)",
                           "record: bulk load w+acc, bulk reduce w+acc"};
#define ROW(M, mult) { \
        float l = timed<M>(d[0], d[0].tab, d[0].tab2, nullptr, nullptr, nullptr, mask, iters, blocks); \
        float p = timed<M>(d[0], d[1].tab, d[1].tab2, nullptr, nullptr, nullptr, mask, iters, blocks); \
        float q = timed<M>(d[0], d[1].tab, d[1].tab2, &d[1], d[0].tab, d[0].tab2, mask, iters, blocks); \
        printf("  %-46s local %7.1f GB/s | peer %7.1f GB/s | peer, both directions at once %7.1f GB/s per GPU  (%.2f M records/s per GPU)\n", names[M], \
               mult * row_payload / l * 1e-6, mult * row_payload / p * 1e-6, mult * row_payload / q * 1e-6, recs / q * 1e-3); }
    ROW(LOAD_BULK, 1) ROW(RED_BULK, 1) ROW(STORE_BULK, 1) ROW(LD16, 1) ROW(RED16, 1) ROW(ATOM16, 1) ROW(RECORD_BULK, 4)
    return 0;
}
