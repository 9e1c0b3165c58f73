// fwgpu_umma.cuh -- tcgen05 (5th-generation tensor core) GEMM tiles for the dense head, sm_100a only.
//
//   C[m][n] (+)= sum_k A(m,k) * B(n,k)          fp32 in, fp32 out, fp32 accumulation in tensor memory
//
// The head's contractions (block_neural.rs:196-341, restated for a sub-batch in fwgpu_head.cuh) must keep predictions
// within 1e-5 of the reference, which a single TF32 pass (10-bit mantissa) does not.  Every operand element v is
// therefore split while it is staged into shared memory,
//   hi = v rounded to TF32 (10 mantissa bits),   lo = v - hi  (exact; the tensor core reads its upper 19 bits),
// and each k-step issues three MMAs into the same accumulator:  hi*hi + lo*hi + hi*lo  (the lo*lo term is < 2^-22 |ab|).
//
// Structure of one CTA (1024 threads, one 128 x 128 output tile, BK = 32; BK = 16 for the update GEMM, which stages the
// squared tiles as well and keeps two accumulators):
//   * all threads load the A / B tile from global memory (either operand may be stored transposed), split it and write
//     it to shared memory in the canonical K-major SWIZZLE_128B / SWIZZLE_64B UMMA layout (umma_chunk_index below);
//   * fence.proxy.async + __syncthreads, then ONE thread issues 4 k-steps x 3 tcgen05.mma.kind::tf32 (M = 128, N = 128,
//     K = 8) and commits them to the stage's mbarrier; the tensor core works while all threads stage the next tile into
//     the other buffer (two stages);
//   * epilogue: tcgen05.ld 32 lanes x 16 columns per warp, fused epilogue, global stores / atomics.
#pragma once
#include "fwgpu_head.cuh"

namespace fwgpu {

#ifdef UMMA_DEBUG_TIMING
__device__ unsigned long long umma_dbg[16];
#define UMMA_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); umma_dbg[i] = t_; } } while (0)
#else
#define UMMA_STAMP(i) do { } while (0)
#endif

constexpr int UMMA_BM = 128, UMMA_BN = 128, UMMA_BK = 32;
constexpr int UMMA_PRODUCER_WARPS = 16, UMMA_THREADS = (UMMA_PRODUCER_WARPS + 2) * 32, UMMA_STAGES = 3;
constexpr int UMMA_STAGE_BYTES = 4 * UMMA_BM * UMMA_BK * 4;     // A_hi, A_lo, B_hi, B_lo of 16 KB each (update GEMM: 8 tiles of 8 KB)
constexpr int UMMA_SMEM_BYTES = UMMA_STAGES * UMMA_STAGE_BYTES + 128; // stages + barriers + tmem address (the epilogue reuses the stages)
constexpr int UMMA_PREFETCH = 3;                                // k-blocks of operand chunks in flight per thread

__device__ __forceinline__ uint32_t umma_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// hi part of the split: v rounded to TF32's 10 mantissa bits, with integer arithmetic.  cvt.rna.tf32.f32 does the same but
// issues at the conversion unit's quarter rate: two of them per element made the producers, not the tensor core, the bound
// of the k-loop (0.97 us of staging per k-block against 0.51 us of MMA work, tools/umma_gemm_test.cu experiment builds).
__device__ __forceinline__ float umma_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

// Operand tiles are K-major with one row = BK floats = 128 bytes (BK = 32, SWIZZLE_128B) or 64 bytes (BK = 16, SWIZZLE_64B):
// swizzle atoms of 8 rows, the 16-byte chunk c of row r stored at chunk position c ^ (r & 7)  [128B]  /  c ^ ((r >> 1) & 3)  [64B]
// (cute Swizzle<3,4,3> / Swizzle<2,4,3>: address bits [4,7) ^= bits [7,10)).  A first version used the un-swizzled
// core-matrix layout: it computed the same numbers 4.4x slower than the tensor core's floor -- an SS-mode TF32 MMA reads
// 8 KB of operands per 64 cycles, i.e. the whole shared-memory bandwidth of the SM, and only swizzled rows deliver it.
template <int BK> __device__ __forceinline__ uint32_t umma_chunk_index(uint32_t row, uint32_t kc)
{
    constexpr uint32_t ROW_CHUNKS = BK / 4;
    const uint32_t sw = BK == 32 ? (row & 7u) : ((row >> 1) & 3u);
    return row * ROW_CHUNKS + (kc ^ sw); // 16-byte units
}

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address, LBO (unused by swizzled
// K-major layouts, 1), SBO = bytes between 8-row groups, version 1, layout type SWIZZLE_128B (2) / SWIZZLE_64B (4)
template <int BK> __device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(((8u * BK * 4u) >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(BK == 32 ? 2u : 4u) << 61;
    return d;
}
// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t UMMA_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UMMA_BN >> 3) << 17) | ((uint32_t)(UMMA_BM >> 4) << 24);

__device__ __forceinline__ void umma_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(UMMA_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(umma_smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void umma_bar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(umma_smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void umma_bar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(umma_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// One operand tile in flight: the 16-byte chunks this thread stages (loaded one k-block ahead, so the global loads overlap
// the barrier, the MMA issue and the wait for the buffer).
// element (row, k) of an operand: TRANS ? src[k * ld + row] : src[row * ld + k]; rows >= n_rows and k >= k_end read as 0;
// ONES: row == n_rows reads as 1 (the bias column of the head's update GEMM: sums of the gradients themselves).
// With 1024 threads a thread owns ONE chunk per operand when BK = 32 (1024 chunks per tile); with BK = 16 (512 chunks)
// the lower half of the block stages A and the upper half B.
template <int BK, bool TRANS> __device__ __forceinline__ void umma_chunk_pos(uint32_t idx, uint32_t &r8, uint32_t &kc, uint32_t &rg)
{
    constexpr uint32_t KC = BK / 4;
    if (!TRANS) { r8 = idx & 7; kc = (idx >> 3) % KC; rg = (idx >> 3) / KC; } // a quarter warp = 8 rows x one 16-byte chunk
    else { r8 = idx & 7; rg = (idx >> 3) & 15; kc = idx >> 7; }                 // a warp = 32 consecutive rows of one k: coalesced
}

template <int BK, bool TRANS, bool ONES>
__device__ __forceinline__ float4 umma_load_chunk(uint32_t idx, const float *__restrict__ src, uint32_t ld, uint32_t row0, uint32_t n_rows,
                                                  uint32_t k0, uint32_t k_end, bool vec_ok)
{
    uint32_t r8, kc, rg;
    umma_chunk_pos<BK, TRANS>(idx, r8, kc, rg);
    const uint32_t row = row0 + rg * 8 + r8, k = k0 + kc * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n_rows) {
        if (!TRANS) {
            const float *s = src + (size_t)row * ld + k;
            if (vec_ok && k + 3 < k_end) v = __ldg(reinterpret_cast<const float4 *>(s));
            else {
                if (k < k_end) v.x = __ldg(s);
                if (k + 1 < k_end) v.y = __ldg(s + 1);
                if (k + 2 < k_end) v.z = __ldg(s + 2);
                if (k + 3 < k_end) v.w = __ldg(s + 3);
            }
        } else {
            const float *s = src + (size_t)k * ld + row;
            if (k < k_end) v.x = __ldg(s);
            if (k + 1 < k_end) v.y = __ldg(s + ld);
            if (k + 2 < k_end) v.z = __ldg(s + 2 * (size_t)ld);
            if (k + 3 < k_end) v.w = __ldg(s + 3 * (size_t)ld);
        }
    } else if (ONES && row == n_rows) {
        v.x = k < k_end ? 1.0f : 0.0f; v.y = k + 1 < k_end ? 1.0f : 0.0f; v.z = k + 2 < k_end ? 1.0f : 0.0f; v.w = k + 3 < k_end ? 1.0f : 0.0f;
    }
    return v;
}

// split a loaded chunk into hi / lo TF32 parts and write them in the canonical layout; SQUARE stages v*v instead of v
template <int BK, bool TRANS, bool SQUARE>
__device__ __forceinline__ void umma_store_chunk(uint32_t idx, float4 v, float4 *hi, float4 *lo)
{
    uint32_t r8, kc, rg;
    umma_chunk_pos<BK, TRANS>(idx, r8, kc, rg);
    if (SQUARE) { v.x = __fmul_rn(v.x, v.x); v.y = __fmul_rn(v.y, v.y); v.z = __fmul_rn(v.z, v.z); v.w = __fmul_rn(v.w, v.w); }
    float4 h, l;
    h.x = umma_tf32(v.x); h.y = umma_tf32(v.y); h.z = umma_tf32(v.z); h.w = umma_tf32(v.w);
    // lo = v - hi is exact in fp32 and |lo| <= 2^-11 |v|; the tensor core reads its upper 19 bits, i.e. drops < 2^-21 |v|
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    const uint32_t chunk = umma_chunk_index<BK>(rg * 8 + r8, kc);
    hi[chunk] = h; lo[chunk] = l;
}

// the MMAs of one staged k-block (BK / 8 k-steps, three split products each)
template <int BK>
__device__ __forceinline__ void umma_issue_block(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, bool first)
{
#pragma unroll
    for (int ks = 0; ks < BK / 8; ks++) {
        const uint32_t off = ks * 32; // 8 floats along K inside the swizzled row
        umma_mma_tf32(tmem_d, umma_desc<BK>(a_hi + off), umma_desc<BK>(b_hi + off), (first && ks == 0) ? 0u : 1u);
        umma_mma_tf32(tmem_d, umma_desc<BK>(a_lo + off), umma_desc<BK>(b_hi + off), 1u);
        umma_mma_tf32(tmem_d, umma_desc<BK>(a_hi + off), umma_desc<BK>(b_lo + off), 1u);
    }
}

__device__ __forceinline__ void umma_bar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma_smem_u32(bar)) : "memory"); }

// Same parameters and epilogues as k_head_gemm (fwgpu_head.cuh).  HEAD_EPI_SUMS: two accumulators, G1 += A B^T and
// G2 += (A.A)(B.B)^T over this block's slice of K, added atomically; column N of B reads as ones and lands in G*_bias.
//
// Warp roles (UMMA_THREADS = 576): warps 0..15 are PRODUCERS (load, split, store one stage, arrive on full[s]), warp 16
// is the MMA warp (waits full[s], one lane issues the stage's MMAs and commits them to empty[s]), warp 17 allocates the
// tensor memory.  UMMA_STAGES = 3 stages of 64 KB decouple them: the tensor core, the stores and the global loads of
// three different k-blocks run concurrently.  After the last commit all 18 warps run the epilogue: the 128 x 128
// accumulator goes TMEM -> registers -> shared memory (row pitch 132 floats) and leaves as coalesced 512-byte rows.
template <bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(UMMA_THREADS, 1) k_umma_gemm(const HeadGemmParams p)
{
    extern __shared__ __align__(1024) unsigned char umma_smem[];
    constexpr bool SUMS = EPI == HEAD_EPI_SUMS;
    constexpr int BK = SUMS ? 16 : 32;                       // SUMS stages the squares as well: half the depth, same bytes
    constexpr uint32_t TILE = UMMA_BM * BK * 4;              // bytes of one hi or lo tile
    constexpr uint32_t TMEM_COLS = SUMS ? 256 : 128;
    constexpr uint32_t PRODUCERS = UMMA_PRODUCER_WARPS * 32; // 512
    constexpr uint32_t CHUNKS = UMMA_BM * BK / 4;            // 16-byte chunks per operand tile: 1024 or 512
    constexpr uint32_t CPT = CHUNKS / PRODUCERS;             // chunks per producer thread and operand: 2 or 1
    uint64_t *full = reinterpret_cast<uint64_t *>(umma_smem + UMMA_STAGES * UMMA_STAGE_BYTES), *empty = full + UMMA_STAGES, *done = empty + UMMA_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m0 = blockIdx.y * UMMA_BM, n0 = blockIdx.x * UMMA_BN;
    uint32_t k_begin = 0, k_end = p.K;
    if (SUMS) { k_begin = blockIdx.z * p.k_split; k_end = min(p.K, k_begin + p.k_split); }
    const uint32_t n_kb = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;

    UMMA_STAMP(0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < UMMA_STAGES; i++) { umma_bar_init(full + i, UMMA_PRODUCER_WARPS); umma_bar_init(empty + i, 1); }
        umma_bar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UMMA_PRODUCER_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma_smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    UMMA_STAMP(1);

    if (warp < UMMA_PRODUCER_WARPS) {
        // ---------------- producers ----------------
        const bool a_vec = !A_T && (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
        const bool b_vec = !B_T && (p.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);
        // register ring: the chunks of the next UMMA_PREFETCH k-blocks are in flight while this one is split and stored
        float4 ca[UMMA_PREFETCH][CPT], cb[UMMA_PREFETCH][CPT];
#pragma unroll
        for (int j = 0; j < UMMA_PREFETCH; j++) {
            const uint32_t kj = k_begin + j * BK;
#pragma unroll
            for (uint32_t c = 0; c < CPT; c++) {
                ca[j][c] = cb[j][c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kj < k_end) {
                    ca[j][c] = umma_load_chunk<BK, A_T, false>(threadIdx.x + c * PRODUCERS, p.A, p.lda, m0, p.M, kj, k_end, a_vec);
                    cb[j][c] = umma_load_chunk<BK, B_T, SUMS>(threadIdx.x + c * PRODUCERS, p.B, p.ldb, n0, p.N, kj, k_end, b_vec);
                }
            }
        }
        uint32_t it = 0;
        for (uint32_t k0 = k_begin; k0 < k_end;) {
#pragma unroll
            for (int j = 0; j < UMMA_PREFETCH; j++) {
                if (k0 >= k_end) break;
                const uint32_t s = it % UMMA_STAGES;
                if (it >= UMMA_STAGES) umma_bar_wait(empty + s, ((it / UMMA_STAGES) - 1) & 1); // the MMAs that read this buffer are done
                unsigned char *st = umma_smem + s * UMMA_STAGE_BYTES;
#ifndef UMMA_EXPERIMENT_NO_STORE
#pragma unroll
                for (uint32_t c = 0; c < CPT; c++) {
                    const uint32_t idx = threadIdx.x + c * PRODUCERS;
                    umma_store_chunk<BK, A_T, false>(idx, ca[j][c], reinterpret_cast<float4 *>(st), reinterpret_cast<float4 *>(st + TILE));
                    umma_store_chunk<BK, B_T, false>(idx, cb[j][c], reinterpret_cast<float4 *>(st + 2 * TILE), reinterpret_cast<float4 *>(st + 3 * TILE));
                    if (SUMS) { // the same tiles squared: block_neural.rs:268-271 accumulates (g_j x_i)^2 per example
                        umma_store_chunk<BK, A_T, true>(idx, ca[j][c], reinterpret_cast<float4 *>(st + 4 * TILE), reinterpret_cast<float4 *>(st + 5 * TILE));
                        umma_store_chunk<BK, B_T, true>(idx, cb[j][c], reinterpret_cast<float4 *>(st + 6 * TILE), reinterpret_cast<float4 *>(st + 7 * TILE));
                    }
                }
#endif
                const uint32_t kn = k0 + UMMA_PREFETCH * BK;
                if (kn < k_end) {
#pragma unroll
                    for (uint32_t c = 0; c < CPT; c++) {
                        ca[j][c] = umma_load_chunk<BK, A_T, false>(threadIdx.x + c * PRODUCERS, p.A, p.lda, m0, p.M, kn, k_end, a_vec);
                        cb[j][c] = umma_load_chunk<BK, B_T, SUMS>(threadIdx.x + c * PRODUCERS, p.B, p.ldb, n0, p.N, kn, k_end, b_vec);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) umma_bar_arrive(full + s);
                k0 += BK; it++;
            }
        }
    } else if (warp == UMMA_PRODUCER_WARPS) {
        // ---------------- MMA warp ----------------
        for (uint32_t it = 0; it < n_kb; it++) {
            const uint32_t s = it % UMMA_STAGES;
            umma_bar_wait(full + s, (it / UMMA_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sb = umma_smem_u32(umma_smem + s * UMMA_STAGE_BYTES);
#ifndef UMMA_EXPERIMENT_NO_MMA // (tools/umma_gemm_test.cu: which side bounds the k-loop)
                umma_issue_block<BK>(tmem, sb, sb + TILE, sb + 2 * TILE, sb + 3 * TILE, it == 0);
                if (SUMS) umma_issue_block<BK>(tmem + 128, sb + 4 * TILE, sb + 5 * TILE, sb + 6 * TILE, sb + 7 * TILE, it == 0);
#endif
                umma_commit(empty + s);
                if (it + 1 == n_kb) umma_commit(done); // MMAs complete in order: this commit covers all of them
            }
            __syncwarp();
        }
    }
    UMMA_STAMP(3);
    if (n_kb) umma_bar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    UMMA_STAMP(4);

    // ---- epilogue, phase 1: warps 0..15 move the accumulator(s) TMEM -> registers -> shared memory ----
    // warp w reads TMEM lanes 32*(w%4).. (its rows) and the 32 columns (w/4)*32 .. ; the stage buffers are free now
    constexpr uint32_t PITCH = UMMA_BN + 4; // floats per staged row: 16-byte stores of 8 consecutive rows hit distinct banks
    float *T = reinterpret_cast<float *>(umma_smem);
    if (warp < UMMA_PRODUCER_WARPS) {
#pragma unroll 1
        for (uint32_t acc_i = 0; acc_i < (SUMS ? 2u : 1u); acc_i++) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((32u * (warp & 3)) << 16) + acc_i * 128 + (warp >> 2) * 32;
            if (n_kb) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) r[j] = 0u;
            }
            float4 *dst = reinterpret_cast<float4 *>(T + (size_t)acc_i * UMMA_BM * PITCH + (size_t)(32 * (warp & 3) + lane) * PITCH + (warp >> 2) * 32);
#pragma unroll
            for (int j = 0; j < 8; j++) dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    UMMA_STAMP(5);
    if (warp == UMMA_PRODUCER_WARPS + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");

    // ---- epilogue, phase 2: every warp takes whole rows; a lane owns 4 consecutive columns (512 bytes per warp and row) ----
    const uint32_t n = n0 + 4 * lane;
    const bool c_vec = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(SUMS ? p.G1 : p.C) & 15) == 0) && (!SUMS || (reinterpret_cast<uintptr_t>(p.G2) & 15) == 0);
    for (uint32_t rr = warp; rr < UMMA_BM; rr += UMMA_THREADS / 32) {
        const uint32_t row = m0 + rr;
        if (row >= p.M) break;
#pragma unroll 1
        for (uint32_t acc_i = 0; acc_i < (SUMS ? 2u : 1u); acc_i++) {
            const float4 q = *reinterpret_cast<const float4 *>(T + (size_t)acc_i * UMMA_BM * PITCH + (size_t)rr * PITCH + 4 * lane);
            float v[4] = {q.x, q.y, q.z, q.w};
            if (SUMS) {
                float *G = acc_i ? p.G2 : p.G1, *Gb = acc_i ? p.G2_bias : p.G1_bias;
                if (c_vec && n + 3 < p.N) {
                    if (v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f || v[3] != 0.0f)
                        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(G + (size_t)row * p.ldc + n), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (v[j] == 0.0f) continue;
                        if (n + j < p.N) atomicAdd(G + (size_t)row * p.ldc + n + j, v[j]);
                        else if (n + j == p.N && Gb) atomicAdd(Gb + row, v[j]);
                    }
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (n + j >= p.N) continue;
                if (EPI == HEAD_EPI_BIAS_ACT) {
                    v[j] = __fadd_rn(p.bias[n + j], v[j]);
                    if (p.relu) v[j] = v[j] < 0.0f ? -0.0f : __fadd_rn(v[j], 0.0f); // block_relu.rs:88-97: output 0, -0.0f is the mask; an exact -0.0 input passes as +0.0
                } else if (EPI == HEAD_EPI_MASK) {
                    if (p.mask_on && __float_as_uint(p.mask_src[(size_t)row * p.ld_mask + n + j]) == 0x80000000u) v[j] = 0.0f; // block_relu.rs:101-108
                } else if (EPI == HEAD_EPI_ADD_DIRECT) {
                    v[j] = __fadd_rn(v[j], __fmul_rn(p.direct_w[n + j], p.dy[row])); // block_misc.rs:452-473
                }
            }
            float *o = p.C + (size_t)row * p.ldc + n;
            if (c_vec && n + 3 < p.N) *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; j++) if (n + j < p.N) o[j] = v[j];
            }
        }
    }
    UMMA_STAMP(6);
}

} // namespace fwgpu
