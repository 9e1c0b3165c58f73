// fwgpu_head.cuh -- sm_100a device code for the dense head of the "one" topology (BASELINE config 5):
//   x = [LR combo outputs, triangle(FFM outputs)]  -> copy -> hidden layers (BlockNeuronLayer + BlockRELU)
//   -> join [h, x] -> one neuron (init One) -> sigmoid          (regressor.rs:191-320)
// replacing, for a sub-batch of B examples evaluated against one weight snapshot,
//   BlockNeuronLayer::forward_backward   (block_neural.rs:196-341)
//   BlockRELU                            (block_relu.rs:79-111)
//   BlockCopy / BlockJoin                (block_misc.rs:435-519)
//   BlockSigmoid                         (block_loss_functions.rs:105-153)
//
// The reference walks one example at a time: for every neuron j and input i it does
//   g = g_j * x_i ; acc_ji += g^2 ; w_ji -= g * LUT[acc_ji]        (block_neural.rs:266-305)
// The device evaluates B examples per pass (B = the examples in flight, 1 in sequential/parity mode):
//   forward   H_l = act(H_{l-1} W_l^T + b_l)                       one tiled fp32 GEMM per layer (bias + ReLU fused)
//   backward  dH_{l-1} = (dZ_l W_l) .* relu'                       one GEMM per layer (mask / direct path fused)
//   update    G1_ji = sum_b g_bj x_bi ,  G2_ji = sum_b (g_bj x_bi)^2      one GEMM per layer computing BOTH sums
//             acc_ji += G2_ji ; w_ji -= G1_ji * LUT[acc_ji]               (per-example squared gradients, as AdaGrad
//                                                                          would have accumulated them one by one)
// For B = 1 every sum has one term and the arithmetic is the reference's, rounding included.
// The contractions in THIS file are fp32 FFMA: for small sub-batches (and the one-example parity mode) the arithmetic is
// the reference's.  Sub-batches of >= 512 rows run the same four GEMM kinds on the tensor cores (fwgpu_umma.cuh,
// tcgen05 with 3xTF32 split operands: single-pass TF32 does not keep predictions within 1e-5).
#pragma once
#include "fwgpu_kernels.cuh"

namespace fwgpu {

enum { HEAD_EPI_BIAS_ACT = 0, HEAD_EPI_MASK = 1, HEAD_EPI_ADD_DIRECT = 2, HEAD_EPI_SUMS = 3 };

struct HeadGemmParams {
    const float *A; uint32_t lda;   // A(m,k): A_T ? A[k*lda + m] : A[m*lda + k]
    const float *B; uint32_t ldb;   // B(n,k): B_T ? B[k*ldb + n] : B[n*ldb + k]
    float *C; uint32_t ldc;         // C[m*ldc + n]
    uint32_t M, N, K;
    uint32_t k_split;               // HEAD_EPI_SUMS: rows of K handled per blockIdx.z
    uint32_t pad_;
    // epilogues
    const float *bias; int relu;    // BIAS_ACT: C = act(acc + bias[n]); a ReLU-clamped output is stored as -0.0f (its mask bit)
    const float *mask_src; uint32_t ld_mask; int mask_on; // MASK: C = mask_src[m][n] is -0.0f ? 0 : acc    (block_relu.rs:101-108)
    const float *direct_w; const float *dy;               // ADD_DIRECT: C = acc + direct_w[n] * dy[m]     (block_misc.rs:452-473)
    float *G1, *G2; float *G1_bias, *G2_bias;             // SUMS: atomicAdd into G1/G2[m*ldc + n]; bias sums from A alone
};

// BM x BN x 16 tiles, 256 threads (16 x 16).  Each thread owns (BM/16) x (BN/16) outputs arranged as 4-wide groups that are
// BM/2 (BN/2) apart, so the 128-bit shared-memory reads of a quarter-warp fall into distinct banks.  The next k-tile is
// fetched into registers while the current one is multiplied (one barrier pair per tile).
template <bool A_T, bool B_T, int EPI, int BM, int BN>
__global__ void __launch_bounds__(256, (BM * BN > 64 * 128) ? 1 : 2) k_head_gemm(const HeadGemmParams p)
{
    constexpr int BK = 16, PAD = 4;
    constexpr int TM = BM / 16, TN = BN / 16;       // outputs per thread: 4 or 8 in each direction
    constexpr int GM = TM / 4, GN = TN / 4;         // 4-wide groups
    constexpr int LA = BM * BK / 256, LB = BN * BK / 256; // tile elements fetched per thread
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const uint32_t tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    uint32_t k_begin = 0, k_end = p.K;
    if (EPI == HEAD_EPI_SUMS) { k_begin = blockIdx.z * p.k_split; k_end = min(p.K, k_begin + p.k_split); }
    float acc[TM][TN], acc2[EPI == HEAD_EPI_SUMS ? TM : 1][EPI == HEAD_EPI_SUMS ? TN : 1], bsum[TM], bsum2[TM];
#pragma unroll
    for (int i = 0; i < TM; i++) {
        bsum[i] = bsum2[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < TN; j++) { acc[i][j] = 0.0f; if (EPI == HEAD_EPI_SUMS) acc2[i][j] = 0.0f; }
    }
    float ra[LA], rb[LB];
    auto fetch = [&](uint32_t k0) {
#pragma unroll
        for (int r = 0; r < LA; r++) {
            const uint32_t e = tid + 256 * r;
            const uint32_t kk = A_T ? e / BM : e % BK, mm = A_T ? e % BM : e / BK;
            const uint32_t gm = m0 + mm, gk = k0 + kk;
            ra[r] = (gm < p.M && gk < k_end) ? (A_T ? p.A[(size_t)gk * p.lda + gm] : p.A[(size_t)gm * p.lda + gk]) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < LB; r++) {
            const uint32_t e = tid + 256 * r;
            const uint32_t kk = B_T ? e / BN : e % BK, nn = B_T ? e % BN : e / BK;
            const uint32_t gn = n0 + nn, gk = k0 + kk;
            rb[r] = (gn < p.N && gk < k_end) ? (B_T ? p.B[(size_t)gk * p.ldb + gn] : p.B[(size_t)gn * p.ldb + gk]) : 0.0f;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int r = 0; r < LA; r++) { const uint32_t e = tid + 256 * r; As[A_T ? e / BM : e % BK][A_T ? e % BM : e / BK] = ra[r]; }
#pragma unroll
        for (int r = 0; r < LB; r++) { const uint32_t e = tid + 256 * r; Bs[B_T ? e / BN : e % BK][B_T ? e % BN : e / BK] = rb[r]; }
    };
    if (k_begin < k_end) fetch(k_begin);
    for (uint32_t k0 = k_begin; k0 < k_end; k0 += BK) {
        stage();
        __syncthreads();
        if (k0 + BK < k_end) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < GM; g++) {
                const float4 v = *reinterpret_cast<const float4 *>(&As[kk][g * (BM / GM) + ty * 4]);
                a[4 * g] = v.x; a[4 * g + 1] = v.y; a[4 * g + 2] = v.z; a[4 * g + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < GN; g++) {
                const float4 v = *reinterpret_cast<const float4 *>(&Bs[kk][g * (BN / GN) + tx * 4]);
                b[4 * g] = v.x; b[4 * g + 1] = v.y; b[4 * g + 2] = v.z; b[4 * g + 3] = v.w;
            }
#pragma unroll
            float a2[EPI == HEAD_EPI_SUMS ? TM : 1], b2[EPI == HEAD_EPI_SUMS ? TN : 1];
            if (EPI == HEAD_EPI_SUMS) {
#pragma unroll
                for (int i = 0; i < TM; i++) a2[i] = __fmul_rn(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < TN; j++) b2[j] = __fmul_rn(b[j], b[j]);
            }
#pragma unroll
            for (int i = 0; i < TM; i++) {
                if (EPI == HEAD_EPI_SUMS) { bsum[i] += a[i]; bsum2[i] += a2[i]; }
#pragma unroll
                for (int j = 0; j < TN; j++) {
                    acc[i][j] = fmaf(a[i], b[j], acc[i][j]); // sum of the per-example gradients g_j * x_i (block_neural.rs:268-269)
                    // sum of their squares as (g_j^2)(x_i^2): one FMA per term instead of multiply, square, add
                    // (optimizer.rs:148-151 squares the product; the two differ by one rounding of ~6e-8 relative)
                    if (EPI == HEAD_EPI_SUMS) acc2[i][j] = fmaf(a2[i], b2[j], acc2[i][j]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const uint32_t gm = m0 + (i / 4) * (BM / GM) + ty * 4 + (i & 3);
        if (gm >= p.M) continue;
        if (EPI == HEAD_EPI_SUMS && blockIdx.x == 0 && tx == 0 && p.G1_bias) {
            if (bsum[i] != 0.0f || bsum2[i] != 0.0f) { atomicAdd(p.G1_bias + gm, bsum[i]); atomicAdd(p.G2_bias + gm, bsum2[i]); }
        }
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const uint32_t gn = n0 + (j / 4) * (BN / GN) + tx * 4 + (j & 3);
            if (gn >= p.N) continue;
            const size_t o = (size_t)gm * p.ldc + gn;
            if (EPI == HEAD_EPI_BIAS_ACT) {
                float y = __fadd_rn(p.bias[gn], acc[i][j]);          // output = bias, then sgemv adds W x (block_neural.rs:207-220)
                // block_relu.rs:88-97: x < 0 -> output 0, derivative 0 (stored as -0.0f = the mask); anything else passes with
                // derivative 1 -- including a pre-activation that is exactly -0.0, which is therefore stored as +0.0
                if (p.relu) y = y < 0.0f ? -0.0f : __fadd_rn(y, 0.0f);
                p.C[o] = y;
            } else if (EPI == HEAD_EPI_MASK) {
                float v = acc[i][j];
                if (p.mask_on && __float_as_uint(p.mask_src[(size_t)gm * p.ld_mask + gn]) == 0x80000000u) v = 0.0f;
                p.C[o] = v;
            } else if (EPI == HEAD_EPI_ADD_DIRECT) {
                p.C[o] = __fadd_rn(acc[i][j], __fmul_rn(p.direct_w[gn], p.dy[gm]));
            } else {
                if (acc[i][j] != 0.0f || acc2[i][j] != 0.0f) { atomicAdd(p.G1 + o, acc[i][j]); atomicAdd(p.G2 + o, acc2[i][j]); }
            }
        }
    }
}

// Final neuron (one output, inputs [h, x]) + sigmoid + logloss gradient, one warp per example
// (block_neural.rs:196-222 with num_neurons = 1, block_loss_functions.rs:105-153), and -- when updating -- the
// gradient of the last hidden layer's pre-activations: dZ[b][j] = w_out[j] * g_b * relu'(h_bj).
struct HeadFinalParams {
    const float *H; uint32_t ldh, n_h;   // last hidden layer's outputs
    const float *X; uint32_t ldx, n_x;   // the head's input (direct path)
    const float *w;                      // final neuron: [n_h + n_x] weights, then the bias
    const float *label, *importance;     // per row
    const uint32_t *out_index;           // where the prediction of row b goes (nullptr: preds[b])
    float *preds; float *dy; float *dZ; uint32_t ldz;
    uint32_t n_rows; int update; int h_relu;
};
__global__ void __launch_bounds__(256) k_head_final(const HeadFinalParams p)
{
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = warp; b < p.n_rows; b += n_warps) {
        const float *h = p.H + (size_t)b * p.ldh, *x = p.X + (size_t)b * p.ldx;
        float s = 0.0f;
        for (uint32_t i = lane; i < p.n_h; i += 32) s = fmaf(p.w[i], h[i], s);
        for (uint32_t i = lane; i < p.n_x; i += 32) s = fmaf(p.w[p.n_h + i], x[i], s);
        s = warp_sum(s);
        const float y = __fadd_rn(p.w[p.n_h + p.n_x], s);
        const float label = p.label[b], importance = p.importance[b];
        float pr, g;
        if (isnan(y)) { pr = logistic(0.0f); g = 0.0f; }
        else if (y < -50.0f) { pr = logistic(-50.0f); g = 0.0f; }
        else if (y > 50.0f) { pr = logistic(50.0f); g = 0.0f; }
        else { pr = logistic(y); g = __fmul_rn(-__fsub_rn(label, pr), importance); }
        if (!(p.update && importance != 0.0f)) g = 0.0f; // regressor.rs:366-370: no update for this example
        if (lane == 0) { p.preds[p.out_index ? p.out_index[b] : b] = pr; p.dy[b] = g; }
        if (p.update) {
            float *dz = p.dZ + (size_t)b * p.ldz;
            for (uint32_t j = lane; j < p.n_h; j += 32) {
                float v = __fmul_rn(p.w[j], g);
                if (p.h_relu && __float_as_uint(h[j]) == 0x80000000u) v = 0.0f;
                dz[j] = v;
            }
        }
    }
}

// Gradient sums of the FINAL neuron (one output, inputs [h, x], block_neural.rs:266-305 with num_neurons = 1):
//   G1[i] += sum_b g_b in_b[i],  G2[i] += sum_b (g_b in_b[i])^2,  and the bias sums from g alone.
// A matrix-vector product: one thread per input column (coalesced rows), blockIdx.y takes a slice of the rows.
struct HeadFinalSumsParams {
    const float *H; uint32_t ldh, n_h; const float *X; uint32_t ldx, n_x;
    const float *dy; uint32_t n_rows, rows_per_block;
    float *G1, *G2; // [n_h + n_x] then the bias
};
__global__ void __launch_bounds__(256) k_head_final_sums(const HeadFinalSumsParams p)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x, n_in = p.n_h + p.n_x;
    const uint32_t b0 = blockIdx.y * p.rows_per_block, b1 = min(p.n_rows, b0 + p.rows_per_block);
    __shared__ float sdy[256];
    float s1 = 0.0f, s2 = 0.0f, t1 = 0.0f, t2 = 0.0f;
    const bool is_h = i < p.n_h;
    const float *col = is_h ? p.H + i : p.X + (i - p.n_h);
    const uint32_t ld = is_h ? p.ldh : p.ldx;
    for (uint32_t base = b0; base < b1; base += 256) {
        __syncthreads();
        sdy[threadIdx.x] = base + threadIdx.x < b1 ? p.dy[base + threadIdx.x] : 0.0f;
        __syncthreads();
        const uint32_t cnt = min(256u, b1 - base);
        if (i < n_in) {
            for (uint32_t r = 0; r < cnt; r++) {
                const float g = __fmul_rn(sdy[r], col[(size_t)(base + r) * ld]);
                s1 += g; s2 = fmaf(g, g, s2);
            }
        } else if (i == n_in) {
            for (uint32_t r = 0; r < cnt; r++) { t1 += sdy[r]; t2 = fmaf(sdy[r], sdy[r], t2); }
        }
    }
    if (i < n_in) { if (s1 != 0.0f || s2 != 0.0f) { atomicAdd(p.G1 + i, s1); atomicAdd(p.G2 + i, s2); } }
    else if (i == n_in) { if (t1 != 0.0f || t2 != 0.0f) { atomicAdd(p.G1 + i, t1); atomicAdd(p.G2 + i, t2); } }
}

// acc += G2 ; w -= G1 * step(acc)  over every head parameter at once; clears G1/G2 for the next sub-batch
// (optimizer.rs:35-37, 76-89, 147-156 with the nn_* hyper-parameters, block_neural.rs:108-109).
__global__ void __launch_bounds__(256) k_head_apply(float *w, float *acc, float *G1, float *G2, size_t n, uint32_t optimizer,
                                                    const float *__restrict__ lut, float lr, float mpt)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float g1 = G1[i], g2 = G2[i];
        if (g1 == 0.0f && g2 == 0.0f) continue;
        G1[i] = 0.0f; G2[i] = 0.0f;
        float upd;
        if (optimizer == OPT_SGD) upd = __fmul_rn(g1, lr);
        else {
            const float a = __fadd_rn(acc[i], g2);
            acc[i] = a;
            upd = opt_step(optimizer, g1, a, lut, lr, mpt);
        }
        w[i] = __fsub_rn(w[i], upd);
    }
}

} // namespace fwgpu
