// fwgpu_kernels.cuh -- sm_100a device code for the LR/FFM learn/predict hot path.
//
// One fused kernel per mini-batch (k_learn) replaces, per example, the reference's
//   BlockLR::forward_backward      (block_lr.rs:123-151)
//   BlockFFM::forward_backward     (block_ffm.rs:122-314)
//   BlockTriangle                  (block_misc.rs:798-884)
//   BlockSigmoid::forward_backward (block_loss_functions.rs:105-153)
//   OptimizerAdagradLUT/Flex/SGD   (optimizer.rs:15-162)
// and k_translate replaces FeatureBufferTranslator::translate (feature_buffer.rs:178-338).
//
// Design (DESIGN.md has the long form):
//   * a group of T threads (32..256) owns one example at a time; groups stride over the batch;
//   * gather: the F field-summed latent rows C[z][0..F*k) = sum_{j in field z} v_j * W[h_j .. h_j+F*k)
//     are built straight from HBM with 128-bit ld.global.cg loads into shared memory -- for the
//     usual one-feature-per-field example that is exactly one coalesced pass over the n rows;
//   * forward: sum_{z<f} <C[f][z-block], C[z][f-block]> (+ the intra-field term of multi-valued
//     fields) + LR dot, reduced with warp shuffles; sigmoid as block_loss_functions.rs:125-141;
//   * backward: per slot grad = g * v_i * (C[z][f-block] - [z==f] v_i w_i[f-block]); AdaGrad
//     accumulator via ATOMG.ADD.F32x4 (returns the old value, so in-flight examples that hit the
//     same slot each see a larger accumulator, as sequential updates would) and the weight step
//     via REDG.ADD.F32x4 -- lock-free, nothing is lost: Hogwild on device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fwgpu {

struct FastDiv { uint32_t d, m, s; };
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f) { return (__umulhi(n, f.m) + n) >> f.s; } // n < 2^31

struct __align__(16) ExMeta {
    float label, importance;
    uint32_t lr_begin, lr_cnt;
    uint32_t ffm_begin, ffm_cnt;
    uint32_t out_index; // where the prediction goes (index into preds)
    uint32_t pad1;
};

enum { OPT_SGD = 0, OPT_FLEX = 1, OPT_LUT = 2 };

// Models with a dense head (regressor.rs:191-320) run the LR/FFM blocks in two passes around the head's GEMMs:
//   PHASE 1  gather + forward only: row b of X receives the head's input [LR combo outputs, triangle(FFM outputs)]
//            (block_lr.rs:28-47, block_ffm.rs:219-261, block_misc.rs:862-884) plus the example's label / importance;
//   PHASE 2  gather again + update: the gradient of every input comes from row b of dX (what the head's backward pass left
//            on the tape), so each field pair has its own gradient instead of the single sigmoid gradient g
//            (block_misc.rs:814-833 mirrors d_tri onto out[f][z] and out[z][f]; block_ffm.rs:265-288; block_lr.rs:135-151).
// PHASE 0 is the fused single pass of head-less models.  Row b = example index - row_base.
struct HeadIO {
    float *X; const float *dX; uint32_t ldx;
    uint32_t n_lr_out;             // number of LR outputs = num_combos; triangle outputs follow
    float *row_label, *row_importance; uint32_t *row_out_index;
    const float *dy;               // PHASE 2: the sigmoid gradient of row b (0 = nothing to update)
    uint32_t row_base;
};
__device__ __forceinline__ uint32_t tri_index(uint32_t a, uint32_t b) { const uint32_t f = a > b ? a : b, z = a > b ? b : a; return f * (f + 1) / 2 + z; }

struct LearnParams {
    // tables (HBM)
    float2 *lr;          // {w, acc} x (1 << bit_precision)            block_lr.rs:19-25
    float *ffm_w;        // (1 << ffm_bit_precision) + F*k (+ pad)      block_ffm.rs:40
    float *ffm_acc;      //                                            block_ffm.rs:41
    const float *lut_lr; // 2048, optimizer.rs:121-144
    const float *lut_ffm;
    // batch (HBM): AoS entries {hash, value bits, combo|field, 0}
    const ExMeta *meta;
    const uint4 *lr_ent;
    const uint4 *ffm_ent;
    float *preds;
    uint32_t n_examples;
    // shape
    uint32_t F, k, Fk, cpr; // cpr = chunks (of VEC floats) per row
    uint32_t n_cap;         // staging capacity: features per example
    FastDiv div_cpr, div_k, div_F;
    uint32_t optimizer;
    float lr_lr, lr_mpt, ffm_lr, ffm_mpt; // learning rate / minus power_t (Flex, SGD)
    int update;
    uint32_t *err_flag;     // bit0: example exceeded n_cap
    uint32_t group_smem_bytes;
    uint32_t max_groups;    // 0 = all resident groups; else cap on examples in flight (concurrency ramp)
    const uint32_t *n_examples_dev; // when set, the number of examples is read from device memory (leftover list)
    int kv;                 // k % VEC == 0: a 16-byte chunk never straddles two field blocks (all BASELINE shapes)
    uint32_t lr_cap;        // LR hashes staged in shared memory for the duplicate check (0 = read them from global)
    int simple_update;      // debug knob: one chunk at a time (ATOMG -> REDG) instead of rounds of four
    int exact_order;        // sum the sigmoid inputs in the reference's tape order (one example in flight: parity mode)
    int phase;              // 0 = fused single pass; 1 / 2 = the two passes around a dense head (HeadIO)
    HeadIO io;
    int sys_scope;          // tables span several GPUs (fwgpu_create_sharded): parity-mode fences must reach the peers' memory
    unsigned long long *stat_examples; // += examples this launch handles (fwgpu_debug_path_counts: which kernel did the work)
};

// ---------------------------------------------------------------------------------------------
template <int VEC> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<1> { using T = float; };

template <int VEC> __device__ __forceinline__ void ldcg_vec(const float *p, float (&o)[VEC]);
template <> __device__ __forceinline__ void ldcg_vec<4>(const float *p, float (&o)[4]) { float4 v = __ldcg(reinterpret_cast<const float4 *>(p)); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <> __device__ __forceinline__ void ldcg_vec<2>(const float *p, float (&o)[2]) { float2 v = __ldcg(reinterpret_cast<const float2 *>(p)); o[0] = v.x; o[1] = v.y; }
template <> __device__ __forceinline__ void ldcg_vec<1>(const float *p, float (&o)[1]) { o[0] = __ldcg(p); }

// atomic add returning the old vector (ATOMG.E.ADD.F32x4 on sm_100a)
template <int VEC> __device__ __forceinline__ void atom_add_vec(float *p, const float (&v)[VEC], float (&old)[VEC]);
template <> __device__ __forceinline__ void atom_add_vec<4>(float *p, const float (&v)[4], float (&old)[4]) { float4 o = atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3])); old[0] = o.x; old[1] = o.y; old[2] = o.z; old[3] = o.w; }
template <> __device__ __forceinline__ void atom_add_vec<2>(float *p, const float (&v)[2], float (&old)[2]) { float2 o = atomicAdd(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1])); old[0] = o.x; old[1] = o.y; }
template <> __device__ __forceinline__ void atom_add_vec<1>(float *p, const float (&v)[1], float (&old)[1]) { old[0] = atomicAdd(p, v[0]); }

// reduction without return (REDG.E.ADD.F32x4)
template <int VEC> __device__ __forceinline__ void red_add_vec(float *p, const float (&v)[VEC]);
template <> __device__ __forceinline__ void red_add_vec<4>(float *p, const float (&v)[4]) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory"); }
template <> __device__ __forceinline__ void red_add_vec<2>(float *p, const float (&v)[2]) { asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory"); }
template <> __device__ __forceinline__ void red_add_vec<1>(float *p, const float (&v)[1]) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory"); }

template <int T> __device__ __forceinline__ void group_sync(int group_in_block)
{
    if (T == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group_in_block + 1), "r"(T) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// expf exactly as glibc (>= 2.27, sysdeps/ieee754/flt-32/e_expf.c, the ARM optimized-routines algorithm) computes it:
// the reference's logistic() (block_loss_functions.rs:15-17) is Rust's f32::exp = libm expf, and AdagradLUT turns a
// 1-ulp difference in the gradient into a different table bucket sooner or later, so the device reproduces the same
// double-precision evaluation instead of calling CUDA's expf (2 ulp).  Valid for |x| < 88 (the sigmoid clamps at 50).
// tests/test_expf.py checks the host restatement of this routine against libm bit for bit.
__device__ __constant__ uint64_t c_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
__device__ __forceinline__ float expf_libm(float x)
{
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0, C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    const double z = __dmul_rn(InvLn2N, (double)x);
    double kd = __dadd_rn(z, SHIFT);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __dsub_rn(z, kd);
    const uint64_t t = c_exp2f_tab[ki & 31] + (ki << 47);
    const double sc = __longlong_as_double((long long)t);
    const double zz = __dadd_rn(__dmul_rn(C0, r), C1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(zz, r2), y);
    y = __dmul_rn(y, sc);
    return __double2float_rn(y);
}
// logistic(t) = (1.0 + (-t).exp()).recip()   (block_loss_functions.rs:15-17)
__device__ __forceinline__ float logistic(float t) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_libm(-t))); }

// new accumulator = old + g*g with two roundings (optimizer.rs:77-79, 148-151).  Never an FMA: the LUT is a step function
// of this sum's top bits, and a fused g*g+old lands in the neighbouring bucket once in ~2^20 updates.
__device__ __forceinline__ float acc_after(float old, float grad) { return __fadd_rn(old, __fmul_rn(grad, grad)); }

// optimizer.rs calculate_update given the accumulator value *after* adding g^2
__device__ __forceinline__ float opt_step(uint32_t optimizer, float grad, float new_acc, const float *__restrict__ lut, float lr, float mpt)
{
    if (optimizer == OPT_LUT) { // optimizer.rs:147-156
        uint32_t key = __float_as_uint(new_acc) >> 20;
        return __fmul_rn(grad, __ldg(lut + key));
    }
    if (optimizer == OPT_FLEX) { // optimizer.rs:76-89
        float u = __fmul_rn(__fmul_rn(grad, lr), powf(new_acc, mpt));
        return (isnan(u) || isinf(u)) ? 0.0f : u;
    }
    return __fmul_rn(grad, lr); // optimizer.rs:35-37
}

// ---------------------------------------------------------------------------------------------
// k_learn<T, VEC>: T threads per example, VEC floats per memory transaction.
// Block = 256 threads = 256/T groups.  Dynamic smem = groups * p.group_smem_bytes.
// Group smem layout (floats): C[F*Fk] | d[n_cap*k] | val[n_cap] | hash[n_cap] | field[n_cap] | fstart[F+1] | red[16] | lrh[lr_cap]
// ---------------------------------------------------------------------------------------------
template <int T, int VEC, int MINB>
__global__ void __launch_bounds__(256, MINB) k_learn(const LearnParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GROUPS = 256 / T;
    constexpr int NW = T / 32;
    const int gib = threadIdx.x / T;      // group in block
    const int tg = threadIdx.x % T;       // thread in group
    const int lane = threadIdx.x & 31;
    const int wg = tg >> 5;               // warp in group

    float *C = reinterpret_cast<float *>(smem_raw + (size_t)gib * p.group_smem_bytes);
    const uint32_t F = p.F, k = p.k, Fk = p.Fk, cpr = p.cpr, ncap = p.n_cap;
    float *d = C + (size_t)F * Fk;
    float *val = d + (size_t)ncap * k;
    uint32_t *hash = reinterpret_cast<uint32_t *>(val + ncap);
    uint32_t *field = hash + ncap;
    uint32_t *fstart = field + ncap;
    float *red = reinterpret_cast<float *>(fstart + F + 1);
    uint32_t *lrh = reinterpret_cast<uint32_t *>(red + 16);

    const float *__restrict__ W = p.ffm_w;

    uint32_t n_groups = gridDim.x * GROUPS;
    if (p.max_groups && p.max_groups < n_groups) n_groups = p.max_groups;
    const uint32_t gid = blockIdx.x * GROUPS + gib;
    if (gid >= n_groups) return; // whole groups leave; the named barriers below are per group

    const uint32_t n_total = p.n_examples_dev ? *p.n_examples_dev : p.n_examples;
    if (gid == 0 && tg == 0 && p.stat_examples && n_total) atomicAdd(p.stat_examples, (unsigned long long)n_total);
    for (uint32_t ex = gid; ex < n_total; ex += n_groups) {
        const ExMeta m = p.meta[ex];
        const uint32_t n = m.ffm_cnt, nlr = m.lr_cnt;
        const uint4 *__restrict__ fe = p.ffm_ent + m.ffm_begin;
        const uint4 *__restrict__ le = p.lr_ent + m.lr_begin;
        if (n > ncap) { // uniform over the group
            if (tg == 0) { atomicOr(p.err_flag, 1u); p.preds[m.out_index] = __int_as_float(0x7fc00000); }
            continue;
        }
        float part = 0.0f;
        bool overlap = false;
        const uint32_t row = p.phase ? m.out_index - p.io.row_base : 0;
        float *xrow = p.phase == 1 ? p.io.X + (size_t)row * p.io.ldx : nullptr;
        const float *dxr = p.phase == 2 ? p.io.dX + (size_t)row * p.io.ldx : nullptr;

        const bool lr_staged = nlr <= p.lr_cap;
        if (lr_staged) for (uint32_t i = tg; i < nlr; i += T) lrh[i] = __ldg(&le[i].x); // visible after the next group_sync
        if (F > 0) {
            // ---- stage the example's feature list ------------------------------------------------
            for (uint32_t i = tg; i < n; i += T) {
                uint4 e = __ldg(fe + i);
                hash[i] = e.x; val[i] = __uint_as_float(e.y); field[i] = e.z;
            }
            // fstart[f] = first entry whose field >= f (entries are sorted by field; feature_buffer.rs:314-335)
            for (uint32_t f = tg; f <= F; f += T) {
                uint32_t i = 0;
                while (i < n && __ldg(&fe[i].z) < f) i++;
                fstart[f] = i;
            }
            group_sync<T>(gib);

            // ---- gather: C[z][:] = sum_{e in field z} v_e * W[h_e : h_e + Fk]  (block_ffm.rs:165-217) ----
            const uint32_t total = F * cpr;
            for (uint32_t idx0 = tg; idx0 < total; idx0 += 4 * T) {
                float w[4][VEC];
                uint32_t zz[4], cc[4], e0[4], e1[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    uint32_t idx = idx0 + u * T;
                    e0[u] = e1[u] = 0;
                    if (idx < total) {
                        zz[u] = fdiv(idx, p.div_cpr);
                        cc[u] = idx - zz[u] * cpr;
                        e0[u] = fstart[zz[u]];
                        e1[u] = fstart[zz[u] + 1];
                        if (e1[u] > e0[u]) ldcg_vec<VEC>(W + hash[e0[u]] + cc[u] * VEC, w[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    uint32_t idx = idx0 + u * T;
                    if (idx >= total) continue;
                    const uint32_t z = zz[u], x0 = cc[u] * VEC;
                    float acc[VEC];
#pragma unroll
                    for (int j = 0; j < VEC; j++) acc[j] = 0.0f;
                    for (uint32_t e = e0[u]; e < e1[u]; e++) {
                        float wv[VEC];
                        if (e == e0[u]) {
#pragma unroll
                            for (int j = 0; j < VEC; j++) wv[j] = w[u][j];
                        } else {
                            ldcg_vec<VEC>(W + hash[e] + x0, wv);
                        }
                        const float v = val[e];
                        const bool own_chunk = p.kv && x0 >= z * k && x0 < z * k + k;
#pragma unroll
                        for (int j = 0; j < VEC; j++) {
                            // first feature assigns w*v, the rest accumulate (separate roundings as in the reference)
                            float t = __fmul_rn(wv[j], v);
                            acc[j] = (e == e0[u]) ? t : __fadd_rn(acc[j], t);
                            if (p.kv) { if (own_chunk) d[e * k + (x0 - z * k) + j] = wv[j]; }
                            else { uint32_t x = x0 + j; if (x >= z * k && x < z * k + k) d[e * k + (x - z * k)] = wv[j]; } // own-field block of feature e
                        }
                    }
#pragma unroll
                    for (int j = 0; j < VEC; j++) C[z * Fk + x0 + j] = acc[j];
                }
            }
            group_sync<T>(gib);

            // Windows of two features of one example may overlap (the mask aligns rows to next_pow2(k) floats
            // only, feature_buffer.rs:142-148, and equal hashes can meet across fields).  The reference then
            // applies the per-slot accumulator updates in feature order (block_ffm.rs:269-287); remember
            // whether this example needs that ordering.
            if (p.update) {
                for (uint32_t a = tg; a < n; a += T)
                    for (uint32_t b = a + 1; b < n; b++) {
                        const uint32_t ha = hash[a], hb = hash[b];
                        const uint32_t diff = ha > hb ? ha - hb : hb - ha;
                        if (diff < Fk) overlap = true;
                    }
            }

            // ---- forward: FFM outputs through the triangle (block_ffm.rs:219-261, block_misc.rs:862-884) ----
            // sum over z<f of 2*out[f][z] + out[f][f], with out[f][z] = 0.5*sum_k C[f][zk..]*C[z][fk..]
            const uint32_t FF = F * F;
            for (uint32_t idx = tg; p.phase != 2 && idx < FF; idx += T) {
                const uint32_t f = fdiv(idx, p.div_F), z = idx - f * F;
                if (z < f) {
                    // 2 * out[f][z] = sum_q w_f[z][q] * (v * contra) with separate roundings (block_ffm.rs:246-257)
                    const float *a = C + f * Fk + z * k, *b = C + z * Fk + f * k;
                    float s = 0.0f;
                    for (uint32_t q = 0; q < k; q++) s = __fadd_rn(s, __fmul_rn(a[q], b[q]));
                    if (p.phase == 1) xrow[p.io.n_lr_out + tri_index(f, z)] = s;
                    else part += s;
                } else if (z == f) {
                    const uint32_t b0 = fstart[f], b1 = fstart[f + 1];
                    if (p.phase == 1 && b1 - b0 <= 1) xrow[p.io.n_lr_out + tri_index(f, f)] = 0.0f;
                    if (b1 - b0 > 1) { // intra-field pairs of a multi-valued field; a lone feature contributes exactly 0
                        const float *cf = C + f * Fk + f * k;
                        float s = 0.0f;
                        for (uint32_t e = b0; e < b1; e++) {
                            const float v = val[e];
                            for (uint32_t q = 0; q < k; q++) {
                                float wq = d[e * k + q];
                                float g_ = __fmul_rn(v, __fsub_rn(cf[q], __fmul_rn(wq, v))); // block_ffm.rs:238-243
                                s = __fadd_rn(s, __fmul_rn(wq, g_));
                            }
                        }
                        if (p.phase == 1) xrow[p.io.n_lr_out + tri_index(f, f)] = __fmul_rn(s, 0.5f);
                        else part += 0.5f * s;
                    }
                }
            }
        }

        // ---- LR forward (block_lr.rs:28-47) -------------------------------------------------------
        if (p.phase == 0) {
            for (uint32_t i = tg; i < nlr; i += T) {
                uint4 e = __ldg(le + i);
                float2 cell = __ldcg(p.lr + e.x);
                part += __fmul_rn(cell.x, __uint_as_float(e.y));
            }
        } else if (p.phase == 1) {
            // one output per combo, features added in buffer order (block_lr.rs:38-45); entries arrive in combo order
            for (uint32_t c = tg; c < p.io.n_lr_out; c += T) {
                float comb = 0.0f;
                for (uint32_t i = 0; i < nlr; i++) {
                    const uint4 e = __ldg(le + i);
                    if (e.z == c) comb = __fadd_rn(comb, __fmul_rn(__ldcg(p.lr + e.x).x, __uint_as_float(e.y)));
                }
                xrow[c] = comb;
            }
            if (tg == 0) { p.io.row_label[row] = m.label; p.io.row_importance[row] = m.importance; p.io.row_out_index[row] = m.out_index; }
            group_sync<T>(gib); // C / staging are reused by the next example
            continue;
        }

        // ---- reduce over the group -----------------------------------------------------------------
        float wsum = warp_sum(part);
        overlap = __any_sync(0xffffffffu, overlap);
        if (NW > 1) {
            if (lane == 0) { red[wg] = wsum; red[8 + wg] = overlap ? 1.0f : 0.0f; }
            group_sync<T>(gib);
            wsum = 0.0f;
            overlap = false;
#pragma unroll
            for (int w_ = 0; w_ < NW; w_++) { wsum += red[w_]; overlap = overlap || (red[8 + w_] != 0.0f); }
        }

        if (p.exact_order) {
            // Parity mode (one example in flight): the sigmoid sums its inputs left to right over the tape
            // [LR combo outputs..., triangle outputs...] (graph.rs:251-284, block_loss_functions.rs:116-120).
            // One thread redoes the forward in exactly that order, so the prediction and the gradient are bit-exact with
            // the reference; AdagradLUT (a step function of the accumulator's top bits) then never lands in another bucket.
            if (tg == 0) {
                float ws = 0.0f, comb = 0.0f;
                uint32_t cur = 0xffffffffu;
                for (uint32_t i = 0; i < nlr; i++) { // block_lr.rs:38-45: out[combo] += w * v, entries arrive in combo order
                    const uint4 e = __ldg(le + i);
                    if (e.z != cur) { if (cur != 0xffffffffu) ws = __fadd_rn(ws, comb); comb = 0.0f; cur = e.z; }
                    comb = __fadd_rn(comb, __fmul_rn(__ldcg(p.lr + e.x).x, __uint_as_float(e.y)));
                }
                if (cur != 0xffffffffu) ws = __fadd_rn(ws, comb);
                // FFM outputs through the triangle, term by term in tape order (block_misc.rs:871-881):
                // for field f: 2*out[f][z] (z < f) then out[f][f], where out[f][z] accumulates, feature by feature of
                // field f, 0.5 * sum_q w_e[z][q] * (v_e * (contra[f][z][q] - [z==f] w_e[f][q] v_e))  (block_ffm.rs:219-261).
                // Raw weights come straight from the table, contra sums from shared memory: faithful for any example.
                for (uint32_t f = 0; f < F; f++) {
                    const uint32_t b0 = fstart[f], b1 = fstart[f + 1];
                    for (uint32_t z = 0; z <= f; z++) {
                        float o = 0.0f;
                        for (uint32_t e = b0; e < b1; e++) {
                            const float v = val[e];
                            const float *wrow = W + hash[e] + z * k;
                            float corr = 0.0f;
                            for (uint32_t q = 0; q < k; q++) {
                                const float wq = __ldcg(wrow + q);
                                float cz = C[z * Fk + f * k + q];
                                if (z == f) cz = __fsub_rn(cz, __fmul_rn(wq, v));
                                corr = __fadd_rn(corr, __fmul_rn(wq, __fmul_rn(v, cz)));
                            }
                            o = __fadd_rn(o, __fmul_rn(corr, 0.5f));
                        }
                        ws = __fadd_rn(ws, z < f ? __fmul_rn(o, 2.0f) : o);
                    }
                }
                red[0] = ws;
            }
            group_sync<T>(gib);
            wsum = red[0];
        }

        // ---- sigmoid + logloss gradient (block_loss_functions.rs:105-153) ---------------------------
        float pr, g;
        if (p.phase == 2) g = __ldg(p.io.dy + row); // the head's sigmoid already produced the prediction and the gradient
        else {
            if (isnan(wsum)) { pr = logistic(0.0f); g = 0.0f; }
            else if (wsum < -50.0f) { pr = logistic(-50.0f); g = 0.0f; }
            else if (wsum > 50.0f) { pr = logistic(50.0f); g = 0.0f; }
            else { pr = logistic(wsum); g = __fmul_rn(-__fsub_rn(m.label, pr), m.importance); }
            if (tg == 0) p.preds[m.out_index] = pr;
        }
        // d_out[f][z]: the sigmoid gradient itself, or with a dense head what its backward pass left for that pair
        auto gpair = [&](uint32_t f, uint32_t z) -> float { return p.phase == 2 ? __ldg(dxr + p.io.n_lr_out + tri_index(f, z)) : g; };

        // regressor.rs:366-370: update && importance != 0; a zero gradient changes nothing
        const bool do_update = p.update && m.importance != 0.0f && g != 0.0f;
        if (do_update) {
            // ---- FFM update (block_ffm.rs:265-288): every d_out[f][z] equals g (triangle backward mirrors it) ----
            if (F > 0) {
                auto update_chunk = [&](uint32_t e, uint32_t c) {
                    const uint32_t f = field[e], h = hash[e], x0 = c * VEC;
                    const float v = val[e];
                    float grad[VEC], gg[VEC], old[VEC];
                    if (p.kv) {
                        // the whole chunk lies in the block towards one field zc
                        const uint32_t zc = fdiv(x0, p.div_k), q0 = x0 - zc * k;
                        const bool own = (zc == f);
                        // a lone feature's own-field block: the gradient is exactly 0, the reference's update a no-op
                        if (own && fstart[f + 1] - fstart[f] == 1) return;
                        const float *cp = C + zc * Fk + f * k + q0;
                        const float gz = gpair(f, zc);
                        bool any = false;
#pragma unroll
                        for (int j = 0; j < VEC; j++) {
                            float cz = cp[j];
                            if (own) cz = __fsub_rn(cz, __fmul_rn(d[e * k + q0 + j], v));
                            grad[j] = __fmul_rn(gz, __fmul_rn(v, cz));
                            gg[j] = __fmul_rn(grad[j], grad[j]);
                            any = any || grad[j] != 0.0f;
                        }
                        if (!any) return; // partner field absent
                    } else {
#pragma unroll
                        for (int j = 0; j < VEC; j++) {
                            const uint32_t x = x0 + j;
                            const uint32_t z = fdiv(x, p.div_k), q = x - z * k;
                            float cz = C[z * Fk + f * k + q];
                            // separate roundings, never an FMA: for a lone feature C[f][f-block] IS fl(w*v), so the
                            // self-interaction must cancel to exactly 0 like the reference's (block_ffm.rs:238-240);
                            // AdaGrad with a zero initial accumulator turns any residue into a full-size step.
                            if (z == f) cz = __fsub_rn(cz, __fmul_rn(d[e * k + q], v));
                            grad[j] = __fmul_rn(gpair(f, z), __fmul_rn(v, cz));
                            gg[j] = __fmul_rn(grad[j], grad[j]);
                        }
                    }
                    float upd[VEC];
                    if (p.optimizer == OPT_SGD) {
#pragma unroll
                        for (int j = 0; j < VEC; j++) upd[j] = -(grad[j] * p.ffm_lr);
                    } else {
                        atom_add_vec<VEC>(p.ffm_acc + h + x0, gg, old);
#pragma unroll
                        for (int j = 0; j < VEC; j++) upd[j] = -opt_step(p.optimizer, grad[j], acc_after(old[j], grad[j]), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    }
                    red_add_vec<VEC>(p.ffm_w + h + x0, upd);
                };
                if (!overlap && p.simple_update) {
                    const uint32_t total = n * cpr;
                    for (uint32_t idx = tg; idx < total; idx += T) {
                        const uint32_t e = fdiv(idx, p.div_cpr);
                        update_chunk(e, idx - e * cpr);
                    }
                } else if (!overlap) {
                    // UB independent chunks per thread and round: all accumulator atomics of a round are in flight
                    // together (one L2 round trip per round instead of one per chunk), then the weight reductions.
                    constexpr int UB = 4;
                    const uint32_t total = n * cpr;
                    // gradient of one chunk, recomputed from shared memory whenever needed (cheaper than holding it in
                    // registers across the atomic round trip: registers decide how many examples an SM keeps in flight)
                    auto chunk_grad = [&](uint32_t idx, float (&gr)[VEC], uint32_t &address) -> bool {
                        const uint32_t e = fdiv(idx, p.div_cpr), c = idx - e * cpr;
                        const uint32_t f = field[e], x0 = c * VEC;
                        const float v = val[e];
                        address = hash[e] + x0;
                        bool any = false;
#pragma unroll
                        for (int j = 0; j < VEC; j++) {
                            const uint32_t x = x0 + j;
                            const uint32_t z = fdiv(x, p.div_k), q = x - z * k;
                            float cz = C[z * Fk + f * k + q];
                            if (z == f) cz = __fsub_rn(cz, __fmul_rn(d[e * k + q], v));
                            gr[j] = __fmul_rn(gpair(f, z), __fmul_rn(v, cz));
                            any = any || (gr[j] != 0.0f);
                        }
                        return any; // an all-zero gradient (own block of a lone feature, absent field) changes nothing
                    };
                    for (uint32_t idx0 = tg; idx0 < total; idx0 += UB * T) {
                        float old[UB][VEC];
                        bool on[UB];
#pragma unroll
                        for (int u = 0; u < UB; u++) {
                            const uint32_t idx = idx0 + u * T;
                            on[u] = false;
                            if (idx >= total) continue;
                            float gr[VEC];
                            uint32_t address;
                            on[u] = chunk_grad(idx, gr, address);
                            if (on[u] && p.optimizer != OPT_SGD) {
                                float gg[VEC];
#pragma unroll
                                for (int j = 0; j < VEC; j++) gg[j] = __fmul_rn(gr[j], gr[j]);
                                atom_add_vec<VEC>(p.ffm_acc + address, gg, old[u]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < UB; u++) {
                            if (!on[u]) continue;
                            float gr[VEC], upd[VEC];
                            uint32_t address;
                            chunk_grad(idx0 + u * T, gr, address);
#pragma unroll
                            for (int j = 0; j < VEC; j++) {
                                if (p.optimizer == OPT_SGD) upd[j] = -(gr[j] * p.ffm_lr);
                                else upd[j] = -opt_step(p.optimizer, gr[j], acc_after(old[u][j], gr[j]), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                            }
                            red_add_vec<VEC>(p.ffm_w + address, upd);
                        }
                    }
                } else {
                    // ordered: a feature's accumulator atomics have returned (their results were consumed above)
                    // before the next feature's are issued
                    for (uint32_t e = 0; e < n; e++) {
                        for (uint32_t c = tg; c < cpr; c += T) update_chunk(e, c);
                        group_sync<T>(gib);
                    }
                }
            }
            // ---- LR update (block_lr.rs:135-151); duplicates of one hash inside an example are applied
            //      in buffer order by the first occurrence's thread (pinned by regressor.rs:629-656) ----
            for (uint32_t i = tg; i < nlr; i += T) {
                const uint4 e = __ldg(le + i);
                bool owner = true;
                if (lr_staged) { for (uint32_t j = 0; j < i; j++) if (lrh[j] == e.x) { owner = false; break; } }
                else { for (uint32_t j = 0; j < i; j++) if (__ldg(&le[j].x) == e.x) { owner = false; break; } }
                if (!owner) continue;
                float *cell = reinterpret_cast<float *>(p.lr + e.x);
                for (uint32_t j = i; j < nlr; j++) {
                    if (lr_staged && j != i && lrh[j] != e.x) continue;
                    const uint4 ej = (j == i) ? e : __ldg(le + j);
                    if (ej.x != e.x) continue;
                    const float grad = (p.phase == 2 ? __ldg(dxr + ej.z) : g) * __uint_as_float(ej.y);
                    float upd;
                    if (p.optimizer == OPT_SGD) upd = grad * p.lr_lr;
                    else {
                        const float gg = __fmul_rn(grad, grad);
                        const float old = atomicAdd(cell + 1, gg);
                        upd = opt_step(p.optimizer, grad, acc_after(old, grad), p.lut_lr, p.lr_lr, p.lr_mpt);
                    }
                    atomicAdd(cell, -upd);
                }
            }
        }
        // Parity mode: the next example must see this one's weight reductions (REDG is fire-and-forget; without the
        // fence its gather can overtake them).  Hogwild mode tolerates that staleness by design.
        if (p.exact_order) { if (p.sys_scope) __threadfence_system(); else __threadfence(); }
        group_sync<T>(gib); // C / meta are reused by the next example
    }
}


// ---------------------------------------------------------------------------------------------
// k_learn_fixed<G, NCH, NLR, OPTK>: the fused fast path for the cache's in-place encoding.
//
// The parser stores a namespace with one feature of weight 1.0 directly in its header slot
// (parser.rs:396-404) -- the case the reference itself special-cases ("value == 1.0", block_ffm.rs:978,
// SPEED.md).  When every field is one namespace and k % 4 == 0, a group of G lanes (a warp, or half a warp for narrow
// models) takes one raw record and does translate + forward + backward + update with the latent rows in registers:
//   lane j (+G t) owns the 16-byte chunk c of row e:  W[h_e + 4c .. 4c+4)  = w_e towards field z = 4c/k
//   its partner chunk  w_z towards field e  is fetched through a padded shared-memory transpose,
//   dot(mine, partner) is both the forward term and (times g) the gradient of my chunk,
//   so a lane issues one LDG.128, one ATOMG.128 and one REDG.128 per chunk and nothing else touches HBM.
// Records that do not fit (a referenced namespace holds several features or weights) are appended to
// a leftover list and go through k_translate + k_learn afterwards.
// ---------------------------------------------------------------------------------------------
struct FixedParams {
    float2 *lr; float *ffm_w; float *ffm_acc; const float *lut_lr; const float *lut_ffm;
    const uint32_t *records; const uint32_t *rec_off; uint32_t off_base, fixed_len;
    uint32_t ex_begin, n_examples; // this launch handles records [ex_begin, ex_begin + n_examples)
    uint32_t F, k, cpr;            // cpr = F*k/4 chunks per row
    FastDiv div_cpr, div_k4;       // by cpr, by k/4
    const uint32_t *field_ns;      // [F] the namespace of each field
    uint32_t n_combos; const uint32_t *combo_off, *combo_ns; const float *combo_weight; uint32_t add_constant;
    uint32_t lr_mask, ffm_mask;
    uint32_t optimizer; float lr_lr, lr_mpt, ffm_lr, ffm_mpt;
    int update;
    float *preds;
    uint32_t *leftover_idx, *leftover_cnt;
    uint32_t max_groups;
    uint32_t rec_smem_floats;      // F * (cpr + 1) * 4: the transpose area of one record
    int sys_scope;                 // sharded tables: parity-mode fences at system scope
    int bias_combine;              // 1: a group sums the bias gradients of FIXED_BIAS_PERIOD records before it applies them (mature models only)
};

// Block = FIXED_WARPS independent warps, one record per G-lane group and round; nothing is exchanged between groups.
//
// LR updates (block_lr.rs:135-151) step from the accumulator value read at gather time -- the cell is loaded as one
// float2 {w, acc} anyway -- so they are two fire-and-forget reductions and no warp waits for an atomic's return value.
// The constant feature is special: EVERY record hits its cell (feature_buffer.rs:270-276).  Same-address atomics
// serialise in L2 (~4 ns each on B200, tools/hotrow_microbench.cu: one bias update per record capped an early version at
// ~180 M records/s), and so do same-address LOADS: 5e8 ld.cg per second of that one cell were the cap of every later
// version until the cell moved into a register (profiles/r01_c2_fixed_indep_top_stalls.txt).  Each group therefore keeps
// the cell in a register, sums the bias gradients of FIXED_BIAS_PERIOD consecutive records and applies them as one update,
//   acc += sum g_i^2 ;  w -= (sum g_i) * LUT[acc],   re-reading the cell then (during the concurrency ramp: every record).
// History: two block-level schemes -- __syncthreads() every round with every warp scanning the others' pairs, then a
// deferred mbarrier hand-over -- left the warps waiting for the slowest one of their block for 26 % / 45 % of the
// stall samples (profiles/r01_c2_fixed_ncu_full.txt, profiles/r01_c2_fixed_mbar_top_stalls.txt).
constexpr int FIXED_WARPS = 16;
constexpr int FIXED_LR_MAX = 64;       // LR entries per record the fast path supports (two per lane)
constexpr int FIXED_BIAS_PERIOD = 16;  // records whose bias updates one group combines

__device__ __forceinline__ void red_add_f32(float *addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }

// one LR cell: acc += G2 ; w -= step(G, acc_seen + G2)
__device__ __forceinline__ void fixed_lr_apply(const FixedParams &p, uint32_t optimizer, uint32_t h, float G, float G2, float acc_seen)
{
    float *cell = reinterpret_cast<float *>(p.lr + h);
    float upd;
    if (optimizer == OPT_SGD) upd = __fmul_rn(G, p.lr_lr);
    else {
        red_add_f32(cell + 1, G2);
        upd = opt_step(optimizer, G, __fadd_rn(acc_seen, G2), p.lut_lr, p.lr_lr, p.lr_mpt);
    }
    red_add_f32(cell, -upd);
}

__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// Parity mode of k_learn_fixed: the triangle outputs added to `ws` in tape order, from a group's shared rows (kept out of line:
// it runs for one record in flight only and must not cost the throughput path registers)
__device__ __noinline__ float tape_order_ffm_sum(float ws, const float4 *S, uint32_t F, uint32_t k4, uint32_t row_stride)
{
    for (uint32_t f = 1; f < F; f++)        // 2 * out[f][z], z < f (block_misc.rs:871-881); the diagonal of a lone feature is 0
        for (uint32_t z = 0; z < f; z++) {
            float corr = 0.0f;              // block_ffm.rs:246-257: correction += w * (v * contra), v = 1.0
            for (uint32_t q = 0; q < k4; q++) {
                const float4 a = S[f * row_stride + z * k4 + q], b = S[z * row_stride + f * k4 + q];
                corr = __fadd_rn(corr, __fmul_rn(a.x, b.x)); corr = __fadd_rn(corr, __fmul_rn(a.y, b.y));
                corr = __fadd_rn(corr, __fmul_rn(a.z, b.z)); corr = __fadd_rn(corr, __fmul_rn(a.w, b.w));
            }
            ws = __fadd_rn(ws, corr);
        }
    return ws;
}

// G   : lanes per record (32, or 16 = two records per warp and round when F <= 16, F*F*k/4 <= 64 chunks and at most 16
//       LR entries: the per-record overhead -- translate, sigmoid, reductions -- is then shared by two records and
//       every wait for L2 covers two records);
// NCH : 16-byte chunks per lane (ceil(F*F*k/4 / G));  NLR : LR entries per lane (ceil((combos + constant) / G));
// OPTK: the optimizer as a compile-time constant (OPT_LUT: no powf code, no optimizer branches) or -1 = p.optimizer
// PARITY: the instantiation for ONE record in flight (p.max_groups == 1, hogwild_max_inflight = 1): the same translate / gather /
//       gradient / optimizer code plus (a) the sigmoid input summed in the reference's tape order, (b) fences between records,
//       (c) LR duplicates inside a record ordered through an atomic's return value.  Kept out of the throughput instantiation
//       because it costs registers there (the fast path is bound by how many records an SM keeps in flight).
template <int G, int NCH, int NLR, int OPTK, bool PARITY>
__global__ void __launch_bounds__(FIXED_WARPS * 32, (NCH >= 3 ? 2 : 3)) k_learn_fixed(const FixedParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = FIXED_WARPS, RPW = 32 / G; // records per warp and round
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sl = lane & (G - 1), sg = lane / G; // lane within the record's group, group within the warp
    const int g0 = sg * G;                        // first lane of my group
    const uint32_t gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << g0);
    float4 *S = reinterpret_cast<float4 *>(smem_raw) + (size_t)(wib * RPW + sg) * (p.rec_smem_floats / 4);
    const uint32_t optimizer = OPTK < 0 ? p.optimizer : (uint32_t)OPTK;
    const uint32_t F = p.F, k = p.k, cpr = p.cpr, row_stride = cpr + 1, k4 = k >> 2;
    const uint32_t n_chunks = F * cpr;
    uint32_t n_groups = gridDim.x * NW * RPW;     // records in flight
    if (p.max_groups && p.max_groups < n_groups) n_groups = p.max_groups;
    const uint32_t wg0 = (blockIdx.x * NW + wib) * RPW; // this warp's first group
    if (wg0 >= n_groups) return;                  // the ramp may leave the last block partly idle
    const bool group_on = wg0 + sg < n_groups;

    // static per-lane geometry: which chunk(s) I own and where my partner lives
    uint32_t my_e[NCH], my_c4[NCH], part_off[NCH], my_off[NCH];
    bool act[NCH], diag[NCH];
#pragma unroll
    for (int t = 0; t < NCH; t++) {
        const uint32_t j = sl + G * t;
        act[t] = j < n_chunks;
        const uint32_t e = act[t] ? fdiv(j, p.div_cpr) : 0, c = act[t] ? j - e * cpr : 0;
        const uint32_t z = fdiv(c, p.div_k4), q4 = c - z * k4; // chunk c = quarter q4 of the block towards field z
        my_e[t] = g0 + e; my_c4[t] = 4 * c;
        my_off[t] = e * row_stride + c;
        part_off[t] = z * row_stride + e * k4 + q4;              // row z, its block towards field e, same quarter
        diag[t] = (z == e);
    }
    const uint32_t my_field_ns = sl < F ? __ldg(p.field_ns + sl) : 0;
    // static per-lane LR entries: lane i (+G) owns combo i; the entry after the last combo is the constant feature
    uint32_t c_o0[NLR], c_len[NLR]; // the combo's namespaces are combo_ns[c_o0 .. c_o0 + c_len); 0 = no entry, 0xffffffff = constant
    float c_w[NLR];
    bool bias_lane = false;
#pragma unroll
    for (int r = 0; r < NLR; r++) {
        const uint32_t i = sl + G * r;
        c_o0[r] = 0; c_len[r] = 0; c_w[r] = 0.0f;
        if (i < p.n_combos) { c_o0[r] = __ldg(p.combo_off + i); c_len[r] = __ldg(p.combo_off + i + 1) - c_o0[r]; c_w[r] = __ldg(p.combo_weight + i); }
        else if (i == p.n_combos && p.add_constant) { c_len[r] = 0xffffffffu; c_w[r] = 1.0f; bias_lane = true; }
    }
    const uint32_t bias_h = 11650396u & p.lr_mask; // feature_buffer.rs:270-276
    // the combined bias update of this group (see above) and the cell as of the last refresh
    const uint32_t bias_period = !p.update ? 0xffffffffu : (p.max_groups || !p.bias_combine) ? 1u : (uint32_t)FIXED_BIAS_PERIOD; // predict: the cell never changes
    float bias_G = 0.0f, bias_G2 = 0.0f;
    float2 bias_cell = bias_lane ? __ldcg(p.lr + bias_h) : make_float2(0.f, 0.f);
    uint32_t bias_n = 0;

    // The record stream comes from HBM: every lane that reads a header slot prefetches its word of the NEXT record of
    // its group into L1 (no register, nothing waits), so the loads below hit L1.  With an offset array the offset of
    // the record after next is prefetched as well.
    auto rec_of = [&](uint32_t e_) -> const uint32_t * {
        return p.records + (p.rec_off ? (size_t)(__ldg(p.rec_off + e_) - p.off_base) : (size_t)e_ * p.fixed_len);
    };
    if (group_on && wg0 + sg < p.n_examples && sl < F) prefetch_l1(rec_of(p.ex_begin + wg0 + sg) + 3 + my_field_ns);

    for (uint32_t base0 = wg0; base0 < p.n_examples; base0 += n_groups) { // group g takes records g, g + n_groups, ...
        const uint32_t base = base0 + sg, ex = p.ex_begin + base;
        bool live = group_on && base < p.n_examples;
        if (group_on && base + n_groups < p.n_examples && sl < F) prefetch_l1(rec_of(ex + n_groups) + 3 + my_field_ns);
        if (p.rec_off && group_on && base + 2 * n_groups < p.n_examples && sl == 0) prefetch_l1(p.rec_off + ex + 2 * n_groups);

        const uint32_t *rec = live ? rec_of(ex) : p.records;
        // ---- translate (feature_buffer.rs:178-338) for in-place slots; anything else -> leftover ----
        const uint32_t slot = (live && sl < F) ? __ldg(rec + 3 + my_field_ns) : 0x80000000u;
        bool bad = (slot & 0x80000000u) && slot != 0x80000000u;
        uint32_t lr_h[NLR];
        bool lr_ok[NLR];
#pragma unroll
        for (int r = 0; r < NLR; r++) {
            lr_h[r] = bias_h; lr_ok[r] = live && c_len[r] != 0;
            if (live && c_len[r] - 1u < 0xfffffffeu) { // a combo: chain its namespaces' hashes (feature_buffer.rs:239-251)
                uint32_t h = 0;
                for (uint32_t o = 0; o < c_len[r]; o++) {
                    const uint32_t s_ = __ldg(rec + 3 + __ldg(p.combo_ns + c_o0[r] + o));
                    if (s_ & 0x80000000u) { lr_ok[r] = false; if (s_ != 0x80000000u) bad = true; }
                    h = o ? ((h * 16777619u) ^ s_) : s_;
                }
                lr_h[r] = h & p.lr_mask;
            }
        }
        if (__ballot_sync(0xffffffffu, bad) & gmask) { // my group's record does not fit the fast path
            if (sl == 0) { const uint32_t at = atomicAdd(p.leftover_cnt, 1u); p.leftover_idx[at] = ex; }
            live = false;
        }
        // Two LR features of ONE record in the same cell (a hash collision between combos, ~1e-4 of the records at -b 18): the
        // reference applies them one after the other (block_lr.rs:140-150), so the second must see the first one's accumulator.
        // Those lanes take their accumulator from an atomic's return value instead of the gathered cell.
        bool lr_dup[NLR];
#pragma unroll
        for (int r = 0; r < NLR; r++) {
            lr_dup[r] = false;
            if (PARITY) {
                const uint32_t key = (live && lr_ok[r] && c_len[r] != 0xffffffffu) ? lr_h[r] : (0x80000000u | (uint32_t)lane);
                lr_dup[r] = __popc(__match_any_sync(0xffffffffu, key) & gmask) > 1;
            }
        }

        // ---- gather: one 128-bit load per chunk ----
        float4 v[NCH];
        uint32_t hbase[NCH]; bool pres[NCH];
#pragma unroll
        for (int t = 0; t < NCH; t++) {
            const uint32_t s_ = __shfl_sync(0xffffffffu, slot, my_e[t]);
            pres[t] = live && act[t] && s_ != 0x80000000u;
            hbase[t] = (s_ & p.ffm_mask) + my_c4[t];
            // own-field chunks are never read: a single-valued field has no interaction with itself (block_ffm.rs:236-244)
            v[t] = (pres[t] && !diag[t]) ? __ldcg(reinterpret_cast<const float4 *>(p.ffm_w + hbase[t])) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float2 lrw[NLR];
#pragma unroll
        for (int r = 0; r < NLR; r++) {
            // the bias cell is read by EVERY record: 5e8 loads/s of one address queue up in its L2 slice (they were the
            // longest stall of the kernel, profiles/r01_c2_fixed_indep_top_stalls.txt), so a group keeps the cell in a
            // register and re-reads it when it applies its combined bias update
            lr_ok[r] = lr_ok[r] && live;
            if (c_len[r] == 0xffffffffu) lrw[r] = bias_cell;
            else lrw[r] = lr_ok[r] ? __ldcg(p.lr + lr_h[r]) : make_float2(0.f, 0.f);
        }
        const float label = live ? (float)__ldg(rec + 1) : 0.0f, importance = live ? __uint_as_float(__ldg(rec + 2)) : 0.0f;
        __syncwarp(); // the previous round's partner reads are done
#pragma unroll
        for (int t = 0; t < NCH; t++) if (act[t]) S[my_off[t]] = v[t];
        __syncwarp();

        // ---- forward ----
        float part = 0.0f;
        float4 pv[NCH];
#pragma unroll
        for (int t = 0; t < NCH; t++) {
            pv[t] = act[t] ? S[part_off[t]] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (act[t] && !diag[t]) {
                float sd = __fmul_rn(v[t].x, pv[t].x);
                sd = __fadd_rn(sd, __fmul_rn(v[t].y, pv[t].y));
                sd = __fadd_rn(sd, __fmul_rn(v[t].z, pv[t].z));
                sd = __fadd_rn(sd, __fmul_rn(v[t].w, pv[t].w));
                part += sd;
            }
        }
        part *= 0.5f; // every unordered field pair is seen from both sides; the triangle keeps 2*out[f][z], z < f
#pragma unroll
        for (int r = 0; r < NLR; r++) if (lr_ok[r]) part += __fmul_rn(lrw[r].x, c_w[r]);
        float wsum = part;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        if (PARITY) {
            // One record in flight = the reference's sequential loop: the sigmoid input is summed in the reference's own order,
            // left to right over the tape [LR combo outputs..., triangle outputs...] (graph.rs:251-284,
            // block_loss_functions.rs:116-120), every lane redoing the same serial sum from the group's shared rows.  The
            // prediction and the gradient are then BIT-exact, and AdagradLUT -- a step function of the accumulator's top bits
            // -- never lands in another bucket.  Everything else (translate, gather, gradients, atomics) is the code above.
            float ws = 0.0f;
#pragma unroll
            for (int r = 0; r < NLR; r++)
                for (int i = 0; i < G; i++) { // combo i + G r: out[combo] = w * value, 0.0 when the combo has no feature (block_lr.rs:38-45)
                    const float t = __shfl_sync(0xffffffffu, lr_ok[r] ? __fmul_rn(lrw[r].x, c_w[r]) : 0.0f, g0 + i);
                    ws = __fadd_rn(ws, t);
                }
            __syncwarp();
            ws = tape_order_ffm_sum(ws, S, F, k4, row_stride);
            wsum = ws;
        }

        float pr, g;
        if (isnan(wsum)) { pr = logistic(0.0f); g = 0.0f; }
        else if (wsum < -50.0f) { pr = logistic(-50.0f); g = 0.0f; }
        else if (wsum > 50.0f) { pr = logistic(50.0f); g = 0.0f; }
        else { pr = logistic(wsum); g = __fmul_rn(-__fsub_rn(label, pr), importance); }
        if (live && sl == 0) p.preds[ex] = pr;

        if (live && p.update && importance != 0.0f && g != 0.0f) {
            // ---- FFM update: grad of my chunk = g * partner chunk (values are 1.0); diagonal chunks get exactly 0 ----
#pragma unroll
            for (int t = 0; t < NCH; t++) {
                if (!pres[t] || diag[t]) continue;
                const float gx = __fmul_rn(g, pv[t].x), gy = __fmul_rn(g, pv[t].y), gz = __fmul_rn(g, pv[t].z), gw = __fmul_rn(g, pv[t].w);
                if (gx == 0.0f && gy == 0.0f && gz == 0.0f && gw == 0.0f) continue; // partner field absent
                float4 upd;
                if (optimizer == OPT_SGD) {
                    upd = make_float4(-__fmul_rn(gx, p.ffm_lr), -__fmul_rn(gy, p.ffm_lr), -__fmul_rn(gz, p.ffm_lr), -__fmul_rn(gw, p.ffm_lr));
                } else {
                    const float4 gg = make_float4(__fmul_rn(gx, gx), __fmul_rn(gy, gy), __fmul_rn(gz, gz), __fmul_rn(gw, gw));
                    const float4 old = atomicAdd(reinterpret_cast<float4 *>(p.ffm_acc + hbase[t]), gg);
                    upd.x = -opt_step(optimizer, gx, __fadd_rn(old.x, gg.x), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.y = -opt_step(optimizer, gy, __fadd_rn(old.y, gg.y), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.z = -opt_step(optimizer, gz, __fadd_rn(old.z, gg.z), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.w = -opt_step(optimizer, gw, __fadd_rn(old.w, gg.w), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                }
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p.ffm_w + hbase[t]), "f"(upd.x), "f"(upd.y), "f"(upd.z), "f"(upd.w) : "memory");
            }
            // ---- LR update; the bias is summed over bias_period records first ----
#pragma unroll
            for (int r = 0; r < NLR; r++) {
                if (!lr_ok[r]) continue;
                const float gl = __fmul_rn(g, c_w[r]);
                if (c_len[r] == 0xffffffffu) { bias_G = __fadd_rn(bias_G, gl); bias_G2 = __fadd_rn(bias_G2, __fmul_rn(gl, gl)); }
                else if (gl != 0.0f) {
                    float acc_seen = lrw[r].y;
                    if (PARITY && lr_dup[r] && optimizer != OPT_SGD) { // ordered by the L2: each duplicate steps from what the previous one left
                        const float gg = __fmul_rn(gl, gl);
                        acc_seen = atomicAdd(reinterpret_cast<float *>(p.lr + lr_h[r]) + 1, gg);
                        red_add_f32(reinterpret_cast<float *>(p.lr + lr_h[r]), -opt_step(optimizer, gl, __fadd_rn(acc_seen, gg), p.lut_lr, p.lr_lr, p.lr_mpt));
                    } else fixed_lr_apply(p, optimizer, lr_h[r], gl, __fmul_rn(gl, gl), acc_seen);
                }
            }
        }
        // One record in flight (hogwild_max_inflight = 1, the setting of the per-example parity tests): REDG is fire-and-forget,
        // so every lane's reductions are fenced and the group re-converges before the next record's gather may read them.
        const bool one_in_flight = PARITY;
        if (++bias_n >= bias_period) {
            if (bias_lane) {
                if (bias_G != 0.0f) fixed_lr_apply(p, optimizer, bias_h, bias_G, bias_G2, bias_cell.y);
                if (one_in_flight) { if (p.sys_scope) __threadfence_system(); else __threadfence(); }
                bias_cell = __ldcg(p.lr + bias_h); // consumed a round later
            }
            bias_G = 0.0f; bias_G2 = 0.0f; bias_n = 0;
        }
        if (one_in_flight) { if (p.sys_scope) __threadfence_system(); else __threadfence(); __syncwarp(); }
    }
    if (bias_lane && bias_G != 0.0f) fixed_lr_apply(p, optimizer, bias_h, bias_G, bias_G2, bias_cell.y);
}

// ---------------------------------------------------------------------------------------------
// k_learn_fixed_cta<UB>: the fused fast path for WIDE models (F*F*k/4 > 128 chunks, e.g. 39 fields x k=8).
// Same contract as k_learn_fixed (raw records in the cache's in-place encoding, one namespace per field,
// k % 4 == 0; anything else -> leftover list), but one 256-thread BLOCK per record and the F rows staged in shared memory:
//   * the record's header slots are read by F threads (prefetched one record ahead), hashes masked in registers;
//   * gather: every 16-byte chunk goes HBM -> shared memory with cp.async.cg (LDGSTS, no register staging), all of a
//     thread's chunks in flight at once, one wait per record;
//   * forward: sum over z<f of <C[f][z-block], C[z][f-block]> with conflict-free LDS.128, shuffle + shared reduction;
//   * update: chunk (e,c) takes its partner chunk C[z][e-block] from shared memory, grad = g * partner,
//     ATOMG.128 on the accumulators (UB chunks in flight per thread), LUT, REDG.128 on the weights; own-field chunks skipped.
// Shared memory per record: F*F*k*4 B (48.7 KB for 39 x 8) -> 4 records in flight per SM.
// ---------------------------------------------------------------------------------------------
struct FixedCtaParams {
    float2 *lr; float *ffm_w; float *ffm_acc; const float *lut_lr; const float *lut_ffm;
    const uint32_t *records; const uint32_t *rec_off; uint32_t off_base, fixed_len;
    uint32_t ex_begin, n_examples;
    uint32_t F, k, Fk, cpr;        // cpr = Fk / 4
    FastDiv div_cpr, div_k4, div_F;
    const uint32_t *field_ns;
    uint32_t n_combos; const uint32_t *combo_off, *combo_ns; const float *combo_weight; uint32_t add_constant;
    uint32_t lr_mask, ffm_mask;
    uint32_t optimizer; float lr_lr, lr_mpt, ffm_lr, ffm_mpt;
    int update;
    float *preds;
    uint32_t *leftover_idx, *leftover_cnt;
    uint32_t max_groups;
    HeadIO io;                     // dense-head models only (PHASE 1 / 2)
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

template <int UB, int PHASE>
__global__ void __launch_bounds__(256, 4) k_learn_fixed_cta(const FixedCtaParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t F = p.F, k = p.k, Fk = p.Fk, cpr = p.cpr, k4 = k >> 2;
    float *C = reinterpret_cast<float *>(smem_raw);
    uint32_t *slots2 = reinterpret_cast<uint32_t *>(C + (size_t)F * Fk); // [2][F], double-buffered by record parity: a thread
    float *red = reinterpret_cast<float *>(slots2 + 2 * F);               // may still read record i's slots while another stores i+1's
    const uint32_t n_chunks = F * cpr;
    const uint32_t n_lr = p.n_combos + (p.add_constant ? 1u : 0u); // <= 256 (host checks)
    uint32_t n_blocks = gridDim.x;
    if (p.max_groups && p.max_groups < n_blocks) n_blocks = p.max_groups;
    if (blockIdx.x >= n_blocks) return;
    const uint32_t my_field_ns = tid < F ? __ldg(p.field_ns + tid) : 0;
    auto rec_ptr = [&](uint32_t ex) { return p.records + (p.rec_off ? (size_t)(p.rec_off[ex] - p.off_base) : (size_t)ex * p.fixed_len); };

    uint32_t ex = p.ex_begin + blockIdx.x;
    const uint32_t ex_end = p.ex_begin + p.n_examples;
    uint32_t slot_next = (ex < ex_end && tid < F) ? __ldg(rec_ptr(ex) + 3 + my_field_ns) : 0x80000000u;

    for (uint32_t it = 0; ex < ex_end; ex += n_blocks, it++) {
        const uint32_t *rec = rec_ptr(ex);
        uint32_t *slots = slots2 + (it & 1) * F;
        // ---- translate (feature_buffer.rs:178-338), in-place slots only ----
        const uint32_t slot = slot_next;
        bool bad = tid < F && (slot & 0x80000000u) && slot != 0x80000000u;
        if (tid < F) slots[tid] = slot;
        uint32_t lr_h = 0; float lr_v = 0.0f; bool lr_ok = false;
        if (tid < p.n_combos) {
            const uint32_t o0 = __ldg(p.combo_off + tid), o1 = __ldg(p.combo_off + tid + 1);
            uint32_t h = 0; bool ok = true;
            for (uint32_t o = o0; o < o1; o++) {
                const uint32_t sl = __ldg(rec + 3 + __ldg(p.combo_ns + o));
                if (sl & 0x80000000u) { ok = false; if (sl != 0x80000000u) bad = true; }
                h = (o == o0) ? sl : ((h * 16777619u) ^ sl);
            }
            lr_ok = ok; lr_h = h & p.lr_mask; lr_v = __ldg(p.combo_weight + tid);
        } else if (tid == p.n_combos && p.add_constant) { lr_ok = true; lr_h = 11650396u & p.lr_mask; lr_v = 1.0f; }
        const float label = (float)__ldg(rec + 1), importance = __uint_as_float(__ldg(rec + 2));
        const int any_bad = __syncthreads_or(bad ? 1 : 0); // also publishes slots[] and closes the previous record's use of C
        // prefetch the next record's slots: their cache lines are warm when the next iteration needs them
        {
            const uint32_t nx = ex + n_blocks;
            slot_next = (nx < ex_end && tid < F) ? __ldg(rec_ptr(nx) + 3 + my_field_ns) : 0x80000000u;
        }
        if (any_bad) {
            // PHASE 2 walks the same records as PHASE 1: the leftover list already holds this one
            if (PHASE != 2 && tid == 0) { const uint32_t at = atomicAdd(p.leftover_cnt, 1u); p.leftover_idx[at] = ex; }
            continue; // uniform
        }
        const uint32_t row = ex - p.io.row_base;
        float g_row = 0.0f;
        if (PHASE == 2) { g_row = __ldg(p.io.dy + row); if (g_row == 0.0f) continue; } // uniform: nothing to update

        // ---- gather: HBM -> shared memory, 16 B per cp.async, everything in flight at once ----
        for (uint32_t idx = tid; idx < n_chunks; idx += 256) {
            const uint32_t e = fdiv(idx, p.div_cpr), c = idx - e * cpr;
            const uint32_t sl = slots[e];
            float *dst = C + e * Fk + 4 * c;
            if (sl != 0x80000000u) cp_async16(dst, p.ffm_w + (sl & p.ffm_mask) + 4 * c);
            else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float lr_w = (PHASE != 2 && lr_ok) ? __ldcg(p.lr + lr_h).x : 0.0f;
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        float g;
        if (PHASE != 2) {
        // ---- forward ----
        float part = lr_ok ? __fmul_rn(lr_w, lr_v) : 0.0f;
        float *xr = PHASE == 1 ? p.io.X + (size_t)row * p.io.ldx : nullptr;
        if (PHASE == 1) {
            if (tid < n_lr) xr[tid] = part; // one feature of value 1.0 per combo: out[combo] = w * combo weight (block_lr.rs:38-45)
            if (tid == 0) { p.io.row_label[row] = label; p.io.row_importance[row] = importance; p.io.row_out_index[row] = ex; }
        }
        const uint32_t FF = F * F;
        for (uint32_t idx = tid; idx < FF; idx += 256) {
            const uint32_t f = fdiv(idx, p.div_F), z = idx - f * F;
            if (PHASE == 1 && z == f) xr[p.io.n_lr_out + tri_index(f, f)] = 0.0f; // a lone feature has no intra-field term
            if (z < f) {
                const float4 *a = reinterpret_cast<const float4 *>(C + f * Fk + z * k), *b = reinterpret_cast<const float4 *>(C + z * Fk + f * k);
                float sd = 0.0f;
                for (uint32_t q = 0; q < k4; q++) {
                    const float4 x = a[q], y = b[q];
                    sd = __fadd_rn(sd, __fmul_rn(x.x, y.x)); sd = __fadd_rn(sd, __fmul_rn(x.y, y.y));
                    sd = __fadd_rn(sd, __fmul_rn(x.z, y.z)); sd = __fadd_rn(sd, __fmul_rn(x.w, y.w));
                }
                if (PHASE == 1) xr[p.io.n_lr_out + tri_index(f, z)] = sd; // = 2 * out[f][z] (block_misc.rs:871-881)
                else part += sd;
            }
        }
        if (PHASE == 1) continue; // uniform; the next iteration's barrier protects C
        float wsum = warp_sum(part);
        if (lane == 0) red[warp] = wsum;
        __syncthreads();
        wsum = 0.0f;
#pragma unroll
        for (int w_ = 0; w_ < 8; w_++) wsum += red[w_];

        float pr;
        if (isnan(wsum)) { pr = logistic(0.0f); g = 0.0f; }
        else if (wsum < -50.0f) { pr = logistic(-50.0f); g = 0.0f; }
        else if (wsum > 50.0f) { pr = logistic(50.0f); g = 0.0f; }
        else { pr = logistic(wsum); g = __fmul_rn(-__fsub_rn(label, pr), importance); }
        if (tid == 0) p.preds[ex] = pr;
        if (!(p.update && importance != 0.0f && g != 0.0f)) continue; // uniform; the next iteration's barrier protects C
        } else {
            g = g_row;
        }
        const float *dxr = PHASE == 2 ? p.io.dX + (size_t)row * p.io.ldx : nullptr;

        // ---- update: UB chunks per thread and round ----
        for (uint32_t idx0 = tid; idx0 < n_chunks; idx0 += UB * 256) {
            float4 gr[UB], old[UB];
            uint32_t addr[UB];
            bool on[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const uint32_t idx = idx0 + u * 256;
                on[u] = false;
                if (idx >= n_chunks) continue;
                const uint32_t e = fdiv(idx, p.div_cpr), c = idx - e * cpr;
                const uint32_t z = fdiv(c, p.div_k4), q4 = c - z * k4;
                const uint32_t sl = slots[e];
                if (z == e || sl == 0x80000000u) continue; // own-field chunk: exactly zero gradient; absent field: no row
                const float4 pv = *reinterpret_cast<const float4 *>(C + z * Fk + e * k + 4 * q4);
                const float gz = PHASE == 2 ? __ldg(dxr + p.io.n_lr_out + tri_index(e, z)) : g; // d_out[e][z] (block_misc.rs:823-832)
                gr[u] = make_float4(__fmul_rn(gz, pv.x), __fmul_rn(gz, pv.y), __fmul_rn(gz, pv.z), __fmul_rn(gz, pv.w));
                if (gr[u].x == 0.0f && gr[u].y == 0.0f && gr[u].z == 0.0f && gr[u].w == 0.0f) continue; // partner absent
                on[u] = true;
                addr[u] = (sl & p.ffm_mask) + 4 * c;
                if (p.optimizer != OPT_SGD)
                    old[u] = atomicAdd(reinterpret_cast<float4 *>(p.ffm_acc + addr[u]),
                                       make_float4(__fmul_rn(gr[u].x, gr[u].x), __fmul_rn(gr[u].y, gr[u].y), __fmul_rn(gr[u].z, gr[u].z), __fmul_rn(gr[u].w, gr[u].w)));
            }
#pragma unroll
            for (int u = 0; u < UB; u++) {
                if (!on[u]) continue;
                float4 upd;
                if (p.optimizer == OPT_SGD) upd = make_float4(-__fmul_rn(gr[u].x, p.ffm_lr), -__fmul_rn(gr[u].y, p.ffm_lr), -__fmul_rn(gr[u].z, p.ffm_lr), -__fmul_rn(gr[u].w, p.ffm_lr));
                else {
                    upd.x = -opt_step(p.optimizer, gr[u].x, acc_after(old[u].x, gr[u].x), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.y = -opt_step(p.optimizer, gr[u].y, acc_after(old[u].y, gr[u].y), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.z = -opt_step(p.optimizer, gr[u].z, acc_after(old[u].z, gr[u].z), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.w = -opt_step(p.optimizer, gr[u].w, acc_after(old[u].w, gr[u].w), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                }
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p.ffm_w + addr[u]), "f"(upd.x), "f"(upd.y), "f"(upd.z), "f"(upd.w) : "memory");
            }
        }
        // ---- LR update (block_lr.rs:135-151) ----
        if (lr_ok) {
            float *cell = reinterpret_cast<float *>(p.lr + lr_h);
            const float grad = __fmul_rn(PHASE == 2 ? __ldg(dxr + tid) : g, lr_v);
            float upd;
            if (p.optimizer == OPT_SGD) upd = __fmul_rn(grad, p.lr_lr);
            else {
                const float old = atomicAdd(cell + 1, __fmul_rn(grad, grad));
                upd = opt_step(p.optimizer, grad, acc_after(old, grad), p.lut_lr, p.lr_lr, p.lr_mpt);
            }
            atomicAdd(cell, -upd);
        }
        // one record in flight (per-example parity tests): this record's reductions are visible before the barrier at the
        // top of the next iteration lets anybody gather
        if (p.max_groups == 1) __threadfence();
    }
}

// ---------------------------------------------------------------------------------------------
// k_learn_rows<PHASE>: the fused path for WIDE models, rebuilt around bulk asynchronous copies (one per row and array)
// instead of one LDGSTS / ATOMG / REDG per 16 bytes.  Same contract as k_learn_fixed_cta (raw records in the cache's
// in-place encoding, one namespace per field, k % 4 == 0; anything else -> leftover list), one 256-thread block per record:
//   * gather: thread e < F issues cp.async.bulk global -> shared for row e of the WEIGHTS and of the ACCUMULATORS
//     (two pieces each: the row without its own-field block, which a single-valued field never uses, block_ffm.rs:236-244);
//     an mbarrier counts the bytes, nobody spends an instruction or a register on the 97 KB in flight;
//   * forward + update are organised by field PAIR: the thread that holds chunk (e, z-block) also holds its partner
//     (z, e-block), so dot(a, b) is the pair's forward term, g * b and g * a are the two gradients, and both chunks are
//     overwritten IN PLACE with the weight step  -g * LUT[acc + g^2]  and the accumulator increment  g^2  -- no thread ever
//     reads what another one rewrites, so there is no barrier between gather and scatter except the sigmoid's reduction;
//   * scatter: thread e issues cp.reduce.async.bulk shared -> global (.add.f32) for row e of both arrays: the L2 adds
//     the whole row, fire and forget.  The optimizer step uses the accumulator snapshot that came with the gather
//     (acc += g^2 is still exact -- nothing is lost -- but concurrent records that hit one slot all step from the value
//     they gathered; with ONE record in flight this is the reference's sequential arithmetic, optimizer.rs:147-156);
//   * the next record's header slots are prefetched while this one computes; two blocks per SM alternate between
//     waiting for their rows and computing.
// The same code serves a hash-range-sharded table (fwgpu_create_sharded): a remote row is pulled with the same bulk copy
// over NVLink and its gradient row is pushed back as ONE bulk reduction that the owner's L2 applies.
// Shared memory per block: 2 * F*F*k*4 B (+ 8 KB LUT): 105 KB for 39 fields x k = 8 -> two blocks per SM.
// ---------------------------------------------------------------------------------------------
struct RowsParams {
    float2 *lr; float *ffm_w; float *ffm_acc; const float *lut_lr; const float *lut_ffm;
    const uint32_t *records; const uint32_t *rec_off; uint32_t off_base, fixed_len;
    uint32_t ex_begin, n_examples;
    uint32_t F, k, Fk, k4;
    uint32_t lpp, n_units;         // lanes per field pair (k/4 when that is 1, 2 or 4, else 1); units = pairs * lpp
    const uint32_t *field_ns;
    uint32_t n_combos; const uint32_t *combo_off, *combo_ns; const float *combo_weight; uint32_t add_constant;
    uint32_t lr_mask, ffm_mask;
    uint32_t optimizer; float lr_lr, lr_mpt, ffm_lr, ffm_mpt;
    int update;
    float *preds;
    uint32_t *leftover_idx, *leftover_cnt;
    uint32_t max_groups;
    HeadIO io;                     // dense-head models only (PHASE 1 / 2)
    // PUSH mode (one model over the GPUs of one box, fwgpu_create_sharded): rows are pulled from their owner's HBM with the same
    // bulk copies; instead of updating in place the block writes the record's RAW gradient rows, each with a 16-byte header
    // {row base, field, -, -}, and pushes row e as ONE bulk store into the inbox of the rank that owns the row
    // (inbox[owner][half][source rank][slot]); the owner applies AdaGrad from its own accumulators (k_apply_inbox).
    unsigned char *inbox;          // one virtual range: rank r's inbox at inbox + r * inbox_rank_stride
    unsigned long long inbox_rank_stride, inbox_src_off; // bytes; inbox_src_off = offset of (half, this rank)'s sub-ring
    uint32_t *push_cnt;            // [world] entries this rank has pushed to each owner during the current chunk (local)
    uint32_t owner_shift, world;   // owner = row base >> owner_shift (>= 32: everything on rank 0), clamped to world - 1
    int sys_scope;                 // sharded tables: parity-mode fences at system scope
};
constexpr int ROWS_MAXU = 8;
constexpr uint32_t ROWS_HDR = 4;   // floats of header in front of each shared-memory row in PUSH mode       // pair units per thread held in registers: n_units <= 8 * 256

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
// shared -> global, element-wise f32 add performed by the L2 that owns the line (local HBM or, for a sharded table, the peer's)
__device__ __forceinline__ void bulk_reduce_add(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

// optimizer.rs calculate_update with the LUT in shared memory
__device__ __forceinline__ float opt_step_s(uint32_t optimizer, float grad, float new_acc, const float *lut_s, float lr, float mpt)
{
    if (optimizer == OPT_LUT) return __fmul_rn(grad, lut_s[__float_as_uint(new_acc) >> 20]);
    if (optimizer == OPT_FLEX) { const float u = __fmul_rn(__fmul_rn(grad, lr), powf(new_acc, mpt)); return (isnan(u) || isinf(u)) ? 0.0f : u; }
    return __fmul_rn(grad, lr);
}

// OPTK: the optimizer as a compile-time constant (OPT_LUT, the reference's default under --adaptive: no powf code) or -1 = optimizer
template <int PHASE, int OPTK, bool PUSH>
__global__ void __launch_bounds__(256, (PHASE == 1 || PUSH ? 4 : 2)) k_learn_rows(const RowsParams p)
{
    const uint32_t optimizer = OPTK < 0 ? p.optimizer : (uint32_t)OPTK;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t F = p.F, k = p.k, Fk = p.Fk, k4 = p.k4, lpp = p.lpp;
    const uint32_t RS = Fk + (PUSH ? ROWS_HDR : 0u);           // shared-memory row stride in floats
    const bool writes = PHASE != 1 && p.update != 0;           // this launch updates the tables (PUSH: sends gradients)
    const bool has_acc = !PUSH && writes && optimizer != OPT_SGD; // accumulator rows travel with the weight rows
    const bool use_lut = !PUSH && writes && optimizer == OPT_LUT;
    float *W = reinterpret_cast<float *>(smem_raw) + (PUSH ? ROWS_HDR : 0u); // row e at W + e * RS (its header right in front of it)
    float *A = W + (size_t)F * RS;
    float *lut_s = A + (has_acc ? (size_t)F * RS : 0);
    // the record's header slots, double-buffered by record parity: a block's fast threads store record i+1's slots (before the
    // barrier that publishes them) while slow ones may still read record i's
    uint32_t *slots2 = reinterpret_cast<uint32_t *>(lut_s + (use_lut ? 2048 : 0));
    const uint32_t slots_stride = (F + 3) & ~3u;
    float *red = reinterpret_cast<float *>(slots2 + 2 * slots_stride);
    uint64_t *bar = reinterpret_cast<uint64_t *>(red + 8);
    float *terms = reinterpret_cast<float *>(bar + 2); // one-record-in-flight mode only: the sigmoid's inputs in tape order

    uint32_t n_blocks = gridDim.x;
    if (p.max_groups && p.max_groups < n_blocks) n_blocks = p.max_groups;
    if (blockIdx.x >= n_blocks) return;
    const bool one_in_flight = p.max_groups == 1; // per-example parity runs: every write is complete before the next record gathers

    // One bulk copy per row and array (the whole row: its own-field block rides along unused -- 2.5 % of the bytes -- so that a
    // row is ONE copy, and is zeroed before it is reduced back).  The copies are issued by 8 warps at once: op o = row (o % F)
    // of array (o / F) belongs to lane o / 8 of warp o % 8, because a warp issues its lanes' bulk copies one after the other.
    const uint32_t n_ops = F * (has_acc ? 2u : 1u);
    const uint32_t my_op = lane * 8 + warp;
    const bool op_on = my_op < n_ops;
    const uint32_t op_row = op_on ? (my_op >= F ? my_op - F : my_op) : 0;
    const bool op_acc = my_op >= F;
    float *const op_smem = (op_acc ? A : W) + (size_t)op_row * RS;
    float *const op_table = op_acc ? p.ffm_acc : p.ffm_w;
    if (tid == 0) mbar_init(bar, n_ops);          // one arrival per op and record
    if (use_lut) for (uint32_t i = tid; i < 2048; i += 256) lut_s[i] = __ldg(p.lut_ffm + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const uint32_t n_lr = p.n_combos + (p.add_constant ? 1u : 0u); // <= 256 (host checks); = the number of LR outputs
    // static geometry: unit u = tid + 256 j is lane (u % lpp) of field pair (u / lpp) = (e, z), e < z; it owns the 16-byte
    // quarters q = u % lpp, + lpp, ... of the chunk pair  a = row e, block towards z   and   b = row z, block towards e
    uint32_t offA[ROWS_MAXU], offB[ROWS_MAXU], tri[ROWS_MAXU], ez[ROWS_MAXU];
#pragma unroll
    for (int j = 0; j < ROWS_MAXU; j++) {
        const uint32_t u = tid + 256u * j;
        offA[j] = 0xffffffffu; offB[j] = 0; tri[j] = 0; ez[j] = 0;
        if (u < p.n_units) {
            const uint32_t pr = u / lpp, q0 = u - pr * lpp;
            uint32_t z = (uint32_t)((1.0f + sqrtf(1.0f + 8.0f * (float)pr)) * 0.5f); // pr = z (z - 1) / 2 + e
            while (z * (z - 1) / 2 > pr) z--;
            while ((z + 1) * z / 2 <= pr) z++;
            const uint32_t e = pr - z * (z - 1) / 2;
            offA[j] = e * RS + z * k + 4 * q0;
            offB[j] = z * RS + e * k + 4 * q0;
            tri[j] = n_lr + z * (z + 1) / 2 + e; // position on the tape / in the head's input (block_misc.rs:871-881)
            ez[j] = e | (z << 16);
        }
    }
    const uint32_t my_field_ns = tid < F ? __ldg(p.field_ns + tid) : 0;
    auto rec_ptr = [&](uint32_t ex) { return p.records + (p.rec_off ? (size_t)(p.rec_off[ex] - p.off_base) : (size_t)ex * p.fixed_len); };

    uint32_t ex = p.ex_begin + blockIdx.x;
    const uint32_t ex_end = p.ex_begin + p.n_examples;
    uint32_t slot_next = (ex < ex_end && tid < F) ? __ldg(rec_ptr(ex) + 3 + my_field_ns) : 0x80000000u;
    uint32_t parity = 0, rec_parity = 0;

    for (; ex < ex_end; ex += n_blocks, rec_parity ^= 1u) {
        const uint32_t *rec = rec_ptr(ex);
        uint32_t *slots = slots2 + rec_parity * slots_stride;
        // ---- translate (feature_buffer.rs:178-338), in-place slots only ----
        const uint32_t slot = slot_next;
        const bool absent = tid < F && slot == 0x80000000u;
        bool bad = tid < F && (slot & 0x80000000u) && !absent;
        uint32_t lr_h = 0; float lr_v = 0.0f; bool lr_ok = false;
        if (tid < p.n_combos) {
            const uint32_t o0 = __ldg(p.combo_off + tid), o1 = __ldg(p.combo_off + tid + 1);
            uint32_t h = 0; bool ok = true;
            for (uint32_t o = o0; o < o1; o++) {
                const uint32_t sl = __ldg(rec + 3 + __ldg(p.combo_ns + o));
                if (sl & 0x80000000u) { ok = false; if (sl != 0x80000000u) bad = true; }
                h = (o == o0) ? sl : ((h * 16777619u) ^ sl);
            }
            lr_ok = ok; lr_h = h & p.lr_mask; lr_v = __ldg(p.combo_weight + tid);
        } else if (tid == p.n_combos && p.add_constant) { lr_ok = true; lr_h = 11650396u & p.lr_mask; lr_v = 1.0f; }
        const float label = (float)__ldg(rec + 1), importance = __uint_as_float(__ldg(rec + 2));
        // the bulk reductions this thread issued for the previous record have finished READING shared memory
        if (op_on) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (tid < F) slots[tid] = slot;
        const int flags = __syncthreads_or((bad ? 1 : 0) | (absent ? 2 : 0)); // publishes slots[]; the buffers are free
        {
            const uint32_t nx = ex + n_blocks;
            slot_next = (nx < ex_end && tid < F) ? __ldg(rec_ptr(nx) + 3 + my_field_ns) : 0x80000000u;
        }
        if (flags & 1) {
            // PHASE 2 walks the same records as PHASE 1: the leftover list already holds this one
            if (PHASE != 2 && tid == 0) { const uint32_t at = atomicAdd(p.leftover_cnt, 1u); p.leftover_idx[at] = ex; }
            continue; // uniform
        }
        const uint32_t row = ex - p.io.row_base;
        float g = 0.0f;
        if (PHASE == 2) { g = __ldg(p.io.dy + row); if (g == 0.0f) continue; } // uniform: nothing to update

        // ---- gather: one bulk copy per row piece and array, completion counted on the mbarrier ----
        const uint32_t op_slot = op_on ? slots[op_row] : 0x80000000u;
        const bool op_present = op_slot != 0x80000000u;
        const uint32_t h_row = op_slot & p.ffm_mask;  // base of MY op's row
        unsigned char *push_dst = nullptr;
        if (op_on) {
            if (op_present) {
                mbar_arrive_expect_tx(bar, Fk * 4u);
                bulk_load(op_smem, op_table + h_row, Fk * 4u, bar);
            } else mbar_arrive(bar);
        }
        const float2 lr_cell = (PHASE != 2 || writes) && lr_ok ? __ldcg(p.lr + lr_h) : make_float2(0.f, 0.f);
        // Two rows of ONE record whose windows overlap (equal or neighbouring hashes: 2.7 % of the records at 39 fields and
        // ffm_bit_precision 24) share slots: the reference updates them feature after feature (block_ffm.rs:269-287), so the
        // second update of a shared slot sees the first one's accumulator.  Such a record takes its accumulators from atomics'
        // return values (update_with_atomics below) instead of the gathered snapshot.  Checked while the rows are in flight.
        bool overlap = false;
        if (!PUSH && writes) {
#pragma unroll
            for (int j = 0; j < ROWS_MAXU; j++) { // every field pair is somebody's unit: the F^2 / 2 comparisons cost two LDS each
                if (offA[j] == 0xffffffffu) continue;
                const uint32_t se = slots[ez[j] & 0xffffu], sz = slots[ez[j] >> 16];
                const uint32_t he = se & p.ffm_mask, hz = sz & p.ffm_mask, diff = he > hz ? he - hz : hz - he;
                overlap = overlap || (diff < Fk && se != 0x80000000u && sz != 0x80000000u);
            }
        }
        if (flags & 2) { // some field is absent: its row reads as zeros (no feature, no interaction)
            for (uint32_t e = 0; e < F; e++)
                if (slots[e] == 0x80000000u) {
                    for (uint32_t i = tid; i < Fk; i += 256) { W[(size_t)e * RS + i] = 0.0f; if (has_acc) A[(size_t)e * RS + i] = 0.0f; }
                }
            __syncthreads();
        }
        mbar_wait(bar, parity);
        parity ^= 1u;

        const float *dxr = PHASE == 2 ? p.io.dX + (size_t)row * p.io.ldx : nullptr;
        if (PHASE != 2) {
            // ---- forward: sum over field pairs of <a, b> (block_ffm.rs:219-261 through the triangle, block_misc.rs:871-881) ----
            float part = lr_ok ? __fmul_rn(lr_cell.x, lr_v) : 0.0f;
            float *xr = PHASE == 1 ? p.io.X + (size_t)row * p.io.ldx : nullptr;
            if (PHASE == 1) {
                if (tid < n_lr) xr[tid] = part; // one feature of value 1.0 per combo: out[combo] = w * combo weight (block_lr.rs:38-45)
                if (tid < F) xr[p.io.n_lr_out + tri_index(tid, tid)] = 0.0f; // a lone feature has no intra-field term
                if (tid == 0) { p.io.row_label[row] = label; p.io.row_importance[row] = importance; p.io.row_out_index[row] = ex; }
            }
            // <a, b> over my quarter(s), every product and sum rounded separately in the order of block_ffm.rs:246-257
            auto fold = [](float acc, const float4 &x, const float4 &y) {
                acc = __fadd_rn(acc, __fmul_rn(x.x, y.x)); acc = __fadd_rn(acc, __fmul_rn(x.y, y.y));
                acc = __fadd_rn(acc, __fmul_rn(x.z, y.z)); return __fadd_rn(acc, __fmul_rn(x.w, y.w));
            };
#pragma unroll
            for (int j = 0; j < ROWS_MAXU; j++) {
                const bool on = offA[j] != 0xffffffffu;
                float sd = 0.0f;
                if (one_in_flight && lpp > 1) {
                    // parity mode: the lanes of a pair chain their quarters in order, so the pair's term is rounded exactly like
                    // the reference's k-long running sum; the complete term ends up in the pair's last lane
                    const float4 x = on ? *reinterpret_cast<const float4 *>(W + offA[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 y = on ? *reinterpret_cast<const float4 *>(W + offB[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    sd = fold(0.0f, x, y);
                    for (uint32_t step = 1; step < lpp; step++) {
                        const float prev = __shfl_up_sync(0xffffffffu, sd, 1);
                        if ((lane & (lpp - 1)) == step) sd = fold(prev, x, y);
                    }
                    if (on && (lane & (lpp - 1)) == lpp - 1) { if (PHASE == 1) xr[tri[j]] = sd; else terms[tri[j]] = sd; }
                    continue;
                }
                if (on) {
                    for (uint32_t q = 0; q * lpp < k4; q++) // one iteration for k = 4, 8, 16
                        sd = fold(sd, *reinterpret_cast<const float4 *>(W + offA[j] + 4 * q * lpp), *reinterpret_cast<const float4 *>(W + offB[j] + 4 * q * lpp));
                }
                if (one_in_flight) { if (on) { if (PHASE == 1) xr[tri[j]] = sd; else terms[tri[j]] = sd; } continue; } // lpp == 1: already in order
                if (PHASE == 1) { // = 2 * out[z][e]: the lanes of one pair are neighbours in the warp
                    if (lpp >= 2) sd += __shfl_xor_sync(0xffffffffu, sd, 1);
                    if (lpp >= 4) sd += __shfl_xor_sync(0xffffffffu, sd, 2);
                    if (on && (lane & (lpp - 1)) == 0) xr[tri[j]] = sd;
                } else part += sd;
            }
            if (PHASE == 1) continue; // uniform; the next iteration's barrier protects the buffers
            float wsum;
            if (one_in_flight) {
                // One record in flight = the reference's sequential loop: the sigmoid input is summed in the reference's own
                // order, left to right over the tape [LR combo outputs..., triangle outputs...] (graph.rs:251-284,
                // block_loss_functions.rs:116-120), so prediction and gradient are BIT-exact and AdagradLUT never lands in another
                // bucket.  Gather, gradients, optimizer step and scatter are the code every other mode runs.
                if (tid < n_lr) terms[tid] = part;                                      // out[combo] = w * value, or 0.0 (block_lr.rs:38-45)
                if (tid < F) terms[n_lr + tri_index(tid, tid)] = 0.0f;                  // a lone feature has no intra-field term
                // (a sharded table in parity mode takes the atomics path too: an atomic's return value proves that the owner has
                //  performed it, which a bulk reduction into a peer's memory does not)
                overlap = __syncthreads_or(overlap ? 1 : 0) != 0 || (p.sys_scope && !PUSH);
                const uint32_t x_len = n_lr + F * (F + 1) / 2;
                wsum = 0.0f;
                for (uint32_t i = 0; i < x_len; i++) wsum = __fadd_rn(wsum, terms[i]);
            } else {
                wsum = warp_sum(part);
                if (lane == 0) red[warp] = wsum;
                overlap = __syncthreads_or(overlap ? 1 : 0) != 0;
                wsum = 0.0f;
#pragma unroll
                for (int w_ = 0; w_ < 8; w_++) wsum += red[w_];
            }

            float pr;
            if (isnan(wsum)) { pr = logistic(0.0f); g = 0.0f; }
            else if (wsum < -50.0f) { pr = logistic(-50.0f); g = 0.0f; }
            else if (wsum > 50.0f) { pr = logistic(50.0f); g = 0.0f; }
            else { pr = logistic(wsum); g = __fmul_rn(-__fsub_rn(label, pr), importance); }
            if (tid == 0) p.preds[ex] = pr;
            if (!(p.update && importance != 0.0f && g != 0.0f)) continue; // uniform; regressor.rs:366-370
        } else if (!PUSH) overlap = __syncthreads_or(overlap ? 1 : 0) != 0;

        // ---- LR accumulators first (block_lr.rs:135-151): the atomic's return value is needed only after the FFM update, so
        //      its round trip to the L2 is hidden; duplicates of one cell inside a record are ordered by the L2 ----
        float lr_grad = 0.0f, lr_gg = 0.0f, lr_old = 0.0f;
        if (lr_ok) {
            lr_grad = __fmul_rn(PHASE == 2 ? __ldg(dxr + tid) : g, lr_v);
            lr_gg = __fmul_rn(lr_grad, lr_grad);
            if (lr_grad != 0.0f && optimizer != OPT_SGD) lr_old = atomicAdd(reinterpret_cast<float *>(p.lr + lr_h) + 1, lr_gg);
        }

        if (overlap) {
            // ---- update_with_atomics: this record's rows share slots.  Chunk by chunk like the single-GPU kernels of round 1:
            //      ATOMG.128 on the accumulators (returns the old value), step, REDG.128 on the weights; nothing is written in
            //      place and nothing is bulk-reduced for this record. ----
            const uint32_t cpr = Fk >> 2;
            for (uint32_t idx = tid; idx < F * cpr; idx += 256) {
                const uint32_t e = idx / cpr, c = idx - e * cpr, z = c / k4, q4 = c - z * k4;
                const uint32_t sl = slots[e];
                if (z == e || sl == 0x80000000u) continue; // own-field chunk: exactly zero gradient; absent field: no row
                const float4 pv = *reinterpret_cast<const float4 *>(W + (size_t)z * RS + e * k + 4 * q4);
                const float gz = PHASE == 2 ? __ldg(dxr + n_lr + tri_index(e, z)) : g;
                const float4 gr = make_float4(__fmul_rn(gz, pv.x), __fmul_rn(gz, pv.y), __fmul_rn(gz, pv.z), __fmul_rn(gz, pv.w));
                if (gr.x == 0.0f && gr.y == 0.0f && gr.z == 0.0f && gr.w == 0.0f) continue; // partner absent
                const uint32_t addr = (sl & p.ffm_mask) + 4 * c;
                float4 upd;
                if (optimizer == OPT_SGD) upd = make_float4(-__fmul_rn(gr.x, p.ffm_lr), -__fmul_rn(gr.y, p.ffm_lr), -__fmul_rn(gr.z, p.ffm_lr), -__fmul_rn(gr.w, p.ffm_lr));
                else {
                    const float4 old = atomicAdd(reinterpret_cast<float4 *>(p.ffm_acc + addr), make_float4(__fmul_rn(gr.x, gr.x), __fmul_rn(gr.y, gr.y), __fmul_rn(gr.z, gr.z), __fmul_rn(gr.w, gr.w)));
                    upd.x = -opt_step(optimizer, gr.x, acc_after(old.x, gr.x), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.y = -opt_step(optimizer, gr.y, acc_after(old.y, gr.y), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.z = -opt_step(optimizer, gr.z, acc_after(old.z, gr.z), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.w = -opt_step(optimizer, gr.w, acc_after(old.w, gr.w), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                }
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p.ffm_w + addr), "f"(upd.x), "f"(upd.y), "f"(upd.z), "f"(upd.w) : "memory");
            }
        } else {
        // ---- update, in place: weights := -step, accumulators := g^2 (block_ffm.rs:265-288, optimizer.rs:147-156) ----
#pragma unroll
        for (int j = 0; j < ROWS_MAXU; j++) {
            if (offA[j] == 0xffffffffu) continue;
            const float gz = PHASE == 2 ? __ldg(dxr + tri[j]) : g; // d_out[e][z] = d_out[z][e] (block_misc.rs:823-832)
            for (uint32_t q = 0; q * lpp < k4; q++) {
                float4 *pa = reinterpret_cast<float4 *>(W + offA[j] + 4 * q * lpp), *pb = reinterpret_cast<float4 *>(W + offB[j] + 4 * q * lpp);
                const float4 a = *pa, b = *pb;
                // gradient of a slot = d_out * value * partner, value = 1.0 (block_ffm.rs:246-257, 269-287)
                const float ga[4] = {__fmul_rn(gz, b.x), __fmul_rn(gz, b.y), __fmul_rn(gz, b.z), __fmul_rn(gz, b.w)};
                const float gb[4] = {__fmul_rn(gz, a.x), __fmul_rn(gz, a.y), __fmul_rn(gz, a.z), __fmul_rn(gz, a.w)};
                float ua[4], ub[4];
                if (has_acc) {
                    float4 *qa = reinterpret_cast<float4 *>(A + offA[j] + 4 * q * lpp), *qb = reinterpret_cast<float4 *>(A + offB[j] + 4 * q * lpp);
                    const float4 ca = *qa, cb = *qb;
                    const float sa[4] = {ca.x, ca.y, ca.z, ca.w}, sb[4] = {cb.x, cb.y, cb.z, cb.w};
                    float g2a[4], g2b[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        g2a[c] = __fmul_rn(ga[c], ga[c]); g2b[c] = __fmul_rn(gb[c], gb[c]);
                        ua[c] = -opt_step_s(optimizer, ga[c], __fadd_rn(sa[c], g2a[c]), lut_s, p.ffm_lr, p.ffm_mpt);
                        ub[c] = -opt_step_s(optimizer, gb[c], __fadd_rn(sb[c], g2b[c]), lut_s, p.ffm_lr, p.ffm_mpt);
                    }
                    *qa = make_float4(g2a[0], g2a[1], g2a[2], g2a[3]);
                    *qb = make_float4(g2b[0], g2b[1], g2b[2], g2b[3]);
                } else if (PUSH) {
#pragma unroll
                    for (int c = 0; c < 4; c++) { ua[c] = ga[c]; ub[c] = gb[c]; } // the owner takes the step (k_apply_inbox)
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) { ua[c] = -__fmul_rn(ga[c], p.ffm_lr); ub[c] = -__fmul_rn(gb[c], p.ffm_lr); }
                }
                *pa = make_float4(ua[0], ua[1], ua[2], ua[3]);
                *pb = make_float4(ub[0], ub[1], ub[2], ub[3]);
            }
        }
        }
        // ---- LR weights: step from the accumulator the atomic returned ----
        if (lr_ok && lr_grad != 0.0f) {
            const float upd = optimizer == OPT_SGD ? __fmul_rn(lr_grad, p.lr_lr) : opt_step(optimizer, lr_grad, __fadd_rn(lr_old, lr_gg), p.lut_lr, p.lr_lr, p.lr_mpt);
            red_add_f32(reinterpret_cast<float *>(p.lr + lr_h), -upd);
        }
        // ---- scatter: the rewritten rows go back as bulk reductions, one per row piece and array ----
        if (PUSH) {
            // one inbox slot per present row, taken from this rank's counter for the owner (one atomic per owner and warp)
            if (op_on && op_present) {
                uint32_t owner = p.owner_shift >= 32 ? 0u : (h_row >> p.owner_shift);
                if (owner >= p.world) owner = p.world - 1; // the spill-over tail lives on the last rank
                const uint32_t peers = __match_any_sync(__activemask(), owner);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(p.push_cnt + owner, (uint32_t)__popc(peers));
                base = __shfl_sync(peers, base, leader);
                const uint32_t slot_i = base + __popc(peers & ((1u << lane) - 1u));
                uint32_t *hdr = reinterpret_cast<uint32_t *>(op_smem) - ROWS_HDR;
                hdr[0] = h_row; hdr[1] = op_row; hdr[2] = 0; hdr[3] = 0;
                push_dst = p.inbox + (size_t)owner * p.inbox_rank_stride + p.inbox_src_off + (size_t)slot_i * ((size_t)RS * 4);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (op_on && op_present) {
                const uint32_t src = smem_u32(op_smem - ROWS_HDR);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(push_dst), "r"(src), "r"(RS * 4u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            continue;
        }
        if (overlap) { if (one_in_flight) { if (p.sys_scope) __threadfence_system(); else __threadfence(); } continue; } // uniform: this record went through the atomics
        if (op_on && op_present) { // the own-field block of my row takes no update: it goes back as zeros
            for (uint32_t q = 0; q < k4; q++) *reinterpret_cast<float4 *>(op_smem + op_row * k + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // my shared-memory writes are visible to the copy engine
        __syncthreads();
        if (op_on && op_present) {
            bulk_reduce_add(op_table + h_row, op_smem, Fk * 4u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (one_in_flight) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // complete, not just read
        }
        if (one_in_flight) { if (p.sys_scope) __threadfence_system(); else __threadfence(); } // the LR reductions too, before the barrier at the top of the next iteration
    }
    if (op_on) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // shared memory must outlive the reductions that read it
}

// ---------------------------------------------------------------------------------------------
// k_apply_inbox: the OWNER's half of the sharded update.  Every entry of this rank's inbox is one gradient row
// {row base, field, -, -, F*k floats} that some rank (this one included) computed for a row this rank owns; a warp takes an
// entry and applies AdaGrad chunk by chunk exactly like the single-GPU kernels do (block_ffm.rs:265-288, optimizer.rs:147-156):
// ATOMG.128 on the local accumulators (returns the old value), LUT step, REDG.128 on the local weights -- all in this GPU's
// own L2; nothing but the 1.26 KB entry ever crossed NVLink.  The own-field block of an entry is skipped (zero gradient).
// ---------------------------------------------------------------------------------------------
struct ApplyParams {
    const unsigned char *inbox_half;  // this rank's inbox, current half: [world][cap] entries of entry_bytes
    const uint32_t *counts_all;       // [world][world] after the all-gather: counts_all[s * world + r] = entries rank s pushed to rank r
    uint32_t world, rank, cap, entry_bytes;
    uint32_t F, k, Fk;
    float *ffm_w, *ffm_acc; const float *lut_ffm;
    uint32_t optimizer; float ffm_lr, ffm_mpt;
};
constexpr int APPLY_UB = 3;
__global__ void __launch_bounds__(256, 4) k_apply_inbox(const ApplyParams p)
{
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t cnt[16], total = 0;
#pragma unroll
    for (uint32_t s_ = 0; s_ < 16; s_++) { cnt[s_] = s_ < p.world ? min(__ldcg(p.counts_all + s_ * p.world + p.rank), p.cap) : 0u; total += cnt[s_]; }
    const uint32_t cpr = p.Fk >> 2, k4 = p.k >> 2;
    for (uint32_t idx = warp; idx < total; idx += n_warps) {
        uint32_t s_ = 0, i = idx;
#pragma unroll
        for (uint32_t t = 0; t < 16; t++) if (s_ == t && i >= cnt[t]) { i -= cnt[t]; s_ = t + 1; }
        const unsigned char *ent = p.inbox_half + ((size_t)s_ * p.cap + i) * p.entry_bytes;
        const uint4 hdr = __ldcg(reinterpret_cast<const uint4 *>(ent));
        const uint32_t h_row = hdr.x, e = hdr.y;
        const float4 *gr4 = reinterpret_cast<const float4 *>(ent + 16);
        // APPLY_UB chunks per lane in flight: all gradient loads, then all accumulator atomics, then the steps and reductions,
        // so that a warp waits for one round trip to the L2 per 96 chunks instead of one per 32
        for (uint32_t c0 = 0; c0 < cpr; c0 += 32 * APPLY_UB) {
            float4 gr[APPLY_UB], old[APPLY_UB];
            bool on[APPLY_UB];
#pragma unroll
            for (int u = 0; u < APPLY_UB; u++) {
                const uint32_t c = c0 + 32 * u + lane;
                on[u] = c < cpr && c / k4 != e; // own-field block: exactly zero gradient (block_ffm.rs:236-244)
                gr[u] = on[u] ? __ldcs(gr4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < APPLY_UB; u++) {
                on[u] = on[u] && !(gr[u].x == 0.0f && gr[u].y == 0.0f && gr[u].z == 0.0f && gr[u].w == 0.0f); // partner field absent
                if (on[u] && p.optimizer != OPT_SGD)
                    old[u] = atomicAdd(reinterpret_cast<float4 *>(p.ffm_acc + h_row + 4 * (c0 + 32 * u + lane)),
                                       make_float4(__fmul_rn(gr[u].x, gr[u].x), __fmul_rn(gr[u].y, gr[u].y), __fmul_rn(gr[u].z, gr[u].z), __fmul_rn(gr[u].w, gr[u].w)));
            }
#pragma unroll
            for (int u = 0; u < APPLY_UB; u++) {
                if (!on[u]) continue;
                float4 upd;
                if (p.optimizer == OPT_SGD) upd = make_float4(-__fmul_rn(gr[u].x, p.ffm_lr), -__fmul_rn(gr[u].y, p.ffm_lr), -__fmul_rn(gr[u].z, p.ffm_lr), -__fmul_rn(gr[u].w, p.ffm_lr));
                else {
                    upd.x = -opt_step(p.optimizer, gr[u].x, acc_after(old[u].x, gr[u].x), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.y = -opt_step(p.optimizer, gr[u].y, acc_after(old[u].y, gr[u].y), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.z = -opt_step(p.optimizer, gr[u].z, acc_after(old[u].z, gr[u].z), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                    upd.w = -opt_step(p.optimizer, gr[u].w, acc_after(old[u].w, gr[u].w), p.lut_ffm, p.ffm_lr, p.ffm_mpt);
                }
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p.ffm_w + h_row + 4 * (c0 + 32 * u + lane)), "f"(upd.x), "f"(upd.y), "f"(upd.z), "f"(upd.w) : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Translate: raw record -> AoS feature lists (feature_buffer.rs:178-338), one thread per example.
// ---------------------------------------------------------------------------------------------
struct TranslateParams {
    const uint32_t *records;   // words
    const uint32_t *rec_off;   // [n+1] word offsets, or nullptr with fixed_len
    uint32_t off_base;         // subtracted from rec_off values (chunked uploads)
    uint32_t fixed_len;
    uint32_t n_examples;
    uint32_t n_namespaces;
    const uint8_t *ns_is_f32;
    uint32_t n_combos;
    const uint32_t *combo_off, *combo_ns;
    const float *combo_weight;
    uint32_t add_constant;
    uint32_t n_fields;
    const uint32_t *field_off, *field_ns;
    uint32_t lr_mask, ffm_mask, ffm_k;
    uint32_t lr_stride, ffm_stride; // entries per example slab
    ExMeta *meta;
    uint4 *lr_ent, *ffm_ent;
    uint32_t *err_flag; // bit1: slab overflow
    const uint32_t *ex_list;  // optional: translate only these examples (the fast kernel's leftovers) ...
    const uint32_t *ex_count; // ... their number, in device memory
};

// feature_reader! (feature_buffer.rs:48-108) for a primitive namespace: number of features and accessors
struct NsView { const uint32_t *rec; uint32_t tok, start, cnt; bool single, f32; };
__device__ __forceinline__ NsView ns_view(const uint32_t *rec, uint32_t ns, const uint8_t *is_f32)
{
    NsView v;
    v.rec = rec;
    v.tok = rec[3 + ns];
    v.f32 = is_f32 ? (is_f32[ns] != 0) : false;
    v.single = (v.tok & 0x80000000u) == 0;
    if (v.single) { v.start = 0; v.cnt = 1; }
    else {
        uint32_t s = (v.tok >> 16) & 0x3fffu, e = v.tok & 0xffffu;
        v.start = s;
        v.cnt = e > s ? (e - s + 1) / 2 : 0; // (start..end).step_by(2)
    }
    return v;
}
__device__ __forceinline__ uint32_t ns_hash(const NsView &v, uint32_t i) { return v.single ? v.tok : v.rec[v.start + 2 * i]; }
__device__ __forceinline__ float ns_val(const NsView &v, uint32_t i) { return (v.single || v.f32) ? 1.0f : __uint_as_float(v.rec[v.start + 2 * i + 1]); }

#define FWGPU_MAX_COMBO_NS 8
__global__ void __launch_bounds__(256) k_translate(const TranslateParams p)
{
    const uint32_t slot_i = blockIdx.x * blockDim.x + threadIdx.x; // position in the output arrays
    if (slot_i >= (p.ex_count ? *p.ex_count : p.n_examples)) return;
    const uint32_t ex = p.ex_list ? p.ex_list[slot_i] : slot_i;     // which record
    const uint32_t *rec = p.records + (p.rec_off ? (size_t)(p.rec_off[ex] - p.off_base) : (size_t)ex * p.fixed_len);
    ExMeta m;
    m.label = (float)rec[1];                 // feature_buffer.rs:190
    m.importance = __uint_as_float(rec[2]);  // :191-192
    m.lr_begin = slot_i * p.lr_stride;
    m.ffm_begin = slot_i * p.ffm_stride;
    m.out_index = ex;
    m.pad1 = 0;
    uint4 *lr = p.lr_ent + (size_t)m.lr_begin;
    uint4 *ffm = p.ffm_ent + (size_t)m.ffm_begin;
    uint32_t nlr = 0, nffm = 0;
    bool overflow = false;

    for (uint32_t c = 0; c < p.n_combos; c++) { // :197-268
        const uint32_t o0 = p.combo_off[c], m_ns = p.combo_off[c + 1] - o0;
        const float w = p.combo_weight[c];
        if (m_ns == 1) {
            NsView v = ns_view(rec, p.combo_ns[o0], p.ns_is_f32);
            for (uint32_t i = 0; i < v.cnt; i++) {
                if (nlr < p.lr_stride) lr[nlr] = make_uint4(ns_hash(v, i) & p.lr_mask, __float_as_uint(ns_val(v, i) * w), c, 0);
                else overflow = true;
                nlr++;
            }
        } else {
            // chained interaction: h = (h * 16777619) ^ h_next, value product; last namespace varies fastest
            NsView vs[FWGPU_MAX_COMBO_NS];
            uint32_t total = 1;
            for (uint32_t j = 0; j < m_ns && j < FWGPU_MAX_COMBO_NS; j++) { vs[j] = ns_view(rec, p.combo_ns[o0 + j], p.ns_is_f32); total *= vs[j].cnt; }
            for (uint32_t t = 0; t < total; t++) {
                uint32_t idxs[FWGPU_MAX_COMBO_NS], r = t;
                for (int j = (int)m_ns - 1; j >= 0; j--) { idxs[j] = r % vs[j].cnt; r /= vs[j].cnt; }
                uint32_t h = ns_hash(vs[0], idxs[0]);
                float val = ns_val(vs[0], idxs[0]);
                for (uint32_t j = 1; j < m_ns; j++) {
                    h = (h * 16777619u) ^ ns_hash(vs[j], idxs[j]);
                    val = val * ns_val(vs[j], idxs[j]);
                }
                if (nlr < p.lr_stride) lr[nlr] = make_uint4(h & p.lr_mask, __float_as_uint(val * w), c, 0);
                else overflow = true;
                nlr++;
            }
        }
    }
    if (p.add_constant) { // :270-276
        if (nlr < p.lr_stride) lr[nlr] = make_uint4(11650396u & p.lr_mask, __float_as_uint(1.0f), p.n_combos, 0);
        else overflow = true;
        nlr++;
    }
    if (p.ffm_k > 0) { // :279-335
        for (uint32_t f = 0; f < p.n_fields; f++) {
            for (uint32_t j = p.field_off[f]; j < p.field_off[f + 1]; j++) {
                NsView v = ns_view(rec, p.field_ns[j], p.ns_is_f32);
                for (uint32_t i = 0; i < v.cnt; i++) {
                    if (nffm < p.ffm_stride) ffm[nffm] = make_uint4(ns_hash(v, i) & p.ffm_mask, __float_as_uint(ns_val(v, i)), f, 0);
                    else overflow = true;
                    nffm++;
                }
            }
        }
    }
    m.lr_cnt = nlr;
    m.ffm_cnt = nffm;
    if (overflow) {
        atomicOr(p.err_flag, 2u);
        if (nlr > p.lr_stride) m.lr_cnt = p.lr_stride;
        // an over-long ffm list keeps its true count so k_learn flags and skips the example
    }
    p.meta[slot_i] = m;
}

// CSR (host layout of fwgpu_batch) -> AoS
struct PackParams {
    uint32_t n_examples;
    const float *labels, *importance;
    const uint32_t *lr_off, *lr_hash, *lr_combo; const float *lr_val;
    const uint32_t *ffm_off, *ffm_hash, *ffm_field; const float *ffm_val;
    uint32_t n_lr, n_ffm;
    ExMeta *meta; uint4 *lr_ent, *ffm_ent;
};
__global__ void __launch_bounds__(256) k_pack(const PackParams p)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.n_examples) {
        ExMeta m;
        m.label = p.labels[i]; m.importance = p.importance[i];
        m.lr_begin = p.lr_off[i]; m.lr_cnt = p.lr_off[i + 1] - p.lr_off[i];
        m.ffm_begin = p.ffm_off ? p.ffm_off[i] : 0; m.ffm_cnt = p.ffm_off ? p.ffm_off[i + 1] - p.ffm_off[i] : 0;
        m.out_index = i;
        m.pad1 = 0;
        p.meta[i] = m;
    }
    if (i < p.n_lr) p.lr_ent[i] = make_uint4(p.lr_hash[i], __float_as_uint(p.lr_val[i]), p.lr_combo[i], 0);
    if (i < p.n_ffm) p.ffm_ent[i] = make_uint4(p.ffm_hash[i], __float_as_uint(p.ffm_val[i]), p.ffm_field[i], 0);
}

// ---------------------------------------------------------------------------------------------
// Table initialisation (block_lr.rs:97-105, block_ffm.rs:784-829)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float merand48(uint64_t seed)
{
    const uint64_t a = 0xeece66d5deece66dULL, c = 2147483647ULL;
    uint64_t s = a * seed + c;
    return __uint_as_float((uint32_t)((s >> 25) & 0x7FFFFFu) | 0x3F800000u) - 1.0f;
}
__global__ void k_init_ffm(float *w, float *acc, uint32_t len, uint32_t first, uint32_t last, float one_over_k_root, float init_acc,
                           float init_width, float zero_band, float center)
{
    // elements [first, last) of the table (a sharded table is initialised range by range, each by its owner); index >= len is padding
    for (size_t i = first + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < last; i += (size_t)gridDim.x * blockDim.x) {
        float wv = 0.0f;
        if (i < len) {
            if (init_width == 0.0f) wv = __fmul_rn(__fsub_rn(1.0f * merand48((uint64_t)len + i), 0.5f), one_over_k_root);
            else {
                float zero_half_band_width = init_width * zero_band * 0.5f;
                float band_width = init_width * (1.0f - zero_band);
                wv = __fsub_rn(__fmul_rn(merand48((uint64_t)i), band_width), band_width * 0.5f);
                if (wv > 0.0f) wv += zero_half_band_width; else wv -= zero_half_band_width;
                wv += center;
            }
        }
        w[i] = wv;
        if (acc) acc[i] = init_acc;
    }
}
__global__ void k_init_lr(float2 *t, size_t len, float init_acc)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) t[i] = make_float2(0.0f, init_acc);
}
__global__ void k_fill(float *p, size_t n, float v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_debug_logistic(const float *in, float *out, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = logistic(in[i]); }
// LR export/import helpers: AoS {w,acc} <-> weights-only
__global__ void k_lr_extract_w(const float2 *t, float *w, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) w[i] = t[i].x; }
__global__ void k_lr_set_w(float2 *t, const float *w, size_t n, float acc) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) t[i] = make_float2(w[i], acc); }

} // namespace fwgpu
