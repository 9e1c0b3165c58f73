// fwgpu.cu -- the C ABI of include/fwgpu.h: context, HBM tables, batch staging, kernel launches.
// Pure CUDA runtime; no torch, no CPU fallback (every entry point needs a live ctx on a GPU).
#include "../../include/fwgpu.h"
#include "fwgpu_kernels.cuh"
#include "fwgpu_head.cuh"
#include "fwgpu_umma.cuh"
#include "fwgpu_shard.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

using namespace fwgpu;

static thread_local std::string g_create_error;

#define CUDA_TRY(ctx, expr)                                                                                   \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                             \
            return FWGPU_ERR_CUDA;                                                                            \
        }                                                                                                     \
    } while (0)

static FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    f.d = d;
    if (d == 0) { f.m = 0; f.s = 0; return f; }
    uint32_t s = 0;
    while ((1ull << s) < d) s++;
    f.s = s;
    f.m = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    return f;
}

// OptimizerAdagradLUT::init (optimizer.rs:121-144); host powf (glibc) exactly as the reference's Rust std does.
static void build_lut(float lr, float power_t, float init_acc, float *lut)
{
    const float minus_power_t = -power_t;
    for (uint32_t x = 0; x < FWGPU_LUT_SIZE; x++) {
        uint32_t b0 = x << 20, b1 = (x + 1) << 20;
        float f0, f1;
        memcpy(&f0, &b0, 4);
        memcpy(&f1, &b1, 4);
        f0 += init_acc;
        f1 += init_acc;
        float val = lr * (powf(f0, minus_power_t) + powf(f1, minus_power_t)) * 0.5f;
        if (std::isnan(val) || std::isinf(val)) val = lr;
        lut[x] = val;
    }
}

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct fwgpu_dataset {
    uint32_t *records = nullptr;
    uint32_t *rec_off = nullptr; // device, [n+1], or null (fixed)
    uint64_t n_words = 0, n_examples = 0;
    uint32_t fixed_len = 0, max_len = 0;
};

struct fwgpu_ctx {
    int device = 0;
    fwgpu_model_desc d{};
    std::vector<uint8_t> ns_is_f32;
    std::vector<uint32_t> combo_off, combo_ns, field_off, field_ns;
    std::vector<float> combo_weight;
    uint32_t n_field_refs = 0;
    uint32_t optimizer = 0; // effective (SGD when immutable)
    uint32_t F = 0, k = 0, Fk = 0, VEC = 1, cpr = 0;
    uint64_t lr_len = 0, ffm_len = 0, ffm_alloc = 0;
    float2 *lr = nullptr;
    float *ffm_w = nullptr, *ffm_acc = nullptr;
    float *lut_dev = nullptr; // 3 * 2048
    float lut_host[3][FWGPU_LUT_SIZE];
    // device copies of the translate spec
    uint8_t *d_ns_is_f32 = nullptr;
    uint32_t *d_combo_off = nullptr, *d_combo_ns = nullptr, *d_field_off = nullptr, *d_field_ns = nullptr;
    float *d_combo_weight = nullptr;
    cudaStream_t stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr, head_side_stream = nullptr;
    cudaEvent_t ev_head_fork = nullptr, ev_head_join = nullptr;
    std::unordered_map<const void *, size_t> smem_optin_set; // kernels whose dynamic shared-memory limit was raised on THIS device (the attribute is per device)
    cudaEvent_t ev_ready[2]{}, ev_free[2]{};
    bool ev_free_recorded[2] = {false, false};
    // predictions leave through their own stream: a chunk's predictions are staged device-to-device on the compute stream
    // (microseconds) and copied to the host while the next chunk is being learned
    cudaEvent_t ev_pred_ready[2]{}, ev_pred_out[2]{};
    bool ev_pred_out_recorded[2] = {false, false};
    DevBuf pred_stage[2];
    // staging
    DevBuf rec[2], rec_off_dev[2], meta, lr_ent, ffm_ent, preds, csr, leftover;
    bool fast_ok = false;     // k_learn_fixed applies to this model (one namespace per field, k % 4 == 0, ...)
    bool fast_enabled = true; // FWGPU_FAST=0 turns the fused kernel off (measurement / debugging)
    bool fast_g16 = true;     // narrow models: two records per warp (FWGPU_G16=0: always one)
    bool fast_cta = false;    // wide model: one block per record (k_learn_fixed_cta) instead of one warp
    bool fast_rows = false;   // ... through k_learn_rows (bulk-copy gather / bulk-reduce scatter) when weights + accumulators of a record fit in shared memory
    int fast_ub = 2;
    uint32_t *err_flag = nullptr;
    uint32_t *err_host = nullptr; // pinned
    int num_sms = 0;
    size_t smem_optin = 0;
    int force_T = 0;
    int minb = 4;
    uint64_t launches = 0;
    uint64_t n_fixed = 0, n_fixed_cta = 0, n_general = 0; // launches per learn-kernel family (fwgpu_debug_path_counts)
    unsigned long long *stat_general_examples = nullptr;  // device: examples the general kernel handled
    uint64_t examples_seen = 0; // examples learned from (update = 1); drives the concurrency ramp
    uint32_t ramp_div = 32;
    uint32_t max_inflight = 0; // 0 = unlimited
    bool ramp_finished = false;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[3];
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[3] = {0, 0, 0};
    uint64_t prof_n[3] = {0, 0, 0};
    // dense head (topology "one", regressor.rs:191-320): hidden layers then the final neuron; all parameters of all
    // layers live in ONE contiguous buffer per kind so the optimizer step is a single launch
    struct HeadLayer { uint32_t n_in, n_out, relu, init; size_t off; };
    std::vector<HeadLayer> head;
    size_t head_params = 0;
    float *head_w = nullptr, *head_acc = nullptr, *head_G1 = nullptr, *head_G2 = nullptr;
    uint32_t x_len = 0, ldx = 0;  // head input: num_combos + F(F+1)/2 (regressor.rs:185-189), padded leading dimension
    uint32_t head_batch = 4096;   // examples per pass around the head's GEMMs = examples in flight
    int head_tile = 0;            // FWGPU_HEAD_TILE: 0 = choose per GEMM, 64 = always 64 x 64 tiles
    uint32_t head_umma_rows = 512; // sub-batches of at least this many rows run the head's GEMMs on the tensor cores (tcgen05, 3xTF32); 0 = never.
                                   // Smaller ones (the sequential / parity mode is 1 row) keep the fp32 FFMA tiles, whose arithmetic is the reference's
    double head_ramp_mul = 2.0;   // sub-batches grow as head_ramp_mul * sqrt(examples_seen) (0 = only the linear ramp)
    uint32_t head_rows_cap = 0;
    DevBuf hX, hdX, hH[FWGPU_MAX_NN_LAYERS], hdZ[FWGPU_MAX_NN_LAYERS], h_label, h_imp, h_outidx, h_dy;
    // hash-range-sharded tables over the GPUs of one box (fwgpu_shard.hpp); null = everything in this GPU's HBM
    ShardGroup *shard = nullptr;
    ShardedArray sh_lr, sh_w, sh_acc;
    // one model over several GPUs, owner-side updates (k_learn_rows<PUSH> + k_apply_inbox): every rank's inbox is one range of
    // sh_inbox; a chunk of at most shard_chunk records per rank is followed by ONE exchange step, the NCCL all-gather of the
    // per-owner push counts (which is also the barrier between "pushed" and "apply")
    bool push_ok = false;
    ShardedArray sh_inbox;
    uint32_t shard_chunk = 8192, inbox_cap = 0, owner_shift = 32;
    uint64_t inbox_rank_bytes = 0, shard_chunks_done = 0;
    uint32_t *push_cnt = nullptr, *counts_all[2] = {nullptr, nullptr};
    bool shard_overlap = false; // apply kernel on a side stream, overlapping the next chunk's push kernel (FWGPU_SHARD_OVERLAP=1)
    cudaStream_t apply_stream = nullptr;
    cudaEvent_t ev_gathered[2]{}, ev_applied[2]{};
    bool ev_applied_rec[2] = {false, false};
    std::string err;
    void set_error(const std::string &s) { err = s; }
};

struct ShardCfg { uint32_t rank, world; const char *rendezvous; uint32_t timeout_ms; };

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function attribute: remember per ctx what was set
template <class K> static cudaError_t ensure_dyn_smem(fwgpu_ctx *c, K kern, size_t smem)
{
    size_t &have = c->smem_optin_set[(const void *)kern];
    if (smem <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) have = smem;
    return e;
}

static fwgpu_status ensure(fwgpu_ctx *c, DevBuf &b, size_t bytes)
{
    if (b.bytes >= bytes) return FWGPU_OK;
    if (b.p) {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->d2h_stream));
        CUDA_TRY(c, cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    CUDA_TRY(c, cudaMalloc(&b.p, want));
    b.bytes = want;
    return FWGPU_OK;
}

template <typename Tp> static fwgpu_status upload_vec(fwgpu_ctx *c, const std::vector<Tp> &v, Tp **out)
{
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(Tp);
    CUDA_TRY(c, cudaMalloc((void **)out, bytes));
    if (!v.empty()) CUDA_TRY(c, cudaMemcpy(*out, v.data(), v.size() * sizeof(Tp), cudaMemcpyHostToDevice));
    return FWGPU_OK;
}

extern "C" const char *fwgpu_version(void) { return "fwgpu 0.1 sm_100a"; }

extern "C" const char *fwgpu_last_error(const fwgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" fwgpu_status fwgpu_host_alloc(void **out, uint64_t bytes)
{
    if (!out) return FWGPU_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return FWGPU_ERR_CUDA; }
    return FWGPU_OK;
}
extern "C" void fwgpu_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" void fwgpu_destroy(fwgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    if (c->head_side_stream) cudaStreamSynchronize(c->head_side_stream);
    if (c->shard) {
        if (c->apply_stream) cudaStreamSynchronize(c->apply_stream);
        c->shard->destroy_array(c->sh_lr); c->shard->destroy_array(c->sh_w); c->shard->destroy_array(c->sh_acc); c->shard->destroy_array(c->sh_inbox);
        cudaFree(c->push_cnt); cudaFree(c->counts_all[0]); cudaFree(c->counts_all[1]);
        for (int i = 0; i < 2; i++) { if (c->ev_gathered[i]) cudaEventDestroy(c->ev_gathered[i]); if (c->ev_applied[i]) cudaEventDestroy(c->ev_applied[i]); }
        if (c->apply_stream) cudaStreamDestroy(c->apply_stream);
        delete c->shard;
        c->lr = nullptr; c->ffm_w = nullptr; c->ffm_acc = nullptr;
    }
    cudaFree(c->lr); cudaFree(c->ffm_w); cudaFree(c->ffm_acc); cudaFree(c->lut_dev);
    cudaFree(c->d_ns_is_f32); cudaFree(c->d_combo_off); cudaFree(c->d_combo_ns); cudaFree(c->d_field_off);
    cudaFree(c->d_field_ns); cudaFree(c->d_combo_weight);
    for (DevBuf *b : {&c->rec[0], &c->rec[1], &c->rec_off_dev[0], &c->rec_off_dev[1], &c->meta, &c->lr_ent, &c->ffm_ent, &c->preds, &c->csr, &c->leftover, &c->pred_stage[0], &c->pred_stage[1]})
        if (b->p) cudaFree(b->p);
    cudaFree(c->head_w); cudaFree(c->head_acc); cudaFree(c->head_G1); cudaFree(c->head_G2);
    for (DevBuf *b : {&c->hX, &c->hdX, &c->h_label, &c->h_imp, &c->h_outidx, &c->h_dy}) if (b->p) cudaFree(b->p);
    for (int i = 0; i < FWGPU_MAX_NN_LAYERS; i++) { if (c->hH[i].p) cudaFree(c->hH[i].p); if (c->hdZ[i].p) cudaFree(c->hdZ[i].p); }
    cudaFree(c->err_flag);
    cudaFree(c->stat_general_examples);
    if (c->err_host) cudaFreeHost(c->err_host);
    for (int i = 0; i < 2; i++) { if (c->ev_ready[i]) cudaEventDestroy(c->ev_ready[i]); if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]); }
    for (int i = 0; i < 2; i++) { if (c->ev_pred_ready[i]) cudaEventDestroy(c->ev_pred_ready[i]); if (c->ev_pred_out[i]) cudaEventDestroy(c->ev_pred_out[i]); }
    for (int kk = 0; kk < 3; kk++) for (auto &pr : c->prof[kk]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->head_side_stream) cudaStreamDestroy(c->head_side_stream);
    if (c->ev_head_fork) cudaEventDestroy(c->ev_head_fork);
    if (c->ev_head_join) cudaEventDestroy(c->ev_head_join);
    delete c;
}

static fwgpu_status create_impl(const fwgpu_model_desc *desc, int device, fwgpu_ctx *c, const ShardCfg *sc)
{
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0) {
        c->set_error(std::string("no CUDA device: ") + (e0 != cudaSuccess ? cudaGetErrorString(e0) : "device count 0") +
                     " (fwgpu has no CPU fallback)");
        return FWGPU_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { c->set_error("device index out of range"); return FWGPU_ERR_INVALID; }
    c->device = device;
    CUDA_TRY(c, cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(c, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { c->set_error("fwgpu kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor)); return FWGPU_ERR_UNSUPPORTED; }
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;

    c->d = *desc;
    const fwgpu_model_desc &d = c->d;
    if (d.bit_precision == 0 || d.bit_precision > 31) { c->set_error("bit_precision must be in 1..31"); return FWGPU_ERR_INVALID; }
    if (d.optimizer > FWGPU_OPT_ADAGRAD_LUT) { c->set_error("unknown optimizer"); return FWGPU_ERR_INVALID; }
    if (d.nn_num_layers > FWGPU_MAX_NN_LAYERS) { c->set_error("too many nn layers"); return FWGPU_ERR_INVALID; }
    c->optimizer = d.immutable ? FWGPU_OPT_SGD : d.optimizer;
    c->F = d.ffm_k > 0 ? d.ffm_num_fields : 0;
    c->k = d.ffm_k;
    c->Fk = c->F * c->k;
    if (d.ffm_k > 0) {
        if (d.ffm_num_fields == 0) { c->set_error("ffm_k > 0 needs ffm_num_fields > 0"); return FWGPU_ERR_INVALID; }
        if (d.ffm_bit_precision == 0 || d.ffm_bit_precision > 31) { c->set_error("ffm_bit_precision must be in 1..31"); return FWGPU_ERR_INVALID; }
        // regressor.rs:23 / block_ffm.rs:97-101: k * F^2 <= FFM_CONTRA_BUF_LEN
        if ((uint64_t)d.ffm_k * d.ffm_num_fields * d.ffm_num_fields > 41472ull) {
            c->set_error("ffm_k * number_of_fields^2 exceeds FFM_CONTRA_BUF_LEN (41472), as in the reference");
            return FWGPU_ERR_UNSUPPORTED;
        }
        uint32_t kp = 1;
        while (kp < d.ffm_k) kp <<= 1;
        uint32_t vec = 4;
        while (vec > 1 && (vec > kp || (c->Fk % vec) != 0)) vec >>= 1;
        c->VEC = vec;
        c->cpr = c->Fk / vec;
    }
    // translate spec copies
    if (d.n_namespaces) {
        if (d.ns_is_f32) c->ns_is_f32.assign(d.ns_is_f32, d.ns_is_f32 + d.n_namespaces);
        else c->ns_is_f32.assign(d.n_namespaces, 0);
    }
    if (d.n_combos) {
        if (!d.combo_off || !d.combo_ns || !d.combo_weight) { c->set_error("combo arrays missing"); return FWGPU_ERR_INVALID; }
        c->combo_off.assign(d.combo_off, d.combo_off + d.n_combos + 1);
        c->combo_ns.assign(d.combo_ns, d.combo_ns + c->combo_off.back());
        c->combo_weight.assign(d.combo_weight, d.combo_weight + d.n_combos);
        for (uint32_t i = 0; i < d.n_combos; i++) {
            uint32_t m = c->combo_off[i + 1] - c->combo_off[i];
            if (m == 0 || m > FWGPU_MAX_COMBO_NS) { c->set_error("a feature combo must have 1..8 namespaces"); return FWGPU_ERR_UNSUPPORTED; }
        }
        for (uint32_t ns : c->combo_ns) if (ns >= d.n_namespaces) { c->set_error("combo namespace index out of range"); return FWGPU_ERR_INVALID; }
    } else {
        c->combo_off.assign(1, 0);
    }
    if (d.ffm_k > 0 && d.field_off && d.field_ns) {
        c->field_off.assign(d.field_off, d.field_off + d.ffm_num_fields + 1);
        c->field_ns.assign(d.field_ns, d.field_ns + c->field_off.back());
        for (uint32_t ns : c->field_ns) if (ns >= d.n_namespaces) { c->set_error("field namespace index out of range"); return FWGPU_ERR_INVALID; }
    } else {
        c->field_off.assign(c->F + 1, 0);
    }
    c->n_field_refs = (uint32_t)c->field_ns.size();
    c->d.ns_is_f32 = nullptr; c->d.combo_off = nullptr; c->d.combo_ns = nullptr; c->d.combo_weight = nullptr;
    c->d.field_off = nullptr; c->d.field_ns = nullptr;

    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    if (!getenv("FWGPU_HEAD_NO_FORK")) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->head_side_stream, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_head_fork, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_head_join, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_ready[i], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_pred_ready[i], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_pred_out[i], cudaEventDisableTiming));
    }
    CUDA_TRY(c, cudaMalloc((void **)&c->err_flag, 4));
    CUDA_TRY(c, cudaMemset(c->err_flag, 0, 4));
    CUDA_TRY(c, cudaMalloc((void **)&c->stat_general_examples, 8));
    CUDA_TRY(c, cudaMemset(c->stat_general_examples, 0, 8));
    CUDA_TRY(c, cudaHostAlloc((void **)&c->err_host, 4, cudaHostAllocDefault));
    *c->err_host = 0;

    // LUTs (built for every optimizer kind; only read under AdagradLUT)
    build_lut(d.learning_rate, d.power_t, d.init_acc_gradient, c->lut_host[0]);
    build_lut(d.ffm_learning_rate, d.ffm_power_t, d.ffm_init_acc_gradient, c->lut_host[1]);
    build_lut(d.nn_learning_rate, d.nn_power_t, d.nn_init_acc_gradient, c->lut_host[2]);
    CUDA_TRY(c, cudaMalloc((void **)&c->lut_dev, sizeof(c->lut_host)));
    CUDA_TRY(c, cudaMemcpy(c->lut_dev, c->lut_host, sizeof(c->lut_host), cudaMemcpyHostToDevice));

    // tables
    c->lr_len = 1ull << d.bit_precision;
    if (d.ffm_k > 0) {
        c->ffm_len = (1ull << d.ffm_bit_precision) + c->Fk; // block_ffm.rs:93-94
        c->ffm_alloc = c->ffm_len + 64;
    }
    // which elements this rank initialises: everything, or (sharded) the hash range it owns
    uint64_t lr_lo = 0, lr_hi = c->lr_len, ffm_lo = 0, ffm_hi = c->ffm_alloc;
    if (sc) {
        if (d.nn_num_layers) { c->set_error("a dense head is replicated per GPU and cannot be combined with a sharded table yet"); return FWGPU_ERR_UNSUPPORTED; }
        c->shard = new ShardGroup();
        ShardGroup &g = *c->shard;
        if (!g.start(sc->rank, sc->world, device, sc->rendezvous, sc->timeout_ms)) { c->set_error("shard group: " + g.error); return FWGPU_ERR_CUDA; }
        std::vector<size_t> sizes;
        auto range = [&](const ShardedArray &a, size_t elem, uint64_t n, uint64_t &lo, uint64_t &hi) {
            lo = std::min<uint64_t>(a.offsets[g.rank] / elem, n);
            hi = a.sizes[g.rank] ? std::min<uint64_t>((a.offsets[g.rank] + a.sizes[g.rank]) / elem, n) : lo;
        };
        shard_plan(c->lr_len * sizeof(float2), 0, g.world, g.granularity, sizes);
        if (!g.create_array(c->sh_lr, sizes)) { c->set_error("sharded LR table: " + g.error); return FWGPU_ERR_CUDA; }
        c->lr = (float2 *)c->sh_lr.va;
        range(c->sh_lr, sizeof(float2), c->lr_len, lr_lo, lr_hi);
        if (d.ffm_k > 0) {
            shard_plan((1ull << d.ffm_bit_precision) * 4, (c->ffm_alloc - (1ull << d.ffm_bit_precision)) * 4, g.world, g.granularity, sizes);
            if (!g.create_array(c->sh_w, sizes)) { c->set_error("sharded FFM weights: " + g.error); return FWGPU_ERR_CUDA; }
            c->ffm_w = (float *)c->sh_w.va;
            if (c->optimizer != FWGPU_OPT_SGD) {
                if (!g.create_array(c->sh_acc, sizes)) { c->set_error("sharded FFM accumulators: " + g.error); return FWGPU_ERR_CUDA; }
                c->ffm_acc = (float *)c->sh_acc.va;
            }
            range(c->sh_w, 4, c->ffm_alloc, ffm_lo, ffm_hi);
        }
    } else {
        CUDA_TRY(c, cudaMalloc((void **)&c->lr, c->lr_len * sizeof(float2)));
        if (d.ffm_k > 0) {
            CUDA_TRY(c, cudaMalloc((void **)&c->ffm_w, c->ffm_alloc * sizeof(float)));
            if (c->optimizer != FWGPU_OPT_SGD) CUDA_TRY(c, cudaMalloc((void **)&c->ffm_acc, c->ffm_alloc * sizeof(float)));
        }
    }
    // initial_data(): Flex starts at init_acc (optimizer.rs:91-93), LUT at 0 (optimizer.rs:158-161)
    const float lr_acc0 = c->optimizer == FWGPU_OPT_ADAGRAD_FLEX ? d.init_acc_gradient : 0.0f;
    if (lr_hi > lr_lo) {
        k_init_lr<<<c->num_sms * 4, 256, 0, c->stream>>>(c->lr + lr_lo, lr_hi - lr_lo, lr_acc0);
        c->launches++;
    }
    if (d.ffm_k > 0 && ffm_hi > ffm_lo) {
        const float ffm_acc0 = c->optimizer == FWGPU_OPT_ADAGRAD_FLEX ? d.ffm_init_acc_gradient : 0.0f;
        const float one_over_k_root = 1.0f / sqrtf((float)d.ffm_k) / 50.0f; // block_ffm.rs:798
        k_init_ffm<<<c->num_sms * 8, 256, 0, c->stream>>>(c->ffm_w, c->ffm_acc, (uint32_t)c->ffm_len, (uint32_t)ffm_lo, (uint32_t)ffm_hi, one_over_k_root,
                                                          ffm_acc0, d.ffm_init_width, d.ffm_init_zero_band, d.ffm_init_center);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    if (d.nn_num_layers > 0) {
        // BlockNeuronLayer::allocate_and_init_weights (block_neural.rs:367-412): weights by init type, biases 0,
        // accumulators at initial_data().  Hu / Xavier: see FWGPU_NN_INIT_* in fwgpu.h (parity unpinned).
        c->x_len = d.num_combos + (c->F ? c->F * (c->F + 1) / 2 : 0);
        c->ldx = (c->x_len + 3) & ~3u;
        uint32_t n_in = c->x_len;
        size_t off = 0;
        for (uint32_t l = 0; l <= d.nn_num_layers; l++) {
            const bool fin = l == d.nn_num_layers;
            fwgpu_ctx::HeadLayer L;
            L.n_in = fin ? n_in + c->x_len : n_in; // join [h, x] (regressor.rs:303-307)
            L.n_out = fin ? 1 : d.nn_width[l];
            L.relu = fin ? 0 : d.nn_relu[l];
            L.init = fin ? FWGPU_NN_INIT_ONE : d.nn_init[l]; // regressor.rs:308-315
            if (L.n_out == 0 || L.n_in >= 16000) { c->set_error("nn layer width must be > 0 and inputs < 16000 (block_neural.rs:27,79)"); return FWGPU_ERR_INVALID; }
            L.off = off;
            off += (size_t)(L.n_in + 1) * L.n_out;
            off = (off + 3) & ~(size_t)3;
            c->head.push_back(L);
            n_in = L.n_out;
        }
        c->head_params = off;
        std::vector<float> w(off, 0.0f);
        uint64_t seed = 12345;
        auto uni = [&]() { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; return (float)((seed >> 40) & 0xFFFFFF) / 16777216.0f; };
        for (auto &L : c->head) {
            const size_t nw = (size_t)L.n_in * L.n_out;
            if (L.init == FWGPU_NN_INIT_ONE) for (size_t i = 0; i < nw; i++) w[L.off + i] = 1.0f;
            else if (L.init == FWGPU_NN_INIT_HU || L.init == FWGPU_NN_INIT_XAVIER) {
                const double sd = L.init == FWGPU_NN_INIT_HU ? sqrt(2.0 / L.n_in) : sqrt(2.0 / (double)nw);
                const double bound = sd * sqrt(3.0);
                for (size_t i = 0; i < nw; i++) w[L.off + i] = (float)((2.0 * uni() - 1.0) * bound);
            }
        }
        CUDA_TRY(c, cudaMalloc((void **)&c->head_w, off * 4));
        CUDA_TRY(c, cudaMemcpy(c->head_w, w.data(), off * 4, cudaMemcpyHostToDevice));
        if (c->optimizer != FWGPU_OPT_SGD) {
            CUDA_TRY(c, cudaMalloc((void **)&c->head_acc, off * 4));
            const float acc0 = c->optimizer == FWGPU_OPT_ADAGRAD_FLEX ? d.nn_init_acc_gradient : 0.0f;
            k_fill<<<c->num_sms, 256, 0, c->stream>>>(c->head_acc, off, acc0);
            c->launches++;
        }
        if (!d.immutable) {
            CUDA_TRY(c, cudaMalloc((void **)&c->head_G1, off * 4));
            CUDA_TRY(c, cudaMalloc((void **)&c->head_G2, off * 4));
            CUDA_TRY(c, cudaMemsetAsync(c->head_G1, 0, off * 4, c->stream));
            CUDA_TRY(c, cudaMemsetAsync(c->head_G2, 0, off * 4, c->stream));
        }
        if (const char *t = getenv("FWGPU_HEAD_BATCH")) c->head_batch = std::max(1, atoi(t));
        if (const char *t = getenv("FWGPU_HEAD_RAMP_MUL")) c->head_ramp_mul = atof(t);
        if (const char *t = getenv("FWGPU_HEAD_TILE")) c->head_tile = atoi(t);
        if (const char *t = getenv("FWGPU_HEAD_UMMA_ROWS")) c->head_umma_rows = (uint32_t)strtoul(t, nullptr, 10);
    }

    fwgpu_status st;
    if ((st = upload_vec(c, c->ns_is_f32, &c->d_ns_is_f32))) return st;
    if ((st = upload_vec(c, c->combo_off, &c->d_combo_off))) return st;
    if ((st = upload_vec(c, c->combo_ns, &c->d_combo_ns))) return st;
    if ((st = upload_vec(c, c->combo_weight, &c->d_combo_weight))) return st;
    if ((st = upload_vec(c, c->field_off, &c->d_field_off))) return st;
    if ((st = upload_vec(c, c->field_ns, &c->d_field_ns))) return st;
    {
        const uint32_t n_lr_max = c->d.n_combos + (c->d.add_constant ? 1u : 0u);
        bool base = (c->k % 4 == 0) && c->n_field_refs == c->F && c->d.n_namespaces > 0;
        for (uint32_t f = 0; base && f < c->F; f++) base = (c->field_off[f + 1] - c->field_off[f]) == 1;
        const uint32_t n_chunks = c->F * (c->Fk / 4);
        bool ok = base && c->F <= 32 && n_lr_max <= 64 && n_chunks <= 128; // warp per record
        if (base && (!ok || !c->head.empty())) {
            // wide model: the block-per-record kernel, if a record's rows fit in shared memory at least twice per SM
            // (models with a dense head always take it: it is the fused kernel that has the two-pass form)
            const size_t smem_need = (size_t)c->F * c->Fk * 4 + (size_t)c->F * 8 + 64;
            c->fast_cta = c->F > 0 && c->F <= 256 && n_lr_max <= 256 && smem_need * 2 <= c->smem_optin;
            ok = c->fast_cta;
        }
        if (c->fast_cta) {
            // k_learn_rows: weight AND accumulator rows of one record in shared memory, pair units held in registers
            const uint32_t k4 = c->k / 4, lpp = (k4 == 1 || k4 == 2 || k4 == 4) ? k4 : 1;
            const size_t n_units = (size_t)c->F * (c->F - 1) / 2 * lpp;
            const size_t smem_rows = (size_t)2 * c->F * c->Fk * 4 + 2048 * 4 + 2 * (size_t)((c->F + 3) & ~3u) * 4 + 64;
            c->fast_rows = c->F >= 2 && n_units <= (size_t)ROWS_MAXU * 256 && smem_rows <= c->smem_optin;
            if (const char *t = getenv("FWGPU_ROWS")) c->fast_rows = c->fast_rows && atoi(t) != 0;
        }
        if (const char *t = getenv("FWGPU_UB")) c->fast_ub = atoi(t);
        c->fast_ok = ok;
        if (const char *t = getenv("FWGPU_G16")) c->fast_g16 = atoi(t) != 0;
        if (const char *t = getenv("FWGPU_FAST")) c->fast_enabled = atoi(t) != 0;
    }
    if (const char *t = getenv("FWGPU_T")) c->force_T = atoi(t);
    if (const char *t = getenv("FWGPU_MINB")) c->minb = atoi(t);
    c->ramp_div = d.hogwild_ramp_div ? d.hogwild_ramp_div : 32;
    if (const char *t = getenv("FWGPU_RAMP_DIV")) c->ramp_div = (uint32_t)strtoul(t, nullptr, 10);
    {
        const bool constant_step = c->optimizer == FWGPU_OPT_SGD || d.power_t == 0.0f || (d.ffm_k > 0 && d.ffm_power_t == 0.0f);
        c->max_inflight = d.hogwild_max_inflight ? d.hogwild_max_inflight : (constant_step ? 16u : 0u);
        if (const char *t = getenv("FWGPU_MAX_INFLIGHT")) c->max_inflight = (uint32_t)strtoul(t, nullptr, 10);
    }
    if (c->shard && c->fast_rows && c->F > 0 && !d.immutable && c->shard->world <= 16 && !getenv("FWGPU_SHARD_DIRECT")) {
        // Owner-side update path.  Collective: every rank takes the same decisions from the same descriptor.
        ShardGroup &g = *c->shard;
        const uint64_t shard_floats = c->sh_w.sizes.size() > 1 && c->sh_w.sizes[1] ? c->sh_w.sizes[0] / 4 : 0;
        bool pow2 = shard_floats && (shard_floats & (shard_floats - 1)) == 0;
        c->owner_shift = 32; // everything on rank 0 (table too small to split)
        if (pow2) { c->owner_shift = 0; while ((1ull << c->owner_shift) < shard_floats) c->owner_shift++; }
        if (shard_floats == 0 || pow2) {
            if (const char *t = getenv("FWGPU_SHARD_CHUNK")) c->shard_chunk = std::max(64, atoi(t));
            if (const char *t = getenv("FWGPU_SHARD_OVERLAP")) c->shard_overlap = atoi(t) != 0;
            if (!g.comm_init()) { c->set_error("shard group (NCCL): " + g.error); return FWGPU_ERR_NCCL; }
            const uint64_t entry_bytes = (uint64_t)(c->Fk + ROWS_HDR) * 4;
            c->inbox_cap = c->shard_chunk * c->F;                               // worst case: every row of a chunk goes to one owner
            c->inbox_rank_bytes = g.round_up(2ull * g.world * c->inbox_cap * entry_bytes); // [half][source][cap]
            std::vector<size_t> sizes(g.world, (size_t)c->inbox_rank_bytes);
            if (!g.create_array(c->sh_inbox, sizes)) { c->set_error("sharded inbox: " + g.error); return FWGPU_ERR_CUDA; }
            CUDA_TRY(c, cudaMalloc((void **)&c->push_cnt, 16 * 4));
            for (int i = 0; i < 2; i++) {
                CUDA_TRY(c, cudaMalloc((void **)&c->counts_all[i], 16 * 16 * 4));
                CUDA_TRY(c, cudaMemset(c->counts_all[i], 0, 16 * 16 * 4));
                CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_gathered[i], cudaEventDisableTiming));
                CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_applied[i], cudaEventDisableTiming));
            }
            CUDA_TRY(c, cudaStreamCreateWithFlags(&c->apply_stream, cudaStreamNonBlocking));
            c->push_ok = true;
        }
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    // nobody trains before every shard is initialised
    if (c->shard && !c->shard->barrier()) { c->set_error("shard group: " + c->shard->error); return FWGPU_ERR_CUDA; }
    return FWGPU_OK;
}

static fwgpu_status create_common(const fwgpu_model_desc *desc, int device, const ShardCfg *sc, fwgpu_ctx **out)
{
    if (!desc || !out) { g_create_error = "null argument"; return FWGPU_ERR_INVALID; }
    *out = nullptr;
    fwgpu_ctx *c = new fwgpu_ctx();
    fwgpu_status st;
    try {
        st = create_impl(desc, device, c, sc);
    } catch (const std::exception &e) {
        c->set_error(e.what());
        st = FWGPU_ERR_INVALID;
    }
    if (st != FWGPU_OK) {
        g_create_error = c->err;
        fwgpu_destroy(c);
        return st;
    }
    *out = c;
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_create(const fwgpu_model_desc *desc, int device, fwgpu_ctx **out) { return create_common(desc, device, nullptr, out); }

extern "C" fwgpu_status fwgpu_create_sharded(const fwgpu_model_desc *desc, int device, uint32_t rank, uint32_t world, const char *rendezvous,
                                             uint32_t timeout_ms, fwgpu_ctx **out)
{
    if (!rendezvous || world == 0 || rank >= world) { g_create_error = "bad shard arguments"; return FWGPU_ERR_INVALID; }
    ShardCfg sc{rank, world, rendezvous, timeout_ms};
    return create_common(desc, device, &sc, out);
}

// Pure layout arithmetic of the sharded tables (no GPU needed; tests/test_dist_cpu.py): how a table of `bytes` (+ `tail_bytes` after
// its last element) splits over `world` ranks at allocation granularity `granularity`, and the shift that maps a float index
// to its owner (32 = everything on rank 0).
extern "C" fwgpu_status fwgpu_debug_shard_plan(uint64_t bytes, uint64_t tail_bytes, uint32_t world, uint64_t granularity, uint64_t *sizes_out, uint32_t *owner_shift_out)
{
    if (!sizes_out || world == 0 || granularity == 0) return FWGPU_ERR_INVALID;
    std::vector<size_t> sizes;
    shard_plan((size_t)bytes, (size_t)tail_bytes, world, (size_t)granularity, sizes);
    for (uint32_t i = 0; i < world; i++) sizes_out[i] = sizes[i];
    if (owner_shift_out) {
        const uint64_t shard_floats = world > 1 && sizes[1] ? sizes[0] / 4 : 0;
        uint32_t sh = 32;
        if (shard_floats && (shard_floats & (shard_floats - 1)) == 0) { sh = 0; while ((1ull << sh) < shard_floats) sh++; }
        *owner_shift_out = sh;
    }
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_shard_barrier(fwgpu_ctx *c)
{
    if (!c) return FWGPU_ERR_INVALID;
    fwgpu_status st = fwgpu_sync(c); // this rank's updates have reached the owners' L2 before anyone moves on
    if (st != FWGPU_OK) return st;
    if (!c->shard) return FWGPU_OK;
    if (!c->shard->barrier()) { c->set_error("shard group: " + c->shard->error); return FWGPU_ERR_CUDA; }
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_shard_info(const fwgpu_ctx *c, uint32_t *rank, uint32_t *world, uint64_t *ffm_first, uint64_t *ffm_count)
{
    if (!c) return FWGPU_ERR_INVALID;
    if (rank) *rank = c->shard ? c->shard->rank : 0;
    if (world) *world = c->shard ? c->shard->world : 1;
    uint64_t lo = 0, n = c->ffm_len;
    if (c->shard && c->ffm_len) {
        const ShardedArray &a = c->sh_w;
        lo = std::min<uint64_t>(a.offsets[c->shard->rank] / 4, c->ffm_len);
        n = std::min<uint64_t>((a.offsets[c->shard->rank] + a.sizes[c->shard->rank]) / 4, c->ffm_len) - lo;
    }
    if (ffm_first) *ffm_first = lo;
    if (ffm_count) *ffm_count = n;
    return FWGPU_OK;
}

extern "C" void *fwgpu_stream(fwgpu_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t fwgpu_launch_count(const fwgpu_ctx *c) { return c ? c->launches : 0; }

static fwgpu_status check_err_flag(fwgpu_ctx *c)
{
    if (*c->err_host) {
        uint32_t f = *c->err_host;
        *c->err_host = 0;
        cudaMemsetAsync(c->err_flag, 0, 4, c->stream);
        cudaStreamSynchronize(c->stream);
        c->set_error(std::string("an example exceeded the staging capacity (") + ((f & 1) ? "ffm features > max_ffm_per_example " : "") +
                     ((f & 2) ? "translate slab overflow " : "") + "); its prediction is NaN and it was not learned. Raise max_ffm_per_example / max_lr_per_example.");
        return FWGPU_ERR_TOO_LARGE;
    }
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_sync(fwgpu_ctx *c)
{
    if (!c) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaMemcpyAsync(c->err_host, c->err_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->d2h_stream));
    if (c->apply_stream) CUDA_TRY(c, cudaStreamSynchronize(c->apply_stream));
    return check_err_flag(c);
}

// ---- profiling -------------------------------------------------------------------------------
static cudaEvent_t get_event(fwgpu_ctx *c)
{
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    fwgpu_ctx *c; int kind; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(fwgpu_ctx *c_, int kind_) : c(c_), kind(kind_)
    {
        if (c->profiling) { a = get_event(c); b = get_event(c); cudaEventRecord(a, c->stream); }
    }
    ~ProfScope()
    {
        if (c->profiling) { cudaEventRecord(b, c->stream); c->prof[kind].push_back({a, b}); }
    }
};
extern "C" fwgpu_status fwgpu_set_profiling(fwgpu_ctx *c, int enabled)
{
    if (!c) return FWGPU_ERR_INVALID;
    c->profiling = enabled != 0;
    return FWGPU_OK;
}
extern "C" fwgpu_status fwgpu_kernel_time(fwgpu_ctx *c, int kind, double *total_ms, uint64_t *launches)
{
    if (!c || kind < 0 || kind > 2) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (auto &pr : c->prof[kind]) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pr.first, pr.second);
        c->prof_ms[kind] += ms;
        c->prof_n[kind]++;
        c->ev_pool.push_back(pr.first);
        c->ev_pool.push_back(pr.second);
    }
    c->prof[kind].clear();
    if (total_ms) *total_ms = c->prof_ms[kind];
    if (launches) *launches = c->prof_n[kind];
    c->prof_ms[kind] = 0;
    c->prof_n[kind] = 0;
    return FWGPU_OK;
}

// ---- learn kernel launch ----------------------------------------------------------------------
template <int T, int VEC, int MINB> static cudaError_t launch_learn_tvm(fwgpu_ctx *c, const LearnParams &p, size_t smem, uint32_t *full_groups)
{
    auto kern = k_learn<T, VEC, MINB>;
    cudaError_t e = ensure_dyn_smem(c, kern, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    constexpr int GROUPS = 256 / T;
    uint32_t need = (p.n_examples + GROUPS - 1) / GROUPS;
    uint32_t grid = std::min<uint32_t>(need, (uint32_t)(c->num_sms * per_sm));
    if (p.max_groups) grid = std::min<uint32_t>(grid, (p.max_groups + GROUPS - 1) / GROUPS);
    if (full_groups) *full_groups = (uint32_t)(c->num_sms * per_sm) * GROUPS;
    if (grid == 0) return cudaSuccess;
    kern<<<grid, 256, smem, c->stream>>>(p);
    c->launches++; c->n_general++;
    return cudaGetLastError();
}

// registers per thread decide how many one-example blocks (T = 256) an SM holds: MINB trades ILP for occupancy
template <int T, int VEC> static cudaError_t launch_learn_tv(fwgpu_ctx *c, const LearnParams &p, size_t smem, uint32_t *full_groups)
{
    if (T == 256) {
        switch (c->minb) {
        case 2: return launch_learn_tvm<T, VEC, 2>(c, p, smem, full_groups);
        case 4: return launch_learn_tvm<T, VEC, 4>(c, p, smem, full_groups);
        default: return launch_learn_tvm<T, VEC, 3>(c, p, smem, full_groups);
        }
    }
    return launch_learn_tvm<T, VEC, 1>(c, p, smem, full_groups);
}

template <int T> static cudaError_t launch_learn_t(fwgpu_ctx *c, const LearnParams &p, size_t smem, uint32_t *full_groups)
{
    switch (c->VEC) {
    case 4: return launch_learn_tv<T, 4>(c, p, smem, full_groups);
    case 2: return launch_learn_tv<T, 2>(c, p, smem, full_groups);
    default: return launch_learn_tv<T, 1>(c, p, smem, full_groups);
    }
}

// shape-dependent part of LearnParams, thread-group width T and dynamic shared memory of k_learn
static fwgpu_status make_learn_params(fwgpu_ctx *c, uint32_t n_cap, int update, uint32_t lr_cap, LearnParams &p, int &T, size_t &smem)
{
    p = LearnParams{};
    p.lr = c->lr; p.ffm_w = c->ffm_w; p.ffm_acc = c->ffm_acc;
    p.lut_lr = c->lut_dev; p.lut_ffm = c->lut_dev + FWGPU_LUT_SIZE;
    p.meta = (const ExMeta *)c->meta.p; p.lr_ent = (const uint4 *)c->lr_ent.p; p.ffm_ent = (const uint4 *)c->ffm_ent.p;
    p.preds = (float *)c->preds.p;
    p.F = c->F; p.k = c->k; p.Fk = c->Fk; p.cpr = c->cpr; p.n_cap = std::max<uint32_t>(n_cap, 1);
    p.div_cpr = make_fastdiv(std::max<uint32_t>(c->cpr, 1)); p.div_k = make_fastdiv(std::max<uint32_t>(c->k, 1)); p.div_F = make_fastdiv(std::max<uint32_t>(c->F, 1));
    p.optimizer = c->optimizer;
    p.lr_lr = c->d.learning_rate; p.lr_mpt = -c->d.power_t; p.ffm_lr = c->d.ffm_learning_rate; p.ffm_mpt = -c->d.ffm_power_t;
    p.update = update; p.err_flag = c->err_flag; p.stat_examples = c->stat_general_examples; p.sys_scope = c->shard ? 1 : 0;
    p.simple_update = 1; // measured on B200 (c3): one chunk at a time with 4 blocks/SM beats rounds of four with 3
    if (const char *t = getenv("FWGPU_SIMPLE_UPDATE")) p.simple_update = atoi(t);
    p.kv = (c->k == 0 || c->k % std::max<uint32_t>(c->VEC, 1) == 0) ? 1 : 0;
    p.lr_cap = std::min<uint32_t>(lr_cap, 1024);
    size_t words = (size_t)c->F * c->Fk + (size_t)p.n_cap * c->k + 3 * (size_t)p.n_cap + (c->F + 1) + 16 + p.lr_cap;
    size_t group_bytes = ((words * 4 + 15) / 16) * 16;
    p.group_smem_bytes = (uint32_t)group_bytes;
    T = 32;
    const uint32_t work = c->F * c->cpr;
    if (work > 1536) T = 256; else if (work > 512) T = 128; else if (work > 128) T = 64;
    if (c->force_T == 32 || c->force_T == 64 || c->force_T == 128 || c->force_T == 256) T = c->force_T;
    smem = group_bytes * (256 / T);
    while (smem > c->smem_optin && T < 256) { T *= 2; smem = group_bytes * (256 / T); }
    if (smem > c->smem_optin) {
        c->set_error("example staging needs " + std::to_string(smem) + " B of shared memory (> " + std::to_string(c->smem_optin) + "): F*F*k or features per example too large");
        return FWGPU_ERR_TOO_LARGE;
    }
    return FWGPU_OK;
}

static fwgpu_status dispatch_learn(fwgpu_ctx *c, const LearnParams &q, int T, size_t smem, uint32_t *full_groups)
{
    cudaError_t e;
    {
        ProfScope ps(c, 0);
        switch (T) {
        case 32: e = launch_learn_t<32>(c, q, smem, full_groups); break;
        case 64: e = launch_learn_t<64>(c, q, smem, full_groups); break;
        case 128: e = launch_learn_t<128>(c, q, smem, full_groups); break;
        default: e = launch_learn_t<256>(c, q, smem, full_groups); break;
        }
    }
    if (e != cudaSuccess) { c->set_error(std::string("k_learn launch: ") + cudaGetErrorString(e)); return FWGPU_ERR_CUDA; }
    return FWGPU_OK;
}

static fwgpu_status head_learn(fwgpu_ctx *c, uint32_t count, int update, uint32_t n_cap, uint32_t lr_cap, const FixedCtaParams *cta,
                               const TranslateParams *tp_left);

static fwgpu_status launch_learn(fwgpu_ctx *c, uint32_t n_examples, uint32_t n_cap, int update, const uint32_t *n_examples_dev = nullptr,
                                 bool count_seen = true, uint32_t lr_cap = 0)
{
    if (n_examples == 0) return FWGPU_OK;
    if (!c->head.empty()) return head_learn(c, n_examples, update, n_cap, lr_cap, nullptr, nullptr);
    LearnParams p;
    int T; size_t smem;
    fwgpu_status st;
    if ((st = make_learn_params(c, n_cap, update, lr_cap, p, T, smem))) return st;
    p.n_examples = n_examples;
    // Concurrency ramp (DESIGN.md "semantics"): a cold model is trained with examples_seen / ramp_div examples
    // in flight; segments double until the whole machine is in use.  Predict-only launches are never limited.
    uint32_t done = 0;
    while (done < n_examples) {
        uint32_t cnt = n_examples - done;
        uint32_t cap = 0;
        if (update && c->ramp_div != 0xffffffffu && !c->ramp_finished) {
            const uint64_t seen = c->examples_seen;
            cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(seen / c->ramp_div, 1), 1u << 30);
            const uint64_t seg_end = std::max<uint64_t>(2 * seen, c->ramp_div);
            if (!n_examples_dev) cnt = (uint32_t)std::min<uint64_t>(cnt, seg_end - seen);
        }
        if (update && c->max_inflight && (cap == 0 || cap > c->max_inflight)) cap = c->max_inflight;
        LearnParams q = p;
        q.meta = p.meta + done;   // ExMeta.out_index is absolute within the chunk, so preds is not offset
        q.n_examples = cnt;
        q.n_examples_dev = n_examples_dev;
        q.max_groups = cap;
        // parity mode: a strictly sequential run (ramp_div >= 2^31-1) or a single-example call sums in the reference's order
        q.exact_order = ((cap == 1 && c->ramp_div >= 0x7fffffffu) || n_examples == 1) ? 1 : 0;
        uint32_t full_groups = 0;
        if ((st = dispatch_learn(c, q, T, smem, &full_groups))) return st;
        if (update) {
            if (count_seen) c->examples_seen += cnt;
            if (cap && full_groups && cap >= full_groups) c->ramp_finished = true;
        }
        done += cnt;
    }
    return FWGPU_OK;
}

// ---- fused fast path (k_learn_fixed) -----------------------------------------------------------
template <int G, int NCH, int NLR, int OPTK, bool PARITY> static cudaError_t launch_fixed_kp(fwgpu_ctx *c, const FixedParams &p, uint32_t *full_groups);
template <int G, int NCH, int NLR, int OPTK> static cudaError_t launch_fixed_k(fwgpu_ctx *c, const FixedParams &p, uint32_t *full_groups)
{
    // one record in flight (update mode): the parity instantiation (tape-order sum, fences, ordered LR duplicates)
    if (p.max_groups == 1 && p.update) return launch_fixed_kp<G, NCH, NLR, OPTK, true>(c, p, full_groups);
    return launch_fixed_kp<G, NCH, NLR, OPTK, false>(c, p, full_groups);
}
template <int G, int NCH, int NLR, int OPTK, bool PARITY> static cudaError_t launch_fixed_kp(fwgpu_ctx *c, const FixedParams &p, uint32_t *full_groups)
{
    auto kern = k_learn_fixed<G, NCH, NLR, OPTK, PARITY>;
    constexpr int NW = FIXED_WARPS, RPW = 32 / G;
    const size_t smem = (size_t)p.rec_smem_floats * 4 * NW * RPW; // the records' row transposes
    cudaError_t e0 = ensure_dyn_smem(c, kern, smem);
    if (e0 != cudaSuccess) return e0;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint32_t per_block = NW * RPW; // records in flight per block
    uint32_t grid = std::min<uint32_t>((p.n_examples + per_block - 1) / per_block, (uint32_t)(c->num_sms * per_sm));
    if (p.max_groups) grid = std::min<uint32_t>(grid, (p.max_groups + per_block - 1) / per_block);
    *full_groups = (uint32_t)(c->num_sms * per_sm) * per_block;
    if (grid == 0) return cudaSuccess;
    kern<<<grid, NW * 32, smem, c->stream>>>(p);
    c->launches++; c->n_fixed++;
    return cudaGetLastError();
}

// AdagradLUT (the reference's default under --adaptive) gets its own instantiation; narrow models (F <= 16, at most 64
// chunks and 16 LR entries) run two records per warp
template <int G, int NCH, int NLR> static cudaError_t launch_fixed_o(fwgpu_ctx *c, const FixedParams &p, uint32_t *full_groups)
{
    if (p.optimizer == OPT_LUT) return launch_fixed_k<G, NCH, NLR, (int)OPT_LUT>(c, p, full_groups);
    return launch_fixed_k<G, NCH, NLR, -1>(c, p, full_groups);
}
template <int G, int NLR> static cudaError_t launch_fixed_g(fwgpu_ctx *c, const FixedParams &p, uint32_t *full_groups)
{
    const uint32_t nch = (p.F * p.cpr + G - 1) / G;
    switch (nch) {
    case 1: return launch_fixed_o<G, 1, NLR>(c, p, full_groups);
    case 2: return launch_fixed_o<G, 2, NLR>(c, p, full_groups);
    case 3: return launch_fixed_o<G, 3, NLR>(c, p, full_groups);
    default: return launch_fixed_o<G, 4, NLR>(c, p, full_groups);
    }
}
static cudaError_t launch_fixed(fwgpu_ctx *c, const FixedParams &p, uint32_t *full_groups)
{
    const uint32_t n_lr = p.n_combos + (p.add_constant ? 1u : 0u), n_chunks = p.F * p.cpr;
    if (c->fast_g16 && p.F <= 16 && n_chunks <= 64 && n_lr <= 16) return launch_fixed_g<16, 1>(c, p, full_groups);
    return n_lr <= 32 ? launch_fixed_g<32, 1>(c, p, full_groups) : launch_fixed_g<32, 2>(c, p, full_groups);
}

template <int UB, int PHASE = 0> static cudaError_t launch_fixed_cta(fwgpu_ctx *c, const FixedCtaParams &p, size_t smem, uint32_t *full_groups)
{
    auto kern = k_learn_fixed_cta<UB, PHASE>;
    cudaError_t e0 = ensure_dyn_smem(c, kern, smem);
    if (e0 != cudaSuccess) return e0;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = std::min<uint32_t>(p.n_examples, (uint32_t)(c->num_sms * per_sm));
    if (p.max_groups) grid = std::min<uint32_t>(grid, p.max_groups);
    *full_groups = (uint32_t)(c->num_sms * per_sm);
    if (grid == 0) return cudaSuccess;
    kern<<<grid, 256, smem, c->stream>>>(p);
    c->launches++; c->n_fixed_cta++;
    return cudaGetLastError();
}

// k_learn_rows: same records, same parameters; shared memory holds the weight rows, and for an updating launch the
// accumulator rows and the FFM look-up table as well
static RowsParams rows_params(const fwgpu_ctx *c, const FixedCtaParams &q)
{
    RowsParams r{};
    r.lr = q.lr; r.ffm_w = q.ffm_w; r.ffm_acc = q.ffm_acc; r.lut_lr = q.lut_lr; r.lut_ffm = q.lut_ffm;
    r.records = q.records; r.rec_off = q.rec_off; r.off_base = q.off_base; r.fixed_len = q.fixed_len;
    r.ex_begin = q.ex_begin; r.n_examples = q.n_examples;
    r.F = q.F; r.k = q.k; r.Fk = q.Fk; r.k4 = q.k / 4;
    r.lpp = (r.k4 == 1 || r.k4 == 2 || r.k4 == 4) ? r.k4 : 1;
    r.n_units = q.F * (q.F - 1) / 2 * r.lpp;
    r.field_ns = q.field_ns; r.n_combos = q.n_combos; r.combo_off = q.combo_off; r.combo_ns = q.combo_ns; r.combo_weight = q.combo_weight;
    r.add_constant = q.add_constant; r.lr_mask = q.lr_mask; r.ffm_mask = q.ffm_mask;
    r.optimizer = q.optimizer; r.lr_lr = q.lr_lr; r.lr_mpt = q.lr_mpt; r.ffm_lr = q.ffm_lr; r.ffm_mpt = q.ffm_mpt;
    r.update = q.update; r.preds = q.preds; r.leftover_idx = q.leftover_idx; r.leftover_cnt = q.leftover_cnt; r.max_groups = q.max_groups;
    r.io = q.io;
    r.sys_scope = c->shard ? 1 : 0;
    return r;
}
template <int PHASE, int OPTK, bool PUSH> static cudaError_t launch_rows_k(fwgpu_ctx *c, const RowsParams &p, uint32_t *full_groups);
template <int PHASE> static cudaError_t launch_rows(fwgpu_ctx *c, const RowsParams &p, uint32_t *full_groups)
{
    if (p.optimizer == OPT_LUT) return launch_rows_k<PHASE, (int)OPT_LUT, false>(c, p, full_groups);
    return launch_rows_k<PHASE, -1, false>(c, p, full_groups);
}
template <int PHASE, int OPTK, bool PUSH> static cudaError_t launch_rows_k(fwgpu_ctx *c, const RowsParams &p, uint32_t *full_groups)
{
    auto kern = k_learn_rows<PHASE, OPTK, PUSH>;
    const bool writes = PHASE != 1 && p.update != 0;
    const size_t rows = (size_t)p.F * (p.Fk + (PUSH ? ROWS_HDR : 0)) * 4 + (PUSH ? 16 : 0);
    const size_t smem = rows + ((!PUSH && writes && p.optimizer != OPT_SGD) ? rows : 0) + ((!PUSH && writes && p.optimizer == OPT_LUT) ? 2048 * 4 : 0) +
                        2 * (size_t)((p.F + 3) & ~3u) * 4 + 8 * 4 + 16 +
                        (p.max_groups == 1 ? (size_t)(p.n_combos + 1 + p.F * (p.F + 1) / 2) * 4 : 0); // parity mode: the tape
    cudaError_t e0 = ensure_dyn_smem(c, kern, smem);
    if (e0 != cudaSuccess) return e0;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (PUSH && c->shard_overlap && per_sm > 2) per_sm = 2; // half of every SM is left to the owner-side apply kernel of the previous chunk
    uint32_t grid = std::min<uint32_t>(p.n_examples, (uint32_t)(c->num_sms * per_sm));
    if (p.max_groups) grid = std::min<uint32_t>(grid, p.max_groups);
    *full_groups = (uint32_t)(c->num_sms * per_sm);
    if (grid == 0) return cudaSuccess;
    kern<<<grid, 256, smem, c->stream>>>(p);
    c->launches++; c->n_fixed_cta++;
    return cudaGetLastError();
}

// One chunk of the sharded owner-side update path: push kernel, exchange step, apply kernel.
//   main stream : [wait: apply of the chunk that last used this inbox half]  zero counters -> k_learn_rows<PUSH> -> [wait: apply of
//                 the previous chunk] -> NCCL all-gather of the per-owner counts (= barrier: every rank has pushed)
//   apply stream: wait all-gather -> k_apply_inbox (overlaps the next chunk's push kernel)
// Inbox halves alternate, so a rank may push chunk i+1 while owners still apply chunk i; chunk i+2 reuses half i only after
// all-gather i+1, which every rank enqueues after its apply of chunk i.
static fwgpu_status shard_push_chunk(fwgpu_ctx *c, RowsParams rp, uint32_t *full_groups)
{
    ShardGroup &g = *c->shard;
    const int half = (int)(c->shard_chunks_done & 1);
    const uint64_t entry_bytes = (uint64_t)(c->Fk + ROWS_HDR) * 4;
    rp.inbox = (unsigned char *)c->sh_inbox.va;
    rp.inbox_rank_stride = c->inbox_rank_bytes;
    rp.inbox_src_off = ((uint64_t)half * g.world + g.rank) * c->inbox_cap * entry_bytes;
    rp.push_cnt = c->push_cnt; rp.owner_shift = c->owner_shift; rp.world = g.world;
    CUDA_TRY(c, cudaMemsetAsync(c->push_cnt, 0, 16 * 4, c->stream));
    cudaError_t e = rp.optimizer == OPT_LUT ? launch_rows_k<0, (int)OPT_LUT, true>(c, rp, full_groups) : launch_rows_k<0, -1, true>(c, rp, full_groups);
    if (e != cudaSuccess) { c->set_error(std::string("k_learn_rows<PUSH> launch: ") + cudaGetErrorString(e)); return FWGPU_ERR_CUDA; }
    // the previous chunk's apply (other half) is ordered before this all-gather: see the reuse argument above
    if (c->ev_applied_rec[half ^ 1]) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_applied[half ^ 1], 0));
    if (!g.all_gather_u32(c->push_cnt, c->counts_all[half], g.world, c->stream)) { c->set_error("shard group (NCCL): " + g.error); return FWGPU_ERR_NCCL; }
    c->launches++;
    cudaStream_t as = c->shard_overlap ? c->apply_stream : c->stream;
    if (c->shard_overlap) {
        CUDA_TRY(c, cudaEventRecord(c->ev_gathered[half], c->stream));
        CUDA_TRY(c, cudaStreamWaitEvent(as, c->ev_gathered[half], 0));
    }
    ApplyParams ap{};
    ap.inbox_half = (const unsigned char *)c->sh_inbox.va + (uint64_t)g.rank * c->inbox_rank_bytes + (uint64_t)half * g.world * c->inbox_cap * entry_bytes;
    ap.counts_all = c->counts_all[half]; ap.world = g.world; ap.rank = g.rank; ap.cap = c->inbox_cap; ap.entry_bytes = (uint32_t)entry_bytes;
    ap.F = c->F; ap.k = c->k; ap.Fk = c->Fk; ap.ffm_w = c->ffm_w; ap.ffm_acc = c->ffm_acc; ap.lut_ffm = c->lut_dev + FWGPU_LUT_SIZE;
    ap.optimizer = c->optimizer; ap.ffm_lr = c->d.ffm_learning_rate; ap.ffm_mpt = -c->d.ffm_power_t;
    static const int apply_blocks_env = getenv("FWGPU_SHARD_APPLY_BLOCKS") ? atoi(getenv("FWGPU_SHARD_APPLY_BLOCKS")) : 0;
    if (!getenv("FWGPU_SHARD_NO_APPLY")) { // (diagnostic: time the push side alone)
        int apply_per_sm = 0;
        CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&apply_per_sm, k_apply_inbox, 256, 0));
        apply_per_sm = std::max(1, apply_blocks_env > 0 ? apply_blocks_env : (c->shard_overlap ? std::min(apply_per_sm, 2) : apply_per_sm)); // one wave
        k_apply_inbox<<<c->num_sms * apply_per_sm, 256, 0, as>>>(ap);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    if (c->shard_overlap) { CUDA_TRY(c, cudaEventRecord(c->ev_applied[half], as)); c->ev_applied_rec[half] = true; }
    c->shard_chunks_done++;
    return FWGPU_OK;
}

// ---- dense head (fwgpu_head.cuh) ----------------------------------------------------------------
template <bool A_T, bool B_T, int EPI, int BM, int BN> static void launch_head_gemm_tile(fwgpu_ctx *c, HeadGemmParams &p, cudaStream_t stream)
{
    uint32_t splits = 1;
    if (EPI == HEAD_EPI_SUMS) {
        // the reduction runs over the sub-batch: split it so that the grid fills the machine about twice
        const uint32_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
        splits = std::max<uint32_t>(1, std::min<uint32_t>((2 * (uint32_t)c->num_sms + tiles - 1) / tiles, (p.K + 63) / 64));
        p.k_split = (((p.K + splits - 1) / splits) + 15) / 16 * 16;
        splits = (p.K + p.k_split - 1) / p.k_split;
    }
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
    k_head_gemm<A_T, B_T, EPI, BM, BN><<<grid, 256, 0, stream>>>(p);
    c->launches++;
}

// tile choice: 128 x 128 (8 x 8 outputs per thread, FFMA-bound) when that still gives every SM a block, else 64 x 64;
// the gradient-sum GEMM keeps two accumulators per output, so its largest tile is 128 x 64
// the same GEMM on the tensor cores (fwgpu_umma.cuh): one 128 x 128 tile per block, the update GEMM split over K so that
// the grid covers the machine
template <bool A_T, bool B_T, int EPI> static void launch_head_umma(fwgpu_ctx *c, HeadGemmParams &p, cudaStream_t stream)
{
    auto kern = k_umma_gemm<A_T, B_T, EPI>;
    if (ensure_dyn_smem(c, kern, UMMA_SMEM_BYTES) != cudaSuccess) { c->set_error("k_umma_gemm: cannot raise the dynamic shared-memory limit"); return; }
    const uint32_t n_cols = p.N + ((EPI == HEAD_EPI_SUMS && p.G1_bias) ? 1u : 0u); // the bias sums ride along as one more column
    const uint32_t tiles = ((p.M + UMMA_BM - 1) / UMMA_BM) * ((n_cols + UMMA_BN - 1) / UMMA_BN);
    uint32_t splits = 1;
    if (EPI == HEAD_EPI_SUMS) {
        splits = std::max<uint32_t>(1, std::min<uint32_t>(((uint32_t)c->num_sms + tiles - 1) / tiles, (p.K + 127) / 128));
        p.k_split = (((p.K + splits - 1) / splits) + 15) / 16 * 16;
        splits = (p.K + p.k_split - 1) / p.k_split;
    }
    dim3 grid((n_cols + UMMA_BN - 1) / UMMA_BN, (p.M + UMMA_BM - 1) / UMMA_BM, splits);
    kern<<<grid, UMMA_THREADS, UMMA_SMEM_BYTES, stream>>>(p);
    c->launches++;
}

template <bool A_T, bool B_T, int EPI> static void launch_head_gemm(fwgpu_ctx *c, HeadGemmParams &p, cudaStream_t stream = nullptr)
{
    if (!stream) stream = c->stream;
    if (c->head_umma_rows && (EPI == HEAD_EPI_SUMS ? p.K : p.M) >= c->head_umma_rows) { launch_head_umma<A_T, B_T, EPI>(c, p, stream); return; }
    const uint64_t big_tiles = (uint64_t)((p.M + 127) / 128) * ((p.N + 127) / 128);
    const bool small = c->head_tile == 64 || (c->head_tile == 0 && (p.M <= 64 || p.N <= 64 || (EPI != HEAD_EPI_SUMS && big_tiles < (uint64_t)c->num_sms * 3 / 4)));
    if (small) launch_head_gemm_tile<A_T, B_T, EPI, 64, 64>(c, p, stream);
    else if constexpr (EPI == HEAD_EPI_SUMS) launch_head_gemm_tile<A_T, B_T, EPI, 128, 64>(c, p, stream);
    else launch_head_gemm_tile<A_T, B_T, EPI, 128, 128>(c, p, stream);
}

// forward (+ backward and optimizer step when update) of the head over `rows` examples whose inputs sit in hX
static fwgpu_status head_pass(fwgpu_ctx *c, uint32_t rows, int update)
{
    ProfScope ps(c, 2);
    const size_t nl = c->head.size(); // hidden layers + final neuron
    float *X = (float *)c->hX.p, *dX = (float *)c->hdX.p, *dy = (float *)c->h_dy.p;
    const float *in = X;
    uint32_t ld_in = c->ldx;
    for (size_t l = 0; l + 1 < nl; l++) { // BlockNeuronLayer forward + BlockRELU (block_neural.rs:196-222, block_relu.rs:79-99)
        const auto &L = c->head[l];
        HeadGemmParams g{};
        g.A = in; g.lda = ld_in; g.B = c->head_w + L.off; g.ldb = L.n_in; g.C = (float *)c->hH[l].p; g.ldc = L.n_out;
        g.M = rows; g.N = L.n_out; g.K = L.n_in; g.bias = c->head_w + L.off + (size_t)L.n_in * L.n_out; g.relu = (int)L.relu;
        launch_head_gemm<false, false, HEAD_EPI_BIAS_ACT>(c, g);
        in = g.C; ld_in = L.n_out;
    }
    const auto &Lf = c->head[nl - 1];
    const auto &Ll = c->head[nl - 2]; // last hidden layer
    {
        HeadFinalParams f{};
        f.H = (const float *)c->hH[nl - 2].p; f.ldh = Ll.n_out; f.n_h = Ll.n_out; f.X = X; f.ldx = c->ldx; f.n_x = c->x_len;
        f.w = c->head_w + Lf.off; f.label = (const float *)c->h_label.p; f.importance = (const float *)c->h_imp.p;
        f.out_index = (const uint32_t *)c->h_outidx.p; f.preds = (float *)c->preds.p; f.dy = dy;
        f.dZ = (float *)c->hdZ[nl - 2].p; f.ldz = Ll.n_out; f.n_rows = rows; f.update = update; f.h_relu = (int)Ll.relu;
        const uint32_t blocks = std::min<uint32_t>((rows + 7) / 8, (uint32_t)c->num_sms * 8);
        k_head_final<<<blocks, 256, 0, c->stream>>>(f);
        c->launches++;
    }
    if (!update) { CUDA_TRY(c, cudaGetLastError()); return FWGPU_OK; }
    // The gradient-sum kernels (this one and the update GEMM of every layer: they read dZ_l / dy and the layer's input) are
    // independent of the GEMMs that carry the error to the layer below (dZ_l and W_l), and each fills less than half of the
    // machine at these sizes: they go to a side stream, forked once their inputs exist and joined before the optimizer step.
    const bool fork = c->head_side_stream != nullptr && c->head_umma_rows && rows >= c->head_umma_rows;
    cudaStream_t sums_stream = fork ? c->head_side_stream : c->stream;
    if (fork) { // dy and the last layer's dZ are complete in main-stream order here
        CUDA_TRY(c, cudaEventRecord(c->ev_head_fork, c->stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->head_side_stream, c->ev_head_fork, 0));
    }
    // final neuron: gradient sums over its inputs [h, x] and its bias (block_neural.rs:266-305 with one neuron)
    {
        HeadFinalSumsParams f{};
        f.H = (const float *)c->hH[nl - 2].p; f.ldh = Ll.n_out; f.n_h = Ll.n_out; f.X = X; f.ldx = c->ldx; f.n_x = c->x_len;
        f.dy = dy; f.n_rows = rows; f.G1 = c->head_G1 + Lf.off; f.G2 = c->head_G2 + Lf.off;
        const uint32_t col_blocks = (Lf.n_in + 1 + 255) / 256;
        const uint32_t row_blocks = std::max<uint32_t>(1, std::min<uint32_t>((rows + 127) / 128, (2 * (uint32_t)c->num_sms + col_blocks - 1) / col_blocks));
        f.rows_per_block = (rows + row_blocks - 1) / row_blocks;
        k_head_final_sums<<<dim3(col_blocks, (rows + f.rows_per_block - 1) / f.rows_per_block), 256, 0, sums_stream>>>(f);
        c->launches++;
    }
    for (size_t li = nl - 1; li-- > 0;) { // hidden layers, last to first
        const auto &L = c->head[li];
        const float *dZ = (const float *)c->hdZ[li].p;
        const float *W = c->head_w + L.off;
        if (fork && li + 2 < nl) { // dZ_l (written by the previous iteration's error GEMM) is complete in main-stream order here
            CUDA_TRY(c, cudaEventRecord(c->ev_head_fork, c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->head_side_stream, c->ev_head_fork, 0));
        }
        // errors for the layer below from the PRE-update weights (block_neural.rs:283-284); the step is applied at the end
        HeadGemmParams g{};
        g.A = dZ; g.lda = L.n_out; g.B = W; g.ldb = L.n_in; g.M = rows; g.N = L.n_in; g.K = L.n_out;
        if (li > 0) {
            g.C = (float *)c->hdZ[li - 1].p; g.ldc = L.n_in; g.mask_src = (const float *)c->hH[li - 1].p; g.ld_mask = L.n_in; g.mask_on = (int)c->head[li - 1].relu;
            launch_head_gemm<false, true, HEAD_EPI_MASK>(c, g);
        } else {
            // BlockCopy backward: d_x = d(path through the layers) + d(direct path into the final neuron) (block_misc.rs:452-473)
            g.C = dX; g.ldc = c->ldx; g.direct_w = c->head_w + Lf.off + Ll.n_out; g.dy = dy;
            launch_head_gemm<false, true, HEAD_EPI_ADD_DIRECT>(c, g);
        }
        HeadGemmParams u{};
        u.A = dZ; u.lda = L.n_out; u.B = li > 0 ? (const float *)c->hH[li - 1].p : X; u.ldb = li > 0 ? L.n_in : c->ldx;
        u.M = L.n_out; u.N = L.n_in; u.K = rows; u.ldc = L.n_in;
        u.G1 = c->head_G1 + L.off; u.G2 = c->head_G2 + L.off;
        u.G1_bias = c->head_G1 + L.off + (size_t)L.n_in * L.n_out; u.G2_bias = c->head_G2 + L.off + (size_t)L.n_in * L.n_out;
        launch_head_gemm<true, true, HEAD_EPI_SUMS>(c, u, sums_stream);
    }
    if (fork) {
        CUDA_TRY(c, cudaEventRecord(c->ev_head_join, c->head_side_stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_head_join, 0));
    }
    k_head_apply<<<(uint32_t)std::min<size_t>((c->head_params + 255) / 256, (size_t)c->num_sms * 8), 256, 0, c->stream>>>(
        c->head_w, c->head_acc, c->head_G1, c->head_G2, c->head_params, c->optimizer, c->lut_dev + 2 * FWGPU_LUT_SIZE, c->d.nn_learning_rate, -c->d.nn_power_t);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return FWGPU_OK;
}

// Regressor::learn for a model with a dense head, over `count` examples: sub-batches of at most head_batch examples
// (the concurrency ramp shrinks the first ones) go through   LR/FFM forward -> head -> LR/FFM update.
// cta != nullptr: raw records through k_learn_fixed_cta, leftovers translated (tp_left) and run through k_learn;
// cta == nullptr: meta / lr_ent / ffm_ent already hold the `count` translated examples.
static fwgpu_status head_learn(fwgpu_ctx *c, uint32_t count, int update, uint32_t n_cap, uint32_t lr_cap, const FixedCtaParams *cta,
                               const TranslateParams *tp_left)
{
    fwgpu_status st;
    const uint32_t cap_rows = std::min<uint32_t>(c->head_batch, std::max<uint32_t>(count, 1));
    if ((st = ensure(c, c->hX, (size_t)cap_rows * c->ldx * 4))) return st;
    if ((st = ensure(c, c->hdX, (size_t)cap_rows * c->ldx * 4))) return st;
    for (size_t l = 0; l + 1 < c->head.size(); l++) {
        if ((st = ensure(c, c->hH[l], (size_t)cap_rows * c->head[l].n_out * 4))) return st;
        if ((st = ensure(c, c->hdZ[l], (size_t)cap_rows * c->head[l].n_out * 4))) return st;
    }
    if ((st = ensure(c, c->h_label, (size_t)cap_rows * 4))) return st;
    if ((st = ensure(c, c->h_imp, (size_t)cap_rows * 4))) return st;
    if ((st = ensure(c, c->h_outidx, (size_t)cap_rows * 4))) return st;
    if ((st = ensure(c, c->h_dy, (size_t)cap_rows * 4))) return st;
    LearnParams p;
    int T; size_t smem;
    if ((st = make_learn_params(c, n_cap, update, lr_cap, p, T, smem))) return st;
    const size_t smem_cta = (size_t)c->F * c->Fk * 4 + (size_t)c->F * 8 + 64;
    uint32_t done = 0;
    while (done < count) {
        uint32_t rows = std::min<uint32_t>(count - done, cap_rows);
        if (update && c->ramp_div != 0xffffffffu && !c->ramp_finished) {
            // Every example of a sub-batch moves EVERY dense weight, so aligned gradients add up: AdaGrad's combined step on a
            // weight is ~ lr * B / sqrt(n) after n examples and the layer's fan-in multiplies the curvature it acts on.  The
            // sub-batch therefore grows like sqrt(n) (on top of the linear cold-start ramp of the sparse tables), DESIGN.md.
            const uint64_t seen = c->examples_seen;
            uint64_t cap = std::max<uint64_t>(seen / c->ramp_div, 1);
            if (c->head_ramp_mul > 0.0) cap = std::min<uint64_t>(cap, std::max<uint64_t>((uint64_t)(c->head_ramp_mul * sqrt((double)seen)), 1));
            rows = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(rows, cap), std::max<uint64_t>(2 * seen, c->ramp_div) - seen);
            if (cap >= c->head_batch) c->ramp_finished = true;
        }
        if (update && c->max_inflight) rows = std::min(rows, c->max_inflight);
        HeadIO io{};
        io.X = (float *)c->hX.p; io.dX = (const float *)c->hdX.p; io.ldx = c->ldx; io.n_lr_out = c->d.num_combos;
        io.row_label = (float *)c->h_label.p; io.row_importance = (float *)c->h_imp.p; io.row_out_index = (uint32_t *)c->h_outidx.p;
        io.dy = (const float *)c->h_dy.p; io.row_base = done;
        LearnParams q = p;
        q.io = io; q.exact_order = 0; q.max_groups = 0;
        uint32_t full_groups = 0;
        for (int phase = 1; phase <= (update ? 2 : 1); phase++) {
            q.phase = phase; q.update = phase == 2 ? 1 : 0;
            if (cta) {
                FixedCtaParams cp = *cta;
                cp.ex_begin = done; cp.n_examples = rows; cp.io = io; cp.update = q.update; cp.max_groups = 0;
                cudaError_t e;
                {
                    ProfScope ps(c, 0);
                    if (phase == 1) {
                        CUDA_TRY(c, cudaMemsetAsync(cp.leftover_cnt, 0, 16, c->stream));
                        e = c->fast_rows ? launch_rows<1>(c, rows_params(c, cp), &full_groups) : launch_fixed_cta<2, 1>(c, cp, smem_cta, &full_groups);
                    } else e = c->fast_rows ? launch_rows<2>(c, rows_params(c, cp), &full_groups) : launch_fixed_cta<2, 2>(c, cp, smem_cta, &full_groups);
                }
                if (e != cudaSuccess) { c->set_error(std::string("k_learn_fixed_cta launch: ") + cudaGetErrorString(e)); return FWGPU_ERR_CUDA; }
                if (phase == 1) { // records the fused kernel cannot take: translate them once, both passes use the result
                    TranslateParams tp = *tp_left;
                    tp.n_examples = rows;
                    ProfScope ps(c, 1);
                    k_translate<<<(rows + 255) / 256, 256, 0, c->stream>>>(tp);
                    c->launches++;
                }
                q.meta = p.meta; q.n_examples = rows; q.n_examples_dev = cp.leftover_cnt;
            } else {
                q.meta = p.meta + done; q.n_examples = rows; q.n_examples_dev = nullptr;
            }
            if ((st = dispatch_learn(c, q, T, smem, &full_groups))) return st;
            if (phase == 1 && (st = head_pass(c, rows, update))) return st;
        }
        if (update) c->examples_seen += rows;
        done += rows;
    }
    return FWGPU_OK;
}

// ---- CSR batch entry --------------------------------------------------------------------------
static fwgpu_status learn_batch_impl(fwgpu_ctx *c, const fwgpu_batch *b, float *preds_out, int update)
{
    if (!c || !b) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (update && c->d.immutable) { c->set_error("This regressor is immutable, you cannot call learn() with update = true"); return FWGPU_ERR_IMMUTABLE; }
    const uint32_t n = b->n_examples;
    if (n == 0) return FWGPU_OK;
    if (!b->labels || !b->importance || !b->lr_off) { c->set_error("batch arrays missing"); return FWGPU_ERR_INVALID; }
    const bool has_ffm = c->F > 0 && b->ffm_off;
    const uint64_t n_lr = b->lr_off[n], n_ffm = has_ffm ? b->ffm_off[n] : 0;
    uint32_t max_ffm = 0;
    if (has_ffm) for (uint32_t i = 0; i < n; i++) max_ffm = std::max(max_ffm, b->ffm_off[i + 1] - b->ffm_off[i]);
    // one device slab: labels | importance | lr_off | lr_hash | lr_val | lr_combo | ffm_off | ffm_hash | ffm_val | ffm_field
    size_t off = 0;
    auto take = [&](size_t words) { size_t o = off; off += ((words + 3) / 4) * 4; return o; };
    size_t o_lab = take(n), o_imp = take(n), o_lro = take(n + 1), o_lrh = take(n_lr), o_lrv = take(n_lr), o_lrc = take(n_lr);
    size_t o_fo = take(n + 1), o_fh = take(n_ffm), o_fv = take(n_ffm), o_ff = take(n_ffm);
    fwgpu_status st;
    if ((st = ensure(c, c->csr, off * 4))) return st;
    if ((st = ensure(c, c->meta, (size_t)n * sizeof(ExMeta)))) return st;
    if ((st = ensure(c, c->lr_ent, std::max<size_t>(n_lr, 1) * 16))) return st;
    if ((st = ensure(c, c->ffm_ent, std::max<size_t>(n_ffm, 1) * 16))) return st;
    if ((st = ensure(c, c->preds, (size_t)n * 4))) return st;
    uint32_t *base = (uint32_t *)c->csr.p;
    auto up = [&](size_t o, const void *src, size_t words) -> cudaError_t {
        if (!words) return cudaSuccess;
        return cudaMemcpyAsync(base + o, src, words * 4, cudaMemcpyHostToDevice, c->stream);
    };
    CUDA_TRY(c, up(o_lab, b->labels, n));
    CUDA_TRY(c, up(o_imp, b->importance, n));
    CUDA_TRY(c, up(o_lro, b->lr_off, n + 1));
    CUDA_TRY(c, up(o_lrh, b->lr_hash, n_lr));
    CUDA_TRY(c, up(o_lrv, b->lr_val, n_lr));
    CUDA_TRY(c, up(o_lrc, b->lr_combo, n_lr));
    if (has_ffm) {
        CUDA_TRY(c, up(o_fo, b->ffm_off, n + 1));
        CUDA_TRY(c, up(o_fh, b->ffm_hash, n_ffm));
        CUDA_TRY(c, up(o_fv, b->ffm_val, n_ffm));
        CUDA_TRY(c, up(o_ff, b->ffm_field, n_ffm));
    }
    PackParams pp{};
    pp.n_examples = n;
    pp.labels = (const float *)(base + o_lab); pp.importance = (const float *)(base + o_imp);
    pp.lr_off = base + o_lro; pp.lr_hash = base + o_lrh; pp.lr_val = (const float *)(base + o_lrv); pp.lr_combo = base + o_lrc;
    pp.ffm_off = has_ffm ? base + o_fo : nullptr; pp.ffm_hash = base + o_fh; pp.ffm_val = (const float *)(base + o_fv); pp.ffm_field = base + o_ff;
    pp.n_lr = (uint32_t)n_lr; pp.n_ffm = (uint32_t)n_ffm;
    pp.meta = (ExMeta *)c->meta.p; pp.lr_ent = (uint4 *)c->lr_ent.p; pp.ffm_ent = (uint4 *)c->ffm_ent.p;
    uint64_t threads = std::max<uint64_t>(std::max<uint64_t>(n, n_lr), n_ffm);
    {
        ProfScope ps(c, 1);
        k_pack<<<(uint32_t)((threads + 255) / 256), 256, 0, c->stream>>>(pp);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    uint32_t n_cap = std::max(max_ffm, c->n_field_refs);
    uint32_t max_lr = 0;
    for (uint32_t i = 0; i < n; i++) max_lr = std::max(max_lr, b->lr_off[i + 1] - b->lr_off[i]);
    if ((st = launch_learn(c, n, n_cap, update, nullptr, true, max_lr))) return st;
    if (preds_out) CUDA_TRY(c, cudaMemcpyAsync(preds_out, c->preds.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_learn_batch(fwgpu_ctx *c, const fwgpu_batch *b, float *preds_out, int update) { return learn_batch_impl(c, b, preds_out, update); }
extern "C" fwgpu_status fwgpu_predict_batch(fwgpu_ctx *c, const fwgpu_batch *b, float *preds_out) { return learn_batch_impl(c, b, preds_out, 0); }

// ---- raw-record entries -----------------------------------------------------------------------
static uint32_t derive_lr_stride(const fwgpu_ctx *c, uint32_t max_dyn_pairs)
{
    if (c->d.max_lr_per_example) return c->d.max_lr_per_example;
    // every combo contributes prod(count(ns)); with single-valued slots that is <= 1 each
    uint64_t s = 0;
    const uint64_t m = std::max<uint32_t>(1, max_dyn_pairs);
    for (uint32_t i = 0; i + 1 < c->combo_off.size(); i++) {
        uint64_t t = 1;
        for (uint32_t j = c->combo_off[i]; j < c->combo_off[i + 1]; j++) t = std::min<uint64_t>(t * m, 1u << 16);
        s += t;
    }
    s += c->d.add_constant ? 1 : 0;
    return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(s, 1), 1u << 16);
}
static uint32_t derive_ffm_stride(const fwgpu_ctx *c, uint32_t max_dyn_pairs)
{
    if (c->d.max_ffm_per_example) return c->d.max_ffm_per_example;
    uint64_t s = (uint64_t)c->n_field_refs * std::max<uint32_t>(1, max_dyn_pairs);
    return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(s, 1), 1u << 16);
}

struct RecView {
    const uint32_t *dev_records; // device pointer to the first word of the slice
    const uint32_t *dev_rec_off; // device, [count+1] (absolute word offsets) or null
    uint32_t off_base;           // subtract from rec_off values
    uint32_t fixed_len;
    uint32_t max_len;            // longest record in words (bounds the dynamic part)
};

static fwgpu_status translate_and_learn(fwgpu_ctx *c, const RecView &rv, uint32_t count, float *preds_host, int update, bool run_learn,
                                        uint32_t *lr_stride_out = nullptr, uint32_t *ffm_stride_out = nullptr)
{
    const uint32_t hdr = 3 + c->d.n_namespaces;
    const uint32_t dyn_pairs = rv.max_len > hdr ? (rv.max_len - hdr + 1) / 2 : 0;
    const uint32_t lr_stride = derive_lr_stride(c, dyn_pairs);
    const uint32_t ffm_stride = c->F ? derive_ffm_stride(c, dyn_pairs) : 1;
    if (lr_stride_out) *lr_stride_out = lr_stride;
    if (ffm_stride_out) *ffm_stride_out = ffm_stride;
    fwgpu_status st;
    if ((st = ensure(c, c->meta, (size_t)count * sizeof(ExMeta)))) return st;
    if ((st = ensure(c, c->lr_ent, (size_t)count * lr_stride * 16))) return st;
    if ((st = ensure(c, c->ffm_ent, (size_t)count * ffm_stride * 16))) return st;
    if ((st = ensure(c, c->preds, (size_t)count * 4))) return st;
    TranslateParams tp{};
    tp.records = rv.dev_records; tp.rec_off = rv.dev_rec_off; tp.fixed_len = rv.fixed_len; tp.n_examples = count;
    tp.off_base = rv.off_base;
    tp.n_namespaces = c->d.n_namespaces; tp.ns_is_f32 = c->d_ns_is_f32;
    tp.n_combos = c->d.n_combos; tp.combo_off = c->d_combo_off; tp.combo_ns = c->d_combo_ns; tp.combo_weight = c->d_combo_weight;
    tp.add_constant = c->d.add_constant; tp.n_fields = c->F; tp.field_off = c->d_field_off; tp.field_ns = c->d_field_ns;
    tp.lr_mask = (uint32_t)((1ull << c->d.bit_precision) - 1); // feature_buffer.rs:140
    uint32_t bits = 0;
    while (c->d.ffm_k > (1u << bits)) bits++;                  // feature_buffer.rs:142-148
    tp.ffm_mask = c->d.ffm_k ? (uint32_t)(((1ull << c->d.ffm_bit_precision) - 1) ^ ((1u << bits) - 1)) : 0;
    tp.ffm_k = c->d.ffm_k;
    tp.lr_stride = lr_stride; tp.ffm_stride = ffm_stride;
    tp.meta = (ExMeta *)c->meta.p; tp.lr_ent = (uint4 *)c->lr_ent.p; tp.ffm_ent = (uint4 *)c->ffm_ent.p; tp.err_flag = c->err_flag;
    if (run_learn && !c->head.empty()) {
        // dense head: two passes around the head's GEMMs per sub-batch (head_learn)
        const bool use_cta = c->fast_ok && c->fast_cta && c->fast_enabled;
        if (use_cta) {
            if ((st = ensure(c, c->leftover, (size_t)(count + 4) * 4))) return st;
            uint32_t *left_cnt = (uint32_t *)c->leftover.p, *left_idx = left_cnt + 4;
            FixedCtaParams cp{};
            cp.lr = c->lr; cp.ffm_w = c->ffm_w; cp.ffm_acc = c->ffm_acc; cp.lut_lr = c->lut_dev; cp.lut_ffm = c->lut_dev + FWGPU_LUT_SIZE;
            cp.records = rv.dev_records; cp.rec_off = rv.dev_rec_off; cp.off_base = rv.off_base; cp.fixed_len = rv.fixed_len;
            cp.F = c->F; cp.k = c->k; cp.Fk = c->Fk; cp.cpr = c->Fk / 4;
            cp.div_cpr = make_fastdiv(std::max<uint32_t>(cp.cpr, 1)); cp.div_k4 = make_fastdiv(std::max<uint32_t>(c->k / 4, 1)); cp.div_F = make_fastdiv(std::max<uint32_t>(c->F, 1));
            cp.field_ns = c->d_field_ns; cp.n_combos = c->d.n_combos; cp.combo_off = c->d_combo_off; cp.combo_ns = c->d_combo_ns;
            cp.combo_weight = c->d_combo_weight; cp.add_constant = c->d.add_constant; cp.lr_mask = tp.lr_mask; cp.ffm_mask = tp.ffm_mask;
            cp.optimizer = c->optimizer; cp.lr_lr = c->d.learning_rate; cp.lr_mpt = -c->d.power_t; cp.ffm_lr = c->d.ffm_learning_rate; cp.ffm_mpt = -c->d.ffm_power_t;
            cp.preds = (float *)c->preds.p; cp.leftover_idx = left_idx; cp.leftover_cnt = left_cnt;
            tp.ex_list = left_idx; tp.ex_count = left_cnt;
            if ((st = head_learn(c, count, update, ffm_stride, lr_stride, &cp, &tp))) return st;
        } else {
            {
                ProfScope ps(c, 1);
                k_translate<<<(count + 255) / 256, 256, 0, c->stream>>>(tp);
                c->launches++;
            }
            CUDA_TRY(c, cudaGetLastError());
            if ((st = head_learn(c, count, update, ffm_stride, lr_stride, nullptr, nullptr))) return st;
        }
        if (preds_host) CUDA_TRY(c, cudaMemcpyAsync(preds_host, c->preds.p, (size_t)count * 4, cudaMemcpyDeviceToHost, c->stream));
        return FWGPU_OK;
    }
    // hogwild_ramp_div 0x7fffffff .. 0xfffffffe = sequential mode (general kernel, reference tape order); 0xffffffff = no ramp
    const bool use_fast = run_learn && c->fast_ok && c->fast_enabled && (c->ramp_div < 0x7fffffffu || c->ramp_div == 0xffffffffu);
    if (use_fast) {
        // fused kernel on the raw records; records it cannot take are listed and go through the general path below
        if ((st = ensure(c, c->leftover, (size_t)(count + 4) * 4))) return st;
        uint32_t *left_cnt = (uint32_t *)c->leftover.p, *left_idx = left_cnt + 4;
        CUDA_TRY(c, cudaMemsetAsync(left_cnt, 0, 16, c->stream));
        FixedParams fp{};
        fp.lr = c->lr; fp.ffm_w = c->ffm_w; fp.ffm_acc = c->ffm_acc; fp.lut_lr = c->lut_dev; fp.lut_ffm = c->lut_dev + FWGPU_LUT_SIZE;
        fp.records = rv.dev_records; fp.rec_off = rv.dev_rec_off; fp.off_base = rv.off_base; fp.fixed_len = rv.fixed_len;
        fp.F = c->F; fp.k = c->k; fp.cpr = c->Fk / 4;
        fp.div_cpr = make_fastdiv(std::max<uint32_t>(fp.cpr, 1)); fp.div_k4 = make_fastdiv(std::max<uint32_t>(c->k / 4, 1));
        fp.field_ns = c->d_field_ns;
        fp.n_combos = c->d.n_combos; fp.combo_off = c->d_combo_off; fp.combo_ns = c->d_combo_ns; fp.combo_weight = c->d_combo_weight;
        fp.add_constant = c->d.add_constant; fp.lr_mask = tp.lr_mask; fp.ffm_mask = tp.ffm_mask;
        fp.optimizer = c->optimizer; fp.lr_lr = c->d.learning_rate; fp.lr_mpt = -c->d.power_t; fp.ffm_lr = c->d.ffm_learning_rate; fp.ffm_mpt = -c->d.ffm_power_t;
        fp.update = update; fp.preds = (float *)c->preds.p; fp.leftover_idx = left_idx; fp.leftover_cnt = left_cnt;
        fp.rec_smem_floats = c->F * (fp.cpr + 1) * 4;
        fp.sys_scope = c->shard ? 1 : 0;
        uint32_t done = 0;
        while (done < count) {
            uint32_t cnt = count - done, cap = 0;
            if (update && c->ramp_div != 0xffffffffu && !c->ramp_finished) {
                const uint64_t seen = c->examples_seen;
                cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(seen / c->ramp_div, 1), 1u << 30);
                cnt = (uint32_t)std::min<uint64_t>(cnt, std::max<uint64_t>(2 * seen, c->ramp_div) - seen);
            }
            if (update && c->max_inflight && (cap == 0 || cap > c->max_inflight)) cap = c->max_inflight;
            // one model over several GPUs: gradients go to the rows' owners chunk by chunk (one record in flight = the parity
            // mode keeps the direct path: its remote bulk reductions complete before the next record gathers)
            const bool push = c->push_ok && update && c->fast_cta && c->fast_rows && cap != 1;
            if (push) cnt = std::min<uint32_t>(cnt, c->shard_chunk);
            fp.ex_begin = done; fp.n_examples = cnt; fp.max_groups = cap;
            {   // Combining the bias updates of 16 records per group multiplies the number of records whose bias gradients are
                // applied from one stale value by 16 (14 208 groups x 16 = 227 K records on c2).  On a young model (small
                // accumulator, large steps) that made the bias oscillate between 2 M and 4 M examples: 0.4-2 % of progressive
                // logloss on 10^7 examples, gone entirely with per-record bias updates whatever the ramp divisor
                // (profiles/r02_c2_ramp_sweep.txt).  It therefore starts once the model has seen 2^24 examples.
                static const long bias_env = getenv("FWGPU_BIAS_PERIOD") ? atol(getenv("FWGPU_BIAS_PERIOD")) : -1;
                fp.bias_combine = bias_env >= 0 ? (bias_env > 1) : (c->examples_seen >= (1ull << 24));
            }
            uint32_t full_groups = 0;
            cudaError_t e;
            if (c->fast_cta) {
                FixedCtaParams cp{};
                cp.lr = fp.lr; cp.ffm_w = fp.ffm_w; cp.ffm_acc = fp.ffm_acc; cp.lut_lr = fp.lut_lr; cp.lut_ffm = fp.lut_ffm;
                cp.records = fp.records; cp.rec_off = fp.rec_off; cp.off_base = fp.off_base; cp.fixed_len = fp.fixed_len;
                cp.ex_begin = done; cp.n_examples = cnt;
                cp.F = c->F; cp.k = c->k; cp.Fk = c->Fk; cp.cpr = c->Fk / 4;
                cp.div_cpr = fp.div_cpr; cp.div_k4 = fp.div_k4; cp.div_F = make_fastdiv(std::max<uint32_t>(c->F, 1));
                cp.field_ns = fp.field_ns; cp.n_combos = fp.n_combos; cp.combo_off = fp.combo_off; cp.combo_ns = fp.combo_ns;
                cp.combo_weight = fp.combo_weight; cp.add_constant = fp.add_constant; cp.lr_mask = fp.lr_mask; cp.ffm_mask = fp.ffm_mask;
                cp.optimizer = fp.optimizer; cp.lr_lr = fp.lr_lr; cp.lr_mpt = fp.lr_mpt; cp.ffm_lr = fp.ffm_lr; cp.ffm_mpt = fp.ffm_mpt;
                cp.update = update; cp.preds = fp.preds; cp.leftover_idx = left_idx; cp.leftover_cnt = left_cnt; cp.max_groups = cap;
                const size_t smem_cta = (size_t)c->F * c->Fk * 4 + (size_t)c->F * 8 + 64;
                ProfScope ps(c, 0);
                if (push) {
                    if ((st = shard_push_chunk(c, rows_params(c, cp), &full_groups))) return st;
                    e = cudaSuccess;
                } else if (c->fast_rows) e = launch_rows<0>(c, rows_params(c, cp), &full_groups);
                else switch (c->fast_ub) {
                case 1: e = launch_fixed_cta<1>(c, cp, smem_cta, &full_groups); break;
                case 4: e = launch_fixed_cta<4>(c, cp, smem_cta, &full_groups); break;
                default: e = launch_fixed_cta<2>(c, cp, smem_cta, &full_groups); break;
                }
            } else {
                ProfScope ps(c, 0);
                e = launch_fixed(c, fp, &full_groups);
            }
            if (e != cudaSuccess) { c->set_error(std::string("k_learn_fixed launch: ") + cudaGetErrorString(e)); return FWGPU_ERR_CUDA; }
            if (update) {
                c->examples_seen += cnt;
                if (cap && full_groups && cap >= full_groups) c->ramp_finished = true;
            }
            done += cnt;
        }
        tp.ex_list = left_idx; tp.ex_count = left_cnt;
        // sharded owner-side path: the call is complete, in stream order, when the owners have applied its last chunk here
        if (c->push_ok && c->shard_overlap && c->shard_chunks_done) {
            const int last = (int)((c->shard_chunks_done - 1) & 1);
            if (c->ev_applied_rec[last]) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_applied[last], 0));
        }
    }
    {
        ProfScope ps(c, 1);
        k_translate<<<(count + 255) / 256, 256, 0, c->stream>>>(tp);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    if (!run_learn) return FWGPU_OK;
    if ((st = launch_learn(c, count, ffm_stride, update, use_fast ? tp.ex_count : nullptr, !use_fast, lr_stride))) return st;
    if (preds_host) CUDA_TRY(c, cudaMemcpyAsync(preds_host, c->preds.p, (size_t)count * 4, cudaMemcpyDeviceToHost, c->stream));
    return FWGPU_OK;
}

static size_t chunk_examples(const fwgpu_ctx *c, uint32_t lr_stride, uint32_t ffm_stride)
{
    size_t slab = (size_t)(lr_stride + ffm_stride) * 16 + sizeof(ExMeta) + 4;
    size_t target = (size_t)384 << 20;
    if (const char *t = getenv("FWGPU_CHUNK_MB")) target = (size_t)atol(t) << 20;
    size_t n = target / slab;
    return std::max<size_t>(n, 4096);
}

static uint32_t host_max_len(const uint32_t *rec_off, uint64_t first, uint64_t count)
{
    uint32_t m = 0;
    for (uint64_t i = first; i < first + count; i++) m = std::max(m, rec_off[i + 1] - rec_off[i]);
    return m;
}

extern "C" fwgpu_status fwgpu_learn_records(fwgpu_ctx *c, const uint32_t *records, uint64_t n_words, const uint32_t *rec_off,
                                            uint32_t n_examples, float *preds_out, int update)
{
    if (!c || (!records && n_examples)) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (update && c->d.immutable) { c->set_error("This regressor is immutable, you cannot call learn() with update = true"); return FWGPU_ERR_IMMUTABLE; }
    if (n_examples == 0) return FWGPU_OK;
    if (c->d.n_namespaces == 0) { c->set_error("model descriptor has no translate spec (n_namespaces == 0)"); return FWGPU_ERR_INVALID; }
    uint32_t fixed_len = 0;
    if (!rec_off) {
        if (n_words % n_examples) { c->set_error("rec_off == NULL needs fixed-length records"); return FWGPU_ERR_INVALID; }
        fixed_len = (uint32_t)(n_words / n_examples);
    }
    const uint32_t hdr = 3 + c->d.n_namespaces;
    const uint32_t max_len = rec_off ? host_max_len(rec_off, 0, n_examples) : fixed_len;
    if (max_len < hdr && !rec_off) { c->set_error("records shorter than header + namespace slots"); return FWGPU_ERR_INVALID; }
    const uint32_t dyn_pairs = max_len > hdr ? (max_len - hdr + 1) / 2 : 0;
    // ~1.3 M records per chunk on c2 (57 MB of records): measured best for the host-record pipeline; 16 MB chunks lose 9 % to launch tails
    const size_t chunk = chunk_examples(c, derive_lr_stride(c, dyn_pairs), c->F ? derive_ffm_stride(c, dyn_pairs) : 1);
    fwgpu_status st;
    uint64_t done = 0;
    int bi = 0;
    while (done < n_examples) {
        const uint32_t cnt = (uint32_t)std::min<uint64_t>(chunk, n_examples - done);
        const uint64_t w0 = rec_off ? rec_off[done] : done * fixed_len;
        const uint64_t w1 = rec_off ? rec_off[done + cnt] : (done + cnt) * fixed_len;
        // copy stream: wait until the translate of the chunk that last used this buffer has finished
        if ((st = ensure(c, c->rec[bi], (size_t)(w1 - w0) * 4))) return st;
        if (rec_off && (st = ensure(c, c->rec_off_dev[bi], (size_t)(cnt + 1) * 4))) return st;
        if (c->ev_free_recorded[bi]) CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_free[bi], 0));
        CUDA_TRY(c, cudaMemcpyAsync(c->rec[bi].p, records + w0, (size_t)(w1 - w0) * 4, cudaMemcpyHostToDevice, c->copy_stream));
        if (rec_off) CUDA_TRY(c, cudaMemcpyAsync(c->rec_off_dev[bi].p, rec_off + done, (size_t)(cnt + 1) * 4, cudaMemcpyHostToDevice, c->copy_stream));
        CUDA_TRY(c, cudaEventRecord(c->ev_ready[bi], c->copy_stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_ready[bi], 0));
        RecView rv{(const uint32_t *)c->rec[bi].p, rec_off ? (const uint32_t *)c->rec_off_dev[bi].p : nullptr, (uint32_t)w0, fixed_len, max_len};
        if ((st = translate_and_learn(c, rv, cnt, nullptr, update, true))) return st;
        CUDA_TRY(c, cudaEventRecord(c->ev_free[bi], c->stream)); // translate (and learn) of this chunk are ordered before it
        c->ev_free_recorded[bi] = true;
        if (preds_out) {
            if ((st = ensure(c, c->pred_stage[bi], (size_t)cnt * 4))) return st;
            if (c->ev_pred_out_recorded[bi]) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_pred_out[bi], 0)); // staging buffer free again
            CUDA_TRY(c, cudaMemcpyAsync(c->pred_stage[bi].p, c->preds.p, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(c, cudaEventRecord(c->ev_pred_ready[bi], c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_pred_ready[bi], 0));
            CUDA_TRY(c, cudaMemcpyAsync(preds_out + done, c->pred_stage[bi].p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, c->d2h_stream));
            CUDA_TRY(c, cudaEventRecord(c->ev_pred_out[bi], c->d2h_stream));
            c->ev_pred_out_recorded[bi] = true;
        }
        done += cnt;
        bi ^= 1;
    }
    // the call is complete, in stream order, when its last predictions have left: whoever waits for (or records an event on) the
    // compute stream afterwards also waits for them
    if (preds_out)
        for (int i = 0; i < 2; i++) if (c->ev_pred_out_recorded[i]) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_pred_out[i], 0));
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_translate_records(fwgpu_ctx *c, const uint32_t *records, uint64_t n_words, const uint32_t *rec_off, uint32_t n,
                                                float *labels, float *importance, uint32_t *lr_off, uint32_t *lr_hash, float *lr_val,
                                                uint32_t *lr_combo, uint64_t lr_cap, uint32_t *ffm_off, uint32_t *ffm_hash, float *ffm_val,
                                                uint32_t *ffm_field, uint64_t ffm_cap)
{
    if (!c || !records || !lr_off || !ffm_off) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (n == 0) { lr_off[0] = 0; ffm_off[0] = 0; return FWGPU_OK; }
    uint32_t fixed_len = 0;
    if (!rec_off) {
        if (n_words % n) { c->set_error("rec_off == NULL needs fixed-length records"); return FWGPU_ERR_INVALID; }
        fixed_len = (uint32_t)(n_words / n);
    }
    const uint32_t max_len = rec_off ? host_max_len(rec_off, 0, n) : fixed_len;
    fwgpu_status st;
    if ((st = ensure(c, c->rec[0], (size_t)n_words * 4))) return st;
    CUDA_TRY(c, cudaMemcpyAsync(c->rec[0].p, records, (size_t)n_words * 4, cudaMemcpyHostToDevice, c->stream));
    if (rec_off) {
        if ((st = ensure(c, c->rec_off_dev[0], (size_t)(n + 1) * 4))) return st;
        CUDA_TRY(c, cudaMemcpyAsync(c->rec_off_dev[0].p, rec_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    }
    RecView rv{(const uint32_t *)c->rec[0].p, rec_off ? (const uint32_t *)c->rec_off_dev[0].p : nullptr, rec_off ? rec_off[0] : 0, fixed_len, max_len};
    uint32_t lr_stride = 0, ffm_stride = 0;
    if ((st = translate_and_learn(c, rv, n, nullptr, 0, false, &lr_stride, &ffm_stride))) return st;
    std::vector<ExMeta> meta(n);
    std::vector<uint4> lre((size_t)n * lr_stride), ffe((size_t)n * ffm_stride);
    CUDA_TRY(c, cudaMemcpyAsync(meta.data(), c->meta.p, (size_t)n * sizeof(ExMeta), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(lre.data(), c->lr_ent.p, lre.size() * 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(ffe.data(), c->ffm_ent.p, ffe.size() * 16, cudaMemcpyDeviceToHost, c->stream));
    if ((st = fwgpu_sync(c))) return st;
    uint64_t nl = 0, nf = 0;
    for (uint32_t i = 0; i < n; i++) {
        labels[i] = meta[i].label;
        importance[i] = meta[i].importance;
        lr_off[i] = (uint32_t)nl;
        ffm_off[i] = (uint32_t)nf;
        if (nl + meta[i].lr_cnt > lr_cap || nf + meta[i].ffm_cnt > ffm_cap) { c->set_error("output capacity too small"); return FWGPU_ERR_INVALID; }
        for (uint32_t j = 0; j < meta[i].lr_cnt; j++) {
            const uint4 &e = lre[(size_t)meta[i].lr_begin + j];
            lr_hash[nl] = e.x; memcpy(&lr_val[nl], &e.y, 4); lr_combo[nl] = e.z; nl++;
        }
        for (uint32_t j = 0; j < meta[i].ffm_cnt; j++) {
            const uint4 &e = ffe[(size_t)meta[i].ffm_begin + j];
            ffm_hash[nf] = e.x; memcpy(&ffm_val[nf], &e.y, 4); ffm_field[nf] = e.z; nf++;
        }
    }
    lr_off[n] = (uint32_t)nl;
    ffm_off[n] = (uint32_t)nf;
    return FWGPU_OK;
}

// ---- resident dataset -------------------------------------------------------------------------
extern "C" fwgpu_status fwgpu_dataset_upload(fwgpu_ctx *c, const uint32_t *records, uint64_t n_words, const uint32_t *rec_off,
                                             uint64_t n_examples, fwgpu_dataset **out)
{
    if (!c || !records || !out || n_examples == 0) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->d.n_namespaces == 0) { c->set_error("model descriptor has no translate spec (n_namespaces == 0)"); return FWGPU_ERR_INVALID; }
    fwgpu_dataset *ds = new fwgpu_dataset();
    ds->n_words = n_words; ds->n_examples = n_examples;
    if (!rec_off) {
        if (n_words % n_examples) { delete ds; c->set_error("rec_off == NULL needs fixed-length records"); return FWGPU_ERR_INVALID; }
        ds->fixed_len = (uint32_t)(n_words / n_examples);
        ds->max_len = ds->fixed_len;
    } else {
        if (n_words >= (1ull << 32)) { delete ds; c->set_error("variable-length datasets are limited to 2^32 words"); return FWGPU_ERR_UNSUPPORTED; }
        ds->max_len = host_max_len(rec_off, 0, n_examples);
    }
    cudaError_t e = cudaMalloc((void **)&ds->records, n_words * 4);
    if (e == cudaSuccess) e = cudaMemcpy(ds->records, records, n_words * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && rec_off) {
        e = cudaMalloc((void **)&ds->rec_off, (n_examples + 1) * 4);
        if (e == cudaSuccess) e = cudaMemcpy(ds->rec_off, rec_off, (n_examples + 1) * 4, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(ds->records); cudaFree(ds->rec_off); delete ds;
        c->set_error(std::string("dataset upload: ") + cudaGetErrorString(e));
        return FWGPU_ERR_CUDA;
    }
    *out = ds;
    return FWGPU_OK;
}

extern "C" void fwgpu_dataset_free(fwgpu_ctx *c, fwgpu_dataset *ds)
{
    if (!ds) return;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    cudaFree(ds->records); cudaFree(ds->rec_off);
    delete ds;
}

extern "C" fwgpu_status fwgpu_dataset_learn(fwgpu_ctx *c, fwgpu_dataset *ds, uint64_t first, uint64_t count, float *preds_out, int update)
{
    if (!c || !ds || first + count > ds->n_examples) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (update && c->d.immutable) { c->set_error("This regressor is immutable, you cannot call learn() with update = true"); return FWGPU_ERR_IMMUTABLE; }
    const uint32_t hdr = 3 + c->d.n_namespaces;
    const uint32_t dyn_pairs = ds->max_len > hdr ? (ds->max_len - hdr + 1) / 2 : 0;
    const size_t chunk = chunk_examples(c, derive_lr_stride(c, dyn_pairs), c->F ? derive_ffm_stride(c, dyn_pairs) : 1);
    fwgpu_status st;
    uint64_t done = 0;
    while (done < count) {
        const uint32_t cnt = (uint32_t)std::min<uint64_t>(chunk, count - done);
        const uint64_t e0 = first + done;
        RecView rv;
        if (ds->rec_off) rv = RecView{ds->records, ds->rec_off + e0, 0, 0, ds->max_len};
        else rv = RecView{ds->records + e0 * ds->fixed_len, nullptr, 0, ds->fixed_len, ds->max_len};
        if ((st = translate_and_learn(c, rv, cnt, preds_out ? preds_out + done : nullptr, update, true))) return st;
        done += cnt;
    }
    return FWGPU_OK;
}

// ---- weights ----------------------------------------------------------------------------------
// Device -> host copy of the first `bytes` of a table; a sharded table is copied owner range by owner range (a copy never
// spans two physical allocations), remote ranges coming over NVLink.
static cudaError_t table_d2h(fwgpu_ctx *c, void *dst, const void *src, size_t bytes, const ShardedArray &a)
{
    if (!c->shard || !a.va) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
    for (uint32_t s = 0; s < c->shard->world; s++) {
        const size_t lo = a.offsets[s], hi = std::min(lo + a.sizes[s], bytes);
        if (hi <= lo) continue;
        cudaError_t e = cudaMemcpyAsync((char *)dst + lo, (const char *)src + lo, hi - lo, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
// Host -> device copy of a whole-table payload: every rank of a sharded model is handed the same payload and writes the
// range it owns (collective call; follow it with fwgpu_shard_barrier before training).
static void own_range(const fwgpu_ctx *c, const ShardedArray &a, size_t bytes, size_t &lo, size_t &hi)
{
    lo = 0; hi = bytes;
    if (c->shard && a.va) { lo = std::min(a.offsets[c->shard->rank], bytes); hi = std::min(a.offsets[c->shard->rank] + a.sizes[c->shard->rank], bytes); }
}
static cudaError_t table_h2d(fwgpu_ctx *c, void *dst, const void *src, size_t bytes, const ShardedArray &a)
{
    size_t lo, hi;
    own_range(c, a, bytes, lo, hi);
    if (hi <= lo) return cudaSuccess;
    return cudaMemcpyAsync((char *)dst + lo, (const char *)src + lo, hi - lo, cudaMemcpyHostToDevice, c->stream);
}

extern "C" fwgpu_status fwgpu_block_len(const fwgpu_ctx *c, int block, uint64_t *n_weights, uint64_t *n_bytes)
{
    if (!c) return FWGPU_ERR_INVALID;
    const bool sgd = c->optimizer == FWGPU_OPT_SGD;
    uint64_t n = 0, bytes = 0;
    if (block == FWGPU_BLOCK_LR) { n = c->lr_len; bytes = n * (sgd ? 4 : 8); }
    else if (block == FWGPU_BLOCK_FFM) { n = c->ffm_len; bytes = n * (sgd ? 4 : 8); }
    else if (block >= FWGPU_BLOCK_NN0 && (size_t)(block - FWGPU_BLOCK_NN0) < c->head.size()) {
        const auto &L = c->head[block - FWGPU_BLOCK_NN0];
        n = (uint64_t)(L.n_in + 1) * L.n_out; bytes = n * (sgd ? 4 : 8); // block_neural.rs:414-438
    }
    else return FWGPU_ERR_INVALID;
    if (n_weights) *n_weights = n;
    if (n_bytes) *n_bytes = bytes;
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_export_block(fwgpu_ctx *c, int block, void *dst, uint64_t dst_bytes)
{
    uint64_t n = 0, bytes = 0;
    if (!c || !dst || fwgpu_block_len(c, block, &n, &bytes) != FWGPU_OK) return FWGPU_ERR_INVALID;
    if (dst_bytes < bytes) { c->set_error("export buffer too small"); return FWGPU_ERR_INVALID; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const bool sgd = c->optimizer == FWGPU_OPT_SGD;
    if (block == FWGPU_BLOCK_LR) {
        if (!sgd) { // {f32 w, f32 acc} x len, block_helpers.rs:23-28
            CUDA_TRY(c, table_d2h(c, dst, c->lr, bytes, c->sh_lr));
        } else {
            fwgpu_status st;
            if ((st = ensure(c, c->csr, n * 4))) return st;
            k_lr_extract_w<<<c->num_sms * 4, 256, 0, c->stream>>>(c->lr, (float *)c->csr.p, n);
            c->launches++;
            CUDA_TRY(c, cudaMemcpyAsync(dst, c->csr.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
        }
    } else if (block >= FWGPU_BLOCK_NN0) { // weights, then accumulators (block_neural.rs:426-438)
        const auto &L = c->head[block - FWGPU_BLOCK_NN0];
        CUDA_TRY(c, cudaMemcpyAsync(dst, c->head_w + L.off, n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (!sgd) CUDA_TRY(c, cudaMemcpyAsync((char *)dst + n * 4, c->head_acc + L.off, n * 4, cudaMemcpyDeviceToHost, c->stream));
    } else {
        if (n == 0) return FWGPU_OK;
        CUDA_TRY(c, table_d2h(c, dst, c->ffm_w, n * 4, c->sh_w)); // weights, then accumulators (block_ffm.rs:835-848)
        if (!sgd) CUDA_TRY(c, table_d2h(c, (char *)dst + n * 4, c->ffm_acc, n * 4, c->sh_acc));
    }
    return fwgpu_sync(c);
}

extern "C" fwgpu_status fwgpu_import_block(fwgpu_ctx *c, int block, const void *src, uint64_t src_bytes, int with_optimizer_state)
{
    uint64_t n = 0, bytes = 0;
    if (!c || !src || fwgpu_block_len(c, block, &n, &bytes) != FWGPU_OK) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const bool sgd = c->optimizer == FWGPU_OPT_SGD;
    const bool with_acc = with_optimizer_state && !sgd;
    const uint64_t need = n * (with_acc ? 8 : 4);
    if (src_bytes < need) { c->set_error("import payload too small"); return FWGPU_ERR_INVALID; }
    if (with_acc) { c->examples_seen = std::max<uint64_t>(c->examples_seen, 1ull << 40); c->ramp_finished = true; }
    if (block == FWGPU_BLOCK_LR) {
        if (with_acc) CUDA_TRY(c, table_h2d(c, c->lr, src, n * 8, c->sh_lr));
        else {
            fwgpu_status st;
            if ((st = ensure(c, c->csr, n * 4))) return st;
            CUDA_TRY(c, cudaMemcpyAsync(c->csr.p, src, n * 4, cudaMemcpyHostToDevice, c->stream));
            const float acc0 = c->optimizer == FWGPU_OPT_ADAGRAD_FLEX ? c->d.init_acc_gradient : 0.0f;
            size_t lo, hi;
            own_range(c, c->sh_lr, n * 8, lo, hi);
            if (hi > lo) {
                k_lr_set_w<<<c->num_sms * 4, 256, 0, c->stream>>>(c->lr + lo / 8, (const float *)c->csr.p + lo / 8, (hi - lo) / 8, acc0);
                c->launches++;
            }
        }
    } else if (block >= FWGPU_BLOCK_NN0) { // block_neural.rs:440-470
        const auto &L = c->head[block - FWGPU_BLOCK_NN0];
        CUDA_TRY(c, cudaMemcpyAsync(c->head_w + L.off, src, n * 4, cudaMemcpyHostToDevice, c->stream));
        if (with_acc) CUDA_TRY(c, cudaMemcpyAsync(c->head_acc + L.off, (const char *)src + n * 4, n * 4, cudaMemcpyHostToDevice, c->stream));
        else if (!sgd) {
            const float acc0 = c->optimizer == FWGPU_OPT_ADAGRAD_FLEX ? c->d.nn_init_acc_gradient : 0.0f;
            k_fill<<<c->num_sms, 256, 0, c->stream>>>(c->head_acc + L.off, n, acc0);
            c->launches++;
        }
    } else {
        if (n == 0) return FWGPU_OK;
        CUDA_TRY(c, table_h2d(c, c->ffm_w, src, n * 4, c->sh_w));
        if (with_acc) CUDA_TRY(c, table_h2d(c, c->ffm_acc, (const char *)src + n * 4, n * 4, c->sh_acc));
        else if (!sgd) {
            const float acc0 = c->optimizer == FWGPU_OPT_ADAGRAD_FLEX ? c->d.ffm_init_acc_gradient : 0.0f;
            size_t lo, hi;
            own_range(c, c->sh_acc, n * 4, lo, hi);
            if (hi > lo) {
                k_fill<<<c->num_sms * 4, 256, 0, c->stream>>>(c->ffm_acc + lo / 4, (hi - lo) / 4, acc0);
                c->launches++;
            }
        }
    }
    return fwgpu_sync(c);
}

extern "C" fwgpu_status fwgpu_debug_logistic(fwgpu_ctx *c, const float *in, float *out, uint64_t n)
{
    if (!c || !in || !out) return FWGPU_ERR_INVALID;
    if (n == 0) return FWGPU_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    fwgpu_status st;
    if ((st = ensure(c, c->csr, n * 8))) return st;
    float *din = (float *)c->csr.p, *dout = din + n;
    CUDA_TRY(c, cudaMemcpyAsync(din, in, n * 4, cudaMemcpyHostToDevice, c->stream));
    k_debug_logistic<<<c->num_sms * 4, 256, 0, c->stream>>>(din, dout, n);
    c->launches++;
    CUDA_TRY(c, cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, c->stream));
    return fwgpu_sync(c);
}

extern "C" fwgpu_status fwgpu_debug_path_counts(fwgpu_ctx *c, uint64_t *out4)
{
    if (!c || !out4) return FWGPU_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    unsigned long long ex = 0;
    CUDA_TRY(c, cudaMemcpy(&ex, c->stat_general_examples, 8, cudaMemcpyDeviceToHost));
    out4[0] = c->n_fixed; out4[1] = c->n_fixed_cta; out4[2] = c->n_general; out4[3] = ex;
    return FWGPU_OK;
}

extern "C" fwgpu_status fwgpu_set_examples_seen(fwgpu_ctx *c, uint64_t n)
{
    if (!c) return FWGPU_ERR_INVALID;
    c->examples_seen = n;
    c->ramp_finished = false;
    return FWGPU_OK;
}
extern "C" uint64_t fwgpu_get_examples_seen(const fwgpu_ctx *c) { return c ? c->examples_seen : 0; }

extern "C" fwgpu_status fwgpu_get_lut(const fwgpu_ctx *c, int which, float *dst)
{
    if (!c || !dst || which < 0 || which > 2) return FWGPU_ERR_INVALID;
    memcpy(dst, c->lut_host[which], sizeof(float) * FWGPU_LUT_SIZE);
    return FWGPU_OK;
}
