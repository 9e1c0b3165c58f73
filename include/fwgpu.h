/*
 * fwgpu.h -- C ABI of the B200-native (sm_100a) replacement for Fwumious Wabbit's
 * LR / FFM learn / predict hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point
 * cites the reference interface it replaces (file:line under the reference's src/).  The
 * reference has no FFI for training (its only extern "C" surface is predict-only, lib.rs:151-235);
 * these are the functions a Rust host would bind with `extern "C"` in place of the trait-object
 * calls it makes today -- INTEGRATION.md shows that binding.
 *
 * Conventions (modelled on lib.rs:47-48, 237-243 and SURVEY.md section 8b):
 *   - every function returns fwgpu_status (0 = ok, < 0 = error); nothing unwinds across the ABI;
 *     fwgpu_last_error(ctx) returns a human-readable message for the last failure on that ctx;
 *   - one ctx per GPU; calls on one ctx are serialised on its compute stream, and are
 *     asynchronous with respect to the host: results (predictions written to host buffers,
 *     exported weights) are valid after fwgpu_sync() returns;
 *   - caller-owned host buffers must stay alive and unmodified until the next fwgpu_sync();
 *     pinned buffers (fwgpu_host_alloc) make the copies truly asynchronous;
 *   - no CPU fallback exists: without a CUDA device fwgpu_create fails with FWGPU_ERR_CUDA.
 */
#ifndef FWGPU_H
#define FWGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t fwgpu_status;
enum {
    FWGPU_OK = 0,
    FWGPU_ERR_INVALID = -1,     /* bad argument / descriptor                                  */
    FWGPU_ERR_CUDA = -2,        /* CUDA runtime error (message in fwgpu_last_error)            */
    FWGPU_ERR_UNSUPPORTED = -3, /* configuration outside what the kernels implement            */
    FWGPU_ERR_TOO_LARGE = -4,   /* an example does not fit the kernel's shared-memory staging  */
    FWGPU_ERR_IMMUTABLE = -5,   /* learn(update=1) on a forward-only ctx (regressor.rs:362-365) */
    FWGPU_ERR_NCCL = -6
};

/* model_instance.rs:24-29 Optimizer */
enum { FWGPU_OPT_SGD = 0, FWGPU_OPT_ADAGRAD_FLEX = 1, FWGPU_OPT_ADAGRAD_LUT = 2 };
/* weight blocks in the order regressor.rs:426-442 serialises them; FWGPU_BLOCK_NN0 + i = hidden layer i,
 * FWGPU_BLOCK_NN0 + nn_num_layers = the final neuron.  A layer holds (n_in + 1) * n_out weights: row-major per neuron,
 * then the biases (block_neural.rs:83-86), followed by as many accumulators (block_neural.rs:426-438). */
enum { FWGPU_BLOCK_LR = 0, FWGPU_BLOCK_FFM = 1, FWGPU_BLOCK_NN0 = 2 /* + layer index */ };

/* block_neural.rs:29-35 InitType.  Hu / Xavier draw from rand_xoshiro + rand_distr in the reference, which no test
 * observes and which cannot be reproduced without those crates ("parity unpinned"): fwgpu draws uniform
 * weights of the same variance from a fixed LCG.  Parity runs import identical weights (fwgpu_import_block). */
enum { FWGPU_NN_INIT_XAVIER = 0, FWGPU_NN_INIT_HU = 1, FWGPU_NN_INIT_ONE = 2, FWGPU_NN_INIT_ZERO = 3 };

#define FWGPU_MAX_NN_LAYERS 8
#define FWGPU_LUT_SIZE 2048 /* optimizer.rs:101-102 */

/*
 * The part of ModelInstance (model_instance.rs:47-97) the hot path reads, plus the flattened
 * feature_combo_descs / ffm_fields that FeatureBufferTranslator (feature_buffer.rs:138-172) needs
 * for the raw-record entry points.  All arrays are copied by fwgpu_create.
 */
typedef struct fwgpu_model_desc {
    /* optimizer hyper-parameters per block: block_lr.rs:63-65, block_ffm.rs:87-91, block_neural.rs:108-109 */
    float learning_rate, power_t, init_acc_gradient;
    float ffm_learning_rate, ffm_power_t, ffm_init_acc_gradient;
    float nn_learning_rate, nn_power_t, nn_init_acc_gradient;
    uint32_t bit_precision;     /* LR table = 1 << bit_precision cells (block_lr.rs:67)            */
    uint32_t ffm_bit_precision; /* FFM table = (1 << ffm_bit_precision) + F*k (block_ffm.rs:93-94) */
    uint32_t ffm_k;             /* 0 = no FFM block (regressor.rs:184)                             */
    uint32_t ffm_num_fields;    /* F = ffm_fields.len()                                            */
    uint32_t num_combos;        /* feature_combo_descs.len() + add_constant (block_lr.rs:52-55)    */
    uint32_t optimizer;         /* FWGPU_OPT_*                                                     */
    uint32_t immutable;         /* 1 = forward-only regressor (regressor.rs:471-534): SGD, no accumulators */
    float ffm_init_width, ffm_init_zero_band, ffm_init_center; /* block_ffm.rs:796-822 */
    /* dense head, topology "one" (regressor.rs:191-320); 0 layers = none */
    uint32_t nn_num_layers;                 /* hidden layers; the final single neuron (init One) is implied   */
    uint32_t nn_width[FWGPU_MAX_NN_LAYERS]; /* --nn i:width:W (default 20, regressor.rs:228-232)               */
    uint32_t nn_relu[FWGPU_MAX_NN_LAYERS];  /* --nn i:activation:relu|none                                      */
    uint32_t nn_init[FWGPU_MAX_NN_LAYERS];  /* FWGPU_NN_INIT_* (--nn i:init:..., default hu; block_neural.rs:367-412) */
    /* translate spec (only needed for fwgpu_*_records) */
    uint32_t n_namespaces;        /* vwmap.rs:31 num_namespaces                               */
    const uint8_t *ns_is_f32;     /* [n_namespaces] NamespaceFormat::F32 (vwmap.rs:16-20)      */
    uint32_t n_combos;            /* feature_combo_descs.len() (without the constant)          */
    const uint32_t *combo_off;    /* [n_combos+1] offsets into combo_ns                        */
    const uint32_t *combo_ns;     /* namespace_index of every descriptor of every combo        */
    const float *combo_weight;    /* [n_combos] FeatureComboDesc.weight                        */
    uint32_t add_constant;        /* add_constant_feature                                      */
    const uint32_t *field_off;    /* [ffm_num_fields+1] offsets into field_ns                  */
    const uint32_t *field_ns;     /* namespace_index of every namespace of every field         */
    uint32_t max_ffm_per_example; /* 0 = derive (one feature per field namespace); raise for multi-valued namespaces */
    uint32_t max_lr_per_example;  /* 0 = derive                                                */
    /* Hogwild concurrency ramp: a freshly initialised model is trained with at most
     * examples_seen / hogwild_ramp_div examples in flight, growing to the full machine; it keeps
     * the cold-start of AdaGrad (accumulators at 0) from overshooting when thousands of examples
     * hit the same weights at once.  0 = default (32); 0x7fffffff = sequential mode (one example in flight, general kernel); 0xffffffff = no ramp.  DESIGN.md "semantics". */
    uint32_t hogwild_ramp_div;
    /* Hard cap on examples in flight.  0 = automatic: unlimited for AdaGrad with power_t > 0 (the accumulators damp
     * concurrent steps on a hot weight), 16 for constant-step models (SGD, or power_t == 0) -- the width of the
     * reference's own Hogwild default (--hogwild_threads 16, main.rs:187-198): with a constant step, m examples in
     * flight multiply the effective learning rate of a hot weight by m. */
    uint32_t hogwild_max_inflight;
} fwgpu_model_desc;

/*
 * A mini-batch of translated examples = B reference FeatureBuffers (feature_buffer.rs:10-31) in
 * CSR form.  lr_*: HashAndValue{hash,value,combo_index}; ffm_*: HashAndValueAndSeq with the plain
 * field index (contra_field_index / ffm_k), sorted by field inside each example exactly as
 * translate emits them.  Hashes are already masked (feature_buffer.rs:140-148).
 */
typedef struct fwgpu_batch {
    uint32_t n_examples;
    const float *labels;       /* [B]  FeatureBuffer.label              */
    const float *importance;   /* [B]  FeatureBuffer.example_importance */
    const uint32_t *lr_off;    /* [B+1] */
    const uint32_t *lr_hash;
    const float *lr_val;
    const uint32_t *lr_combo;
    const uint32_t *ffm_off;   /* [B+1] (may be NULL when ffm_k == 0) */
    const uint32_t *ffm_hash;
    const float *ffm_val;
    const uint32_t *ffm_field;
} fwgpu_batch;

typedef struct fwgpu_ctx fwgpu_ctx;
typedef struct fwgpu_dataset fwgpu_dataset; /* records resident in HBM */

/* ---- lifecycle ---------------------------------------------------------------------------- */

/* Regressor::new = new_without_weights + allocate_and_init_weights (regressor.rs:153-165, 173-345):
 * allocates the tables in HBM and initialises them like the reference (LR zeros block_lr.rs:97-105,
 * FFM merand48 block_ffm.rs:784-829), builds the AdaGrad look-up tables (optimizer.rs:121-144). */
fwgpu_status fwgpu_create(const fwgpu_model_desc *desc, int device, fwgpu_ctx **out);
/* The same regressor as ONE model over the GPUs of one NVLink/NVSwitch box (BASELINE config 4): the LR / FFM tables are
 * hash-range-sharded, rank r owning indices [r*len/world, (r+1)*len/world) (+ the spill-over tail of block_ffm.rs:93-94 on
 * the last rank).  One process per GPU calls this collectively (rank in [0, world)); the reference's counterpart is the
 * single table all Hogwild workers share (hogwild.rs:24-103).  It takes the place of SURVEY 8b's
 * fwgpu_comm_init [ncclUniqueId, rank, nranks]: the ranks rendezvous over the unix sockets "<rendezvous>.<rank>" (owner-only),
 * exchange their shards' memory handles (CUDA VMM, POSIX fds) and the NCCL unique id there, and map all ranges into one
 * virtual range, so a row has the same address on every rank.
 *   - predict, and training with one record in flight (parity mode): rows are pulled from the owner's HBM by bulk copies over
 *     NVLink inside the learn kernel; parity-mode updates go back as bulk reductions applied by the owner's L2;
 *   - training (wide models: the bulk-copy kernel): every rank pulls the rows its records need, computes, and PUSHES each
 *     row's gradient (1.26 KB, one bulk store) into the owner's inbox; after every chunk of records (default 8192 per rank)
 *     ONE exchange step -- an NCCL all-gather of the per-owner entry counts, which is also the barrier -- and the owners
 *     apply AdaGrad from their own accumulators (owner-side update, no remote atomics).  These calls are COLLECTIVE: every
 *     rank must pass the same number of examples to fwgpu_learn_records / fwgpu_dataset_learn(update = 1);
 *   - narrow models (warp-per-record kernel) keep addressing remote rows directly (remote 128-bit loads / atomics).
 * Every other entry point works unchanged on such a ctx; fwgpu_import_block is collective (each rank writes the range it
 * owns); call fwgpu_shard_barrier before fwgpu_export_block so that every owner has applied what it was sent.
 * Fails with FWGPU_ERR_NCCL when libnccl.so.2 cannot be loaded.  Tables too small to split into whole allocation granules
 * live on rank 0.  timeout_ms 0 = 60 s. */
fwgpu_status fwgpu_create_sharded(const fwgpu_model_desc *desc, int device, uint32_t rank, uint32_t world,
                                  const char *rendezvous, uint32_t timeout_ms, fwgpu_ctx **out);
/* The layout arithmetic behind fwgpu_create_sharded, as a pure function (no GPU): sizes_out[world] = bytes of the table each
 * rank's HBM holds (equal hash ranges when they are whole allocation granules, else everything on rank 0; the tail goes to
 * the last rank), *owner_shift_out = s such that float index i of the table lives on rank min(i >> s, world - 1) (32 = rank 0). */
fwgpu_status fwgpu_debug_shard_plan(uint64_t bytes, uint64_t tail_bytes, uint32_t world, uint64_t granularity, uint64_t *sizes_out, uint32_t *owner_shift_out);
/* fwgpu_sync + wait until every rank of the shard group has done the same (no-op group of one for unsharded ctxs). */
fwgpu_status fwgpu_shard_barrier(fwgpu_ctx *ctx);
/* rank / world of the ctx's shard group and the FFM index range [first, first+count) this rank's HBM holds. */
fwgpu_status fwgpu_shard_info(const fwgpu_ctx *ctx, uint32_t *rank, uint32_t *world, uint64_t *ffm_first, uint64_t *ffm_count);
void fwgpu_destroy(fwgpu_ctx *ctx);
const char *fwgpu_last_error(const fwgpu_ctx *ctx); /* ctx may be NULL: last create() failure */
fwgpu_status fwgpu_sync(fwgpu_ctx *ctx);
void *fwgpu_stream(fwgpu_ctx *ctx);            /* the ctx's cudaStream_t, for event timing by the caller */
uint64_t fwgpu_launch_count(const fwgpu_ctx *ctx); /* kernels launched by this ctx so far */

/* ---- the hot path ------------------------------------------------------------------------- */

/* Regressor::learn (regressor.rs:356-379) over a mini-batch: forward, prediction out, and when
 * update != 0 (and importance != 0) backward + per-weight optimizer update, Hogwild-style with
 * device atomics.  preds_out: host pointer, [n_examples], the probability before the update. */
fwgpu_status fwgpu_learn_batch(fwgpu_ctx *ctx, const fwgpu_batch *batch, float *preds_out, int update);

/* Regressor::predict (regressor.rs:381-395) over a mini-batch. */
fwgpu_status fwgpu_predict_batch(fwgpu_ctx *ctx, const fwgpu_batch *batch, float *preds_out);

/* FeatureBufferTranslator::translate + Regressor::learn, i.e. the body of the reference's train
 * loop (main.rs:240-256) and of a Hogwild worker (hogwild.rs:93-100), on raw parser/cache records
 * (parser.rs:57-74).  records: n_words u32 words holding n_examples records back to back;
 * rec_off: [n_examples+1] word offsets of each record, or NULL when every record has the fixed
 * length n_words / n_examples.  Translation runs on the device and is bit-exact with the reference. */
fwgpu_status fwgpu_learn_records(fwgpu_ctx *ctx, const uint32_t *records, uint64_t n_words,
                                 const uint32_t *rec_off, uint32_t n_examples, float *preds_out, int update);

/* The translated form of a record batch, for bit-exact comparison with the reference's translate
 * (feature_buffer.rs:178-338).  Fills caller arrays (host); capacities in entries; counts returned
 * through lr_off/ffm_off [n_examples+1].  Synchronous. */
fwgpu_status fwgpu_translate_records(fwgpu_ctx *ctx, const uint32_t *records, uint64_t n_words,
                                     const uint32_t *rec_off, uint32_t n_examples,
                                     float *labels, float *importance,
                                     uint32_t *lr_off, uint32_t *lr_hash, float *lr_val, uint32_t *lr_combo, uint64_t lr_cap,
                                     uint32_t *ffm_off, uint32_t *ffm_hash, float *ffm_val, uint32_t *ffm_field, uint64_t ffm_cap);

/* Records kept resident in HBM (the input cache of cache.rs held on the device): upload once,
 * then learn over [first, first+count) slices without host traffic.  preds_out may be NULL. */
fwgpu_status fwgpu_dataset_upload(fwgpu_ctx *ctx, const uint32_t *records, uint64_t n_words,
                                  const uint32_t *rec_off, uint64_t n_examples, fwgpu_dataset **out);
fwgpu_status fwgpu_dataset_learn(fwgpu_ctx *ctx, fwgpu_dataset *ds, uint64_t first, uint64_t count,
                                 float *preds_out, int update);
void fwgpu_dataset_free(fwgpu_ctx *ctx, fwgpu_dataset *ds);

/* ---- weights (regressor.rs:426-469 write_weights_to_buf / overwrite_weights_from_buf) ------ */

/* BlockTrait::get_serialized_len (block_lr.rs:257-259, block_ffm.rs:831-833, block_neural.rs:414-416):
 * number of weights of a block; bytes = what the reference writes for that block
 * (LR: len * 8 {w,acc} or len * 4 for SGD; FFM/NN: len*4 weights then len*4 accumulators, SGD: weights only). */
fwgpu_status fwgpu_block_len(const fwgpu_ctx *ctx, int block, uint64_t *n_weights, uint64_t *n_bytes);
/* write_weights_to_buf for one block in the reference byte layout (block_helpers.rs:99-124). Synchronous. */
fwgpu_status fwgpu_export_block(fwgpu_ctx *ctx, int block, void *dst, uint64_t dst_bytes);
/* read_weights_from_buf (block_helpers.rs:43-60).  with_optimizer_state = 0 reads a weights-only
 * payload (read_weights_from_buf_into_forward_only, block_lr.rs:277-292, block_ffm.rs:879-899). */
fwgpu_status fwgpu_import_block(fwgpu_ctx *ctx, int block, const void *src, uint64_t src_bytes, int with_optimizer_state);
/* Number of examples this ctx has learned from (drives the concurrency ramp).  Importing weights
 * with optimizer state marks the model as trained (no ramp); set it explicitly when needed. */
fwgpu_status fwgpu_set_examples_seen(fwgpu_ctx *ctx, uint64_t n);
uint64_t fwgpu_get_examples_seen(const fwgpu_ctx *ctx);
/* the AdaGrad LUT of a block (which: 0 lr, 1 ffm, 2 nn), 2048 floats, for inspection */
fwgpu_status fwgpu_get_lut(const fwgpu_ctx *ctx, int which, float *dst2048);

/* ---- measurement helpers ------------------------------------------------------------------ */

/* When enabled, every hot-path kernel launch is bracketed by CUDA events on the ctx stream;
 * fwgpu_kernel_time returns the accumulated device time and launch count per kernel kind
 * (0 = learn/predict kernel, 1 = translate kernels). */
fwgpu_status fwgpu_set_profiling(fwgpu_ctx *ctx, int enabled);
fwgpu_status fwgpu_kernel_time(fwgpu_ctx *ctx, int kind, double *total_ms, uint64_t *launches);

/* Which kernel family did the work so far (tests assert that the kernels the bench times are the ones under test):
 * out4 = { launches of the warp-per-record fused kernel (k_learn_fixed), launches of the block-per-record fused kernel
 * (k_learn_fixed_cta), launches of the general kernel (k_learn), examples the general kernel processed }.  Synchronous. */
fwgpu_status fwgpu_debug_path_counts(fwgpu_ctx *ctx, uint64_t *out4);

/* Debug/verification: out[i] = logistic(in[i]) computed by the device routine the kernels use
 * (block_loss_functions.rs:15-17 with glibc-exact expf).  Host pointers, synchronous. */
fwgpu_status fwgpu_debug_logistic(fwgpu_ctx *ctx, const float *in, float *out, uint64_t n);

/* pinned host memory for batch buffers (SURVEY.md section 8b "Batch layout") */
fwgpu_status fwgpu_host_alloc(void **out, uint64_t bytes);
void fwgpu_host_free(void *p);

/* library identity: "fwgpu <version> sm_100a" */
const char *fwgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
