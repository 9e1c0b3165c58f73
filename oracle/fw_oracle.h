/*
 * fw_oracle.h -- CPU oracle for the Fwumious Wabbit LR/FFM learn/predict hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's algorithm
 * (outbrain-inc/fwumious_wabbit, Rust) for the path SURVEY.md section 8 names.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (fwumious_wabbit_b200/, libfwgpu.so) never links, imports or calls it.
 *
 * Parity status: pinned against the golden values of the reference's own unit tests
 * (tests/test_oracle_goldens.py lists each one with its reference file:line).  Two pieces
 * are "parity unpinned" because no reference test observes them: the merand48 FFM init
 * (block_ffm.rs:796-806, third-party crate merand48 0.1.0) and the neural-layer random
 * init (block_neural.rs:385-406, rand_xoshiro/rand_distr); both are restated from their
 * published definitions and parity runs load identical weights on both sides.
 *
 * Every function cites the reference file:line it follows (paths relative to src/).
 */
#ifndef FW_ORACLE_H
#define FW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (parser.rs:12-21, feature_buffer.rs:6-8, optimizer.rs:101-102) ---- */
#define FWO_HEADER_LEN 3u
#define FWO_LABEL_OFFSET 1u
#define FWO_IMPORTANCE_OFFSET 2u
#define FWO_IS_NOT_SINGLE_MASK 0x80000000u
#define FWO_MASK31 0x7fffffffu
#define FWO_NO_FEATURES 0x80000000u
#define FWO_NO_LABEL 0xffu
#define FWO_FLOAT32_ONE 1065353216u
#define FWO_VOWPAL_FNV_PRIME 16777619u
#define FWO_CONSTANT_HASH 11650396u
#define FWO_LUT_BITS 11
#define FWO_LUT_SIZE 2048

enum { FWO_OPT_SGD = 0, FWO_OPT_ADAGRAD_FLEX = 1, FWO_OPT_ADAGRAD_LUT = 2 };
enum { FWO_GRAPH_REGRESSOR = 0, FWO_GRAPH_FFM_BLOCK_ONLY = 1 };
enum { FWO_NN_INIT_XAVIER = 0, FWO_NN_INIT_HU = 1, FWO_NN_INIT_ONE = 2, FWO_NN_INIT_ZERO = 3 };

/* feature_buffer.rs:10-31 */
typedef struct { uint32_t hash; float value; uint32_t combo_index; } fwo_lr_feat;
typedef struct { uint32_t hash; float value; uint32_t contra_field_index; /* = field*k */ } fwo_ffm_feat;
typedef struct {
    float label;
    float example_importance;
    uint64_t example_number;
    uint32_t n_lr;
    const fwo_lr_feat *lr;
    uint32_t n_ffm;
    const fwo_ffm_feat *ffm;
} fwo_feature_buffer;

#define FWO_MAX_NN_LAYERS 8
/* The hyper-parameters of model_instance.rs:47-97 that the hot path reads. */
typedef struct {
    float learning_rate, power_t, init_acc_gradient;             /* LR block   */
    float ffm_learning_rate, ffm_power_t, ffm_init_acc_gradient; /* FFM block  */
    float nn_learning_rate, nn_power_t, nn_init_acc_gradient;    /* NN blocks  */
    uint32_t bit_precision;
    uint32_t ffm_bit_precision;
    uint32_t ffm_k;          /* 0 = no FFM block */
    uint32_t ffm_num_fields; /* F */
    uint32_t num_combos;     /* feature_combo_descs.len() + (add_constant_feature ? 1 : 0) */
    uint32_t optimizer;      /* FWO_OPT_* */
    uint32_t graph;          /* FWO_GRAPH_* */
    float ffm_init_width, ffm_init_zero_band, ffm_init_center;   /* block_ffm.rs:796-822 */
    uint32_t nn_num_layers;  /* 0 = no head (regressor.rs:191) ; topology "one" only */
    uint32_t nn_width[FWO_MAX_NN_LAYERS];
    uint32_t nn_relu[FWO_MAX_NN_LAYERS];
    uint32_t nn_init[FWO_MAX_NN_LAYERS];
    float nn_maxnorm[FWO_MAX_NN_LAYERS];
} fwo_model_desc;

/* Flat description of what FeatureBufferTranslator reads from ModelInstance
 * (feature_combo_descs, ffm_fields, add_constant_feature; feature_buffer.rs:138-172). */
typedef struct {
    uint32_t n_namespaces;
    const uint8_t *ns_is_f32;     /* [n_namespaces] NamespaceFormat::F32 ? 1 : 0 (vwmap.rs:16-20) */
    uint32_t n_combos;            /* without the constant */
    const uint32_t *combo_off;    /* [n_combos+1] into combo_ns */
    const uint32_t *combo_ns;     /* namespace indices */
    const float *combo_weight;    /* [n_combos] */
    uint32_t add_constant;
    uint32_t n_fields;
    const uint32_t *field_off;    /* [n_fields+1] into field_ns */
    const uint32_t *field_ns;
    uint32_t bit_precision, ffm_bit_precision, ffm_k;
} fwo_translate_spec;

typedef struct fwo_regressor fwo_regressor;

/* ---- hashing (third-party fasthash 0.4 murmur3::hash32[_with_seed] = MurmurHash3_x86_32;
 *      call sites parser.rs:82-83, 382-385) ---- */
uint32_t fwo_murmur3_32(const void *key, size_t len, uint32_t seed);

/* ---- optimizer (optimizer.rs) ---- */
void fwo_lut_build(float learning_rate, float power_t, float initial_acc_gradient, float *lut2048);
float fwo_opt_update(uint32_t optimizer, float lr, float minus_power_t, const float *lut,
                     float gradient, float *acc);

/* ---- merand48 (third-party crate merand48 0.1.0 = VW's merand48; block_ffm.rs:801,811) ---- */
float fwo_merand48(uint64_t seed);

/* ---- text parser: one VW line -> u32 record (parser.rs:214-461) ----
 * names: n_namespaces C strings (vw names), index = namespace_index; ns_is_f32 as above.
 * Returns record length in words (>0), 0 for an empty line/EOF, <0 on error (err filled).
 * FWO_PARSE_FLUSH (-2) for the "flush" command (parser.rs:227-238). */
#define FWO_PARSE_ERROR (-1)
#define FWO_PARSE_FLUSH (-2)
#define FWO_PARSE_HOGWILD_LOAD (-3)
int fwo_parse_line(const char *const *ns_names, const uint8_t *ns_is_f32, uint32_t n_namespaces,
                   uint32_t namespace_skip_prefix, const char *line, size_t line_len,
                   uint32_t *out, size_t out_cap, char *err, size_t err_cap);

/* ---- translate: record -> feature buffer (feature_buffer.rs:178-338) ----
 * Returns 0 on success; counts written to *n_lr / *n_ffm; -1 if capacity exceeded. */
int fwo_translate(const fwo_translate_spec *spec, const uint32_t *record,
                  fwo_lr_feat *lr, uint32_t lr_cap, uint32_t *n_lr,
                  fwo_ffm_feat *ffm, uint32_t ffm_cap, uint32_t *n_ffm,
                  float *label, float *importance);
uint32_t fwo_lr_hash_mask(uint32_t bit_precision);
uint32_t fwo_ffm_hash_mask(uint32_t ffm_bit_precision, uint32_t ffm_k);

/* ---- regressor (regressor.rs:173-395) ---- */
fwo_regressor *fwo_regressor_new(const fwo_model_desc *desc); /* allocate_and_init_weights included */
void fwo_regressor_free(fwo_regressor *r);
/* Regressor::learn (regressor.rs:356-379): update && importance != 0 -> forward_backward,
 * otherwise the predict-order forward. */
float fwo_learn(fwo_regressor *r, const fwo_feature_buffer *fb, int update);
/* Regressor::predict (regressor.rs:381-395). */
float fwo_predict(fwo_regressor *r, const fwo_feature_buffer *fb);
/* block_helpers.rs:161-173 slearn2: the training-order chain even when update == 0. */
float fwo_forward_backward(fwo_regressor *r, const fwo_feature_buffer *fb, int update);

/* weight access (layouts: block_helpers.rs:17-28; block_lr.rs:97-105; block_ffm.rs:784-792) */
uint32_t fwo_lr_len(const fwo_regressor *r);  /* 1 << bit_precision */
float *fwo_lr_table(fwo_regressor *r);        /* AoS {w, acc} x lr_len (acc unused for SGD) */
uint32_t fwo_ffm_len(const fwo_regressor *r); /* (1 << ffm_bit_precision) + F*k */
float *fwo_ffm_weights(fwo_regressor *r);
float *fwo_ffm_acc(fwo_regressor *r);
uint32_t fwo_nn_layer_count(const fwo_regressor *r); /* hidden layers + final neuron */
uint32_t fwo_nn_layer_len(const fwo_regressor *r, uint32_t layer); /* (n_in+1)*n_out */
float *fwo_nn_weights(fwo_regressor *r, uint32_t layer);
float *fwo_nn_acc(fwo_regressor *r, uint32_t layer);
const float *fwo_lut(const fwo_regressor *r, int which /*0 lr, 1 ffm, 2 nn*/);
/* one neuron layer (+ relu) in isolation, as block_neural.rs:507-581 / block_relu.rs:156-173 test it */
int fwo_test_neuron_layer(uint32_t optimizer, float lr, float power_t, float init_acc, uint32_t n_in, uint32_t n_out, uint32_t init,
                          uint32_t relu, const float *x, const float *d_out, uint32_t n_steps, float *outs, float *d_in);

/* ---- batch drivers used by tests and the CPU baseline ----
 * Fixed-capacity CSR batch identical to include/fwgpu.h's fwgpu_batch (plain arrays). */
typedef struct {
    uint32_t n_examples;
    const float *labels;
    const float *importance;
    const uint32_t *lr_off;   /* [n+1] */
    const uint32_t *lr_hash;
    const float *lr_val;
    const uint32_t *lr_combo;
    const uint32_t *ffm_off;  /* [n+1] */
    const uint32_t *ffm_hash;
    const float *ffm_val;
    const uint32_t *ffm_field; /* plain field index (contra_field_index / k) */
} fwo_batch;

/* Sequential (reference default mode, main.rs:213-258): learn example by example, write the
 * prediction made before each update. */
void fwo_learn_batch_sequential(fwo_regressor *r, const fwo_batch *b, float *preds, int update);

/* Hogwild (hogwild.rs:24-103): n_threads workers, each translate + learn(update=true) on the
 * shared tables without locks.  records: back-to-back u32 records; rec_off[n_records+1] word
 * offsets.  Returns elapsed seconds of the parallel section.  preds may be NULL (the reference
 * produces none in this mode, main.rs:242-247). */
double fwo_hogwild_run(fwo_regressor *r, const fwo_translate_spec *spec, const uint32_t *records,
                       const uint64_t *rec_off, uint64_t n_records, uint32_t n_threads, float *preds);

/* Analysis tool (no reference counterpart): emulates a device that keeps `wave` examples in flight
 * (all scored on one weight snapshot, then updated; mode 0 = per-example updates in order,
 * mode 1 = per-slot aggregated update).  See fw_oracle.c. */
void fwo_learn_records_wave(fwo_regressor *r, const fwo_translate_spec *spec, const uint32_t *records,
                            const uint64_t *rec_off, uint64_t n_records, uint32_t wave, int mode, float *preds);
/* Test tool: batched semantics of a model with a dense head (sub-batch of `wave` examples against one snapshot; dense layers
 * take summed gradients G1 / G2, sparse tables are updated example by example).  wave == 1 is fwo_learn. */
void fwo_learn_records_head_wave(fwo_regressor *r, const fwo_translate_spec *spec, const uint32_t *records,
                                 const uint64_t *rec_off, uint64_t n_records, uint32_t wave, float *preds);

#ifdef __cplusplus
}
#endif
#endif
