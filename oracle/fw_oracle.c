/*
 * fw_oracle.c -- CPU oracle: plain-C restatement of the Fwumious Wabbit hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see fw_oracle.h).  Never linked into the product.
 * Compile with -ffp-contract=off: the reference is Rust, which never contracts a*b+c into an
 * FMA on its own, and the default release build has no +fma target feature (build.sh:10), so
 * every multiply and add below must round separately to reproduce the reference's goldens.
 *
 * All "reference" citations are file:line under /root/reference/src/.
 */
#include "fw_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * MurmurHash3_x86_32 (Austin Appleby, public domain) -- what fasthash 0.4 murmur3::hash32 and
 * hash32_with_seed compute.  Call sites: parser.rs:82-83 (namespace seed, seed 0) and
 * parser.rs:382-385 (feature hash seeded with the namespace hash).  Pinned by the parser
 * known answers (parser.rs:497,553,570,724,1040) in tests/test_oracle_goldens.py.
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int8_t r) { return (x << r) | (x >> (32 - r)); }

uint32_t fwo_murmur3_32(const void *key, size_t len, uint32_t seed)
{
    const uint8_t *data = (const uint8_t *)key;
    const size_t nblocks = len / 4;
    uint32_t h1 = seed;
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    for (size_t i = 0; i < nblocks; i++) {
        uint32_t k1;
        memcpy(&k1, data + 4 * i, 4);
        k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
        h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
    }
    const uint8_t *tail = data + nblocks * 4;
    uint32_t k1 = 0;
    switch (len & 3) {
    case 3: k1 ^= (uint32_t)tail[2] << 16; /* fallthrough */
    case 2: k1 ^= (uint32_t)tail[1] << 8;  /* fallthrough */
    case 1: k1 ^= tail[0];
        k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint32_t)len;
    h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
    return h1;
}

/* ------------------------------------------------------------------------------------------
 * merand48 (crate merand48 0.1.0 == Vowpal Wabbit's merand48, named in SPEED.md:38-39).
 * One LCG step from the seed, top mantissa bits -> [0,1).  PARITY UNPINNED: no reference test
 * observes an init value (all FFM tests overwrite weights with 1.0).
 * ---------------------------------------------------------------------------------------- */
float fwo_merand48(uint64_t seed)
{
    const uint64_t a = 0xeece66d5deece66dULL, c = 2147483647ULL;
    uint64_t s = a * seed + c;
    uint32_t bits = (uint32_t)((s >> 25) & 0x7FFFFF) | 0x3F800000u;
    float f;
    memcpy(&f, &bits, 4);
    return f - 1.0f;
}

/* ------------------------------------------------------------------------------------------
 * Optimizers (optimizer.rs)
 * ---------------------------------------------------------------------------------------- */
static inline float bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

/* OptimizerAdagradLUT::init, optimizer.rs:121-144 */
void fwo_lut_build(float learning_rate, float power_t, float initial_acc_gradient, float *lut)
{
    float minus_power_t = -power_t;
    for (uint32_t x = 0; x < FWO_LUT_SIZE; x++) {
        float float_x = bits2f(x << (31 - FWO_LUT_BITS)) + initial_acc_gradient;
        float float_x_plus_one = bits2f((x + 1) << (31 - FWO_LUT_BITS)) + initial_acc_gradient;
        float val = learning_rate * (powf(float_x, minus_power_t) + powf(float_x_plus_one, minus_power_t)) * 0.5f;
        if (isnan(val) || isinf(val)) val = learning_rate;
        lut[x] = val;
    }
}

/* calculate_update for the three optimizers: SGD optimizer.rs:35-37, AdagradFlex :76-89,
 * AdagradLUT :147-156. */
float fwo_opt_update(uint32_t optimizer, float lr, float minus_power_t, const float *lut,
                     float gradient, float *acc)
{
    if (optimizer == FWO_OPT_SGD) return gradient * lr;
    float gradient_squared = gradient * gradient;
    float new_acc = *acc + gradient_squared;
    *acc = new_acc;
    if (optimizer == FWO_OPT_ADAGRAD_LUT) {
        uint32_t key = f2bits(new_acc) >> (31 - FWO_LUT_BITS);
        return gradient * lut[key];
    }
    float update = gradient * lr * powf(new_acc, minus_power_t);
    if (isnan(update) || isinf(update)) return 0.0f;
    return update;
}

typedef struct {
    uint32_t kind;
    float lr, minus_power_t, init_acc;
    float lut[FWO_LUT_SIZE];
} fwo_opt;

static void opt_init(fwo_opt *o, uint32_t kind, float lr, float power_t, float init_acc)
{
    o->kind = kind; o->lr = lr; o->minus_power_t = -power_t; o->init_acc = init_acc;
    memset(o->lut, 0, sizeof(o->lut));
    if (kind == FWO_OPT_ADAGRAD_LUT) fwo_lut_build(lr, power_t, init_acc, o->lut);
}
/* initial_data(): Flex optimizer.rs:91-93 = initial_acc_gradient; LUT :158-161 = 0 (folded in LUT) */
static float opt_initial_data(const fwo_opt *o)
{
    return o->kind == FWO_OPT_ADAGRAD_FLEX ? o->init_acc : 0.0f;
}
static inline float opt_update(const fwo_opt *o, float g, float *acc)
{
    return fwo_opt_update(o->kind, o->lr, o->minus_power_t, o->lut, g, acc);
}

/* ------------------------------------------------------------------------------------------
 * Text parser (parser.rs:214-461).  Namespace lookup is a linear scan (the reference uses a
 * radix tree, radix_tree.rs; any exact-match map is equivalent).
 * ---------------------------------------------------------------------------------------- */
static int parse_f32(const char *s, size_t a, size_t b, float *out)
{
    /* parser.rs:110-139 parse_float_or_error: "NONE" -> NaN, else Rust str::parse::<f32>() */
    if (b - a == 4 && memcmp(s + a, "NONE", 4) == 0) { *out = NAN; return 0; }
    if (b <= a || b - a > 63) return -1;
    char tmp[64];
    memcpy(tmp, s + a, b - a);
    tmp[b - a] = 0;
    /* Rust's f32 parser rejects leading/trailing whitespace and hex floats; strtof accepts more.
     * Reject the extras that matter for VW files. */
    for (size_t i = 0; i < b - a; i++) {
        char c = tmp[i];
        if (!((c >= '0' && c <= '9') || c == '.' || c == '-' || c == '+' || c == 'e' || c == 'E' ||
              c == 'i' || c == 'n' || c == 'f' || c == 'a' || c == 'N' || c == 'I' || c == 'F' || c == 'A' ||
              c == 't' || c == 'y' || c == 'T' || c == 'Y'))
            return -1;
    }
    char *end = NULL;
    float v = strtof(tmp, &end);
    if (end == tmp || *end != 0) return -1;
    *out = v;
    return 0;
}

int fwo_parse_line(const char *const *ns_names, const uint8_t *ns_is_f32, uint32_t n_namespaces,
                   uint32_t namespace_skip_prefix, const char *p, size_t line_len,
                   uint32_t *out, size_t out_cap, char *err, size_t err_cap)
{
#define PERR(...) do { if (err && err_cap) snprintf(err, err_cap, __VA_ARGS__); return FWO_PARSE_ERROR; } while (0)
    if (line_len == 0) return 0;
    size_t bufpos = n_namespaces + FWO_HEADER_LEN;
    if (out_cap < bufpos) PERR("output buffer too small");
    size_t olen = bufpos;
    for (size_t i = 0; i < bufpos; i++) out[i] = FWO_NO_FEATURES; /* parser.rs:220-221 */
    size_t tmp_read_buf_size = line_len; /* includes the trailing '\n' like read_until(0x0a) */
    size_t i_start, i_end = 0;

    switch ((unsigned char)p[0]) { /* parser.rs:227-268 */
    case 0x31: out[FWO_LABEL_OFFSET] = 1; break;
    case 0x2d: out[FWO_LABEL_OFFSET] = 0; break;
    case 0x7c: out[FWO_LABEL_OFFSET] = FWO_NO_LABEL; break;
    default:
        if (tmp_read_buf_size >= 5 && memcmp(p, "flush", 5) == 0) return FWO_PARSE_FLUSH;
        if (tmp_read_buf_size >= strlen("hogwild_load ")) {
            /* parse_cmd (parser.rs:141-156): split on runs of spaces; exactly two tokens and the
             * first one "hogwild_load" -> command, any other token count -> error. */
            size_t ntok = 0, i = 0, first_len = 0;
            while (i < tmp_read_buf_size) {
                size_t s = i;
                while (i < tmp_read_buf_size && p[i] != 0x20) i++;
                if (ntok == 0) first_len = i - s;
                ntok++;
                while (i < tmp_read_buf_size && p[i] == 0x20) i++;
            }
            if (ntok == 2 && first_len == 12 && memcmp(p, "hogwild_load", 12) == 0)
                return FWO_PARSE_HOGWILD_LOAD;
        }
        PERR("Cannot parse an example");
    }

    size_t rowlen = tmp_read_buf_size - 1; /* parser.rs:270 ignore last newline byte */
    if (out[FWO_LABEL_OFFSET] == FWO_NO_LABEL) {
        out[FWO_IMPORTANCE_OFFSET] = FWO_FLOAT32_ONE;
    } else {
        while (p[i_end] != 0x20 && i_end < rowlen) i_end++;
        while (p[i_end] == 0x20 && i_end < rowlen) i_end++;
        if (p[i_end] == 0x7c) {
            out[FWO_IMPORTANCE_OFFSET] = FWO_FLOAT32_ONE;
        } else {
            i_start = i_end;
            while (p[i_end] != 0x20 && i_end < rowlen) i_end++;
            float importance;
            if (parse_f32(p, i_start, i_end, &importance) != 0)
                PERR("Failed parsing example importance: %.*s", (int)(i_end - i_start), p + i_start);
            if (importance < 0.0f) PERR("Example importance cannot be negative: %g! ", importance);
            out[FWO_IMPORTANCE_OFFSET] = f2bits(importance);
        }
    }
    while (p[i_end] != 0x7c && i_end < rowlen) i_end++; /* parser.rs:311-313 */

    uint32_t cur_seed = 0;
    size_t cur_index_offset = FWO_HEADER_LEN;
    int cur_is_f32 = 0;
    size_t bufpos_namespace_start = 0;
    float cur_ns_weight = 1.0f;
    uint32_t cur_num_features = 0;

    while (i_end < rowlen) { /* parser.rs:321-453 */
        while (p[i_end] == 0x20 && i_end < rowlen) i_end++;
        i_start = i_end;
        while (p[i_end] != 0x20 && p[i_end] != 0x3a && i_end < rowlen) i_end++;
        size_t i_end_first_part = i_end;
        while (p[i_end] != 0x20 && i_end < rowlen) i_end++;

        if (p[i_start] == 0x7c) {
            i_start += 1;
            if (i_end_first_part != i_end) {
                if (parse_f32(p, i_end_first_part + 1, i_end, &cur_ns_weight) != 0)
                    PERR("Failed parsing namespace weight: %.*s", (int)(i_end - i_end_first_part - 1), p + i_end_first_part + 1);
            } else {
                cur_ns_weight = 1.0f;
            }
            size_t nlen = i_end_first_part - i_start;
            uint32_t found = UINT32_MAX;
            for (uint32_t n = 0; n < n_namespaces; n++) {
                if (ns_names[n] && strlen(ns_names[n]) == nlen && memcmp(ns_names[n], p + i_start, nlen) == 0) { found = n; break; }
            }
            if (found == UINT32_MAX)
                PERR("Feature name was not predeclared in vw_namespace_map.csv: %.*s", (int)nlen, p + i_start);
            cur_seed = fwo_murmur3_32(ns_names[found], nlen, 0); /* parser.rs:82-83 */
            cur_index_offset = found + FWO_HEADER_LEN;
            cur_is_f32 = ns_is_f32 ? ns_is_f32[found] : 0;
            cur_num_features = 0;
            bufpos_namespace_start = olen;
        } else {
            uint32_t h = fwo_murmur3_32(p + i_start, i_end_first_part - i_start, cur_seed) & FWO_MASK31;
            float feature_weight = 1.0f;
            if (i_end_first_part != i_end) {
                if (parse_f32(p, i_end_first_part + 1, i_end, &feature_weight) != 0)
                    PERR("Failed parsing feature weight: %.*s", (int)(i_end - i_end_first_part - 1), p + i_end_first_part + 1);
            }
            if (cur_num_features == 0 && !cur_is_f32 && cur_ns_weight == 1.0f && feature_weight == 1.0f) {
                out[cur_index_offset] = h;
            } else {
                uint32_t feature_output = out[cur_index_offset];
                if (olen + 4 > out_cap) PERR("record too long");
                if (cur_num_features == 1 && (feature_output & FWO_IS_NOT_SINGLE_MASK) == 0) {
                    out[olen++] = feature_output;
                    out[olen++] = FWO_FLOAT32_ONE;
                }
                out[olen++] = h;
                if (cur_is_f32) {
                    size_t float_start = i_start + namespace_skip_prefix;
                    float float_value = NAN;
                    if (i_end_first_part != float_start) {
                        if (float_start > i_end_first_part ||
                            parse_f32(p, float_start, i_end_first_part, &float_value) != 0)
                            PERR("Failed parsing feature value to float (for float namespace): %.*s",
                                 (int)(i_end_first_part - i_start), p + i_start);
                    }
                    out[olen++] = f2bits(float_value);
                    if (cur_ns_weight * feature_weight != 1.0f)
                        PERR("Namespaces that are f32 can not have weight attached neither to namespace nor to a single feature (basically they can' use :weight syntax");
                } else {
                    out[olen++] = f2bits(cur_ns_weight * feature_weight);
                }
                out[cur_index_offset] = FWO_IS_NOT_SINGLE_MASK | (uint32_t)((bufpos_namespace_start << 16) + olen);
            }
            cur_num_features += 1;
        }
        i_end += 1;
    }
    out[0] = (uint32_t)olen;
    return (int)olen;
#undef PERR
}

/* ------------------------------------------------------------------------------------------
 * Translate (feature_buffer.rs:138-338)
 * ---------------------------------------------------------------------------------------- */
uint32_t fwo_lr_hash_mask(uint32_t bit_precision) { return (1u << bit_precision) - 1u; } /* :140 */

uint32_t fwo_ffm_hash_mask(uint32_t ffm_bit_precision, uint32_t ffm_k) /* :142-148 */
{
    uint32_t bits = 0;
    while (ffm_k > (1u << bits)) bits++;
    uint32_t dimensions_mask = (1u << bits) - 1u;
    return ((1u << ffm_bit_precision) - 1u) ^ dimensions_mask;
}

/* feature_reader! (feature_buffer.rs:48-108), primitive namespaces only: iterate the features
 * of one namespace slot.  Returns count; (hash,value) pairs written to hv (cap pairs). */
typedef struct { uint32_t hash; float value; } hv_t;
static uint32_t read_namespace(const uint32_t *rec, uint32_t ns_index, int is_f32, hv_t *hv, uint32_t cap)
{
    uint32_t first_token = rec[ns_index + FWO_HEADER_LEN];
    if ((first_token & FWO_IS_NOT_SINGLE_MASK) == 0) {
        if (cap < 1) return UINT32_MAX;
        hv[0].hash = first_token; hv[0].value = 1.0f;
        return 1;
    }
    uint32_t start = (first_token >> 16) & 0x3fff;
    uint32_t end = first_token & 0xffff;
    uint32_t n = 0;
    for (uint32_t off = start; off < end; off += 2) {
        if (n >= cap) return UINT32_MAX;
        hv[n].hash = rec[off];
        hv[n].value = is_f32 ? 1.0f : bits2f(rec[off + 1]);
        n++;
    }
    return n;
}

#define FWO_MAX_NS_FEATS 4096
int fwo_translate(const fwo_translate_spec *spec, const uint32_t *rec,
                  fwo_lr_feat *lr, uint32_t lr_cap, uint32_t *n_lr_out,
                  fwo_ffm_feat *ffm, uint32_t ffm_cap, uint32_t *n_ffm_out,
                  float *label, float *importance)
{
    const uint32_t lr_mask = fwo_lr_hash_mask(spec->bit_precision);
    uint32_t n_lr = 0, n_ffm = 0;
    *label = (float)rec[FWO_LABEL_OFFSET];            /* :190 */
    *importance = bits2f(rec[FWO_IMPORTANCE_OFFSET]); /* :191-192 */

    static __thread hv_t *buf_a = NULL, *buf_b = NULL, *buf_ns = NULL;
    if (!buf_a) {
        buf_a = (hv_t *)malloc(sizeof(hv_t) * FWO_MAX_NS_FEATS);
        buf_b = (hv_t *)malloc(sizeof(hv_t) * FWO_MAX_NS_FEATS);
        buf_ns = (hv_t *)malloc(sizeof(hv_t) * FWO_MAX_NS_FEATS);
    }

    for (uint32_t c = 0; c < spec->n_combos; c++) { /* :197-268 */
        const uint32_t *nss = spec->combo_ns + spec->combo_off[c];
        uint32_t num_ns = spec->combo_off[c + 1] - spec->combo_off[c];
        float w = spec->combo_weight[c];
        if (num_ns == 1) {
            uint32_t n = read_namespace(rec, nss[0], spec->ns_is_f32 ? spec->ns_is_f32[nss[0]] : 0, buf_a, FWO_MAX_NS_FEATS);
            if (n == UINT32_MAX || n_lr + n > lr_cap) return -1;
            for (uint32_t i = 0; i < n; i++) {
                lr[n_lr].hash = buf_a[i].hash & lr_mask;
                lr[n_lr].value = buf_a[i].value * w;
                lr[n_lr].combo_index = c;
                n_lr++;
            }
        } else {
            hv_t *in = buf_a, *out = buf_b;
            uint32_t n_in = read_namespace(rec, nss[0], spec->ns_is_f32 ? spec->ns_is_f32[nss[0]] : 0, in, FWO_MAX_NS_FEATS);
            if (n_in == UINT32_MAX) return -1;
            for (uint32_t j = 1; j < num_ns; j++) {
                uint32_t n_ns = read_namespace(rec, nss[j], spec->ns_is_f32 ? spec->ns_is_f32[nss[j]] : 0, buf_ns, FWO_MAX_NS_FEATS);
                if (n_ns == UINT32_MAX) return -1;
                uint32_t n_out = 0;
                for (uint32_t a = 0; a < n_in; a++) {
                    uint32_t half_hash = in[a].hash * FWO_VOWPAL_FNV_PRIME; /* wrapping, :239 */
                    for (uint32_t b = 0; b < n_ns; b++) {
                        if (n_out >= FWO_MAX_NS_FEATS) return -1;
                        out[n_out].hash = buf_ns[b].hash ^ half_hash;
                        out[n_out].value = in[a].value * buf_ns[b].value;
                        n_out++;
                    }
                }
                hv_t *t = in; in = out; out = t;
                n_in = n_out;
            }
            if (n_lr + n_in > lr_cap) return -1;
            for (uint32_t i = 0; i < n_in; i++) {
                lr[n_lr].hash = in[i].hash & lr_mask;
                lr[n_lr].value = in[i].value * w;
                lr[n_lr].combo_index = c;
                n_lr++;
            }
        }
    }
    if (spec->add_constant) { /* :270-276 */
        if (n_lr + 1 > lr_cap) return -1;
        lr[n_lr].hash = FWO_CONSTANT_HASH & lr_mask;
        lr[n_lr].value = 1.0f;
        lr[n_lr].combo_index = spec->n_combos;
        n_lr++;
    }
    if (spec->ffm_k > 0) { /* :279-335 */
        const uint32_t ffm_mask = fwo_ffm_hash_mask(spec->ffm_bit_precision, spec->ffm_k);
        for (uint32_t f = 0; f < spec->n_fields; f++) {
            for (uint32_t j = spec->field_off[f]; j < spec->field_off[f + 1]; j++) {
                uint32_t ns = spec->field_ns[j];
                uint32_t n = read_namespace(rec, ns, spec->ns_is_f32 ? spec->ns_is_f32[ns] : 0, buf_a, FWO_MAX_NS_FEATS);
                if (n == UINT32_MAX || n_ffm + n > ffm_cap) return -1;
                for (uint32_t i = 0; i < n; i++) {
                    ffm[n_ffm].hash = buf_a[i].hash & ffm_mask;
                    ffm[n_ffm].value = buf_a[i].value;
                    ffm[n_ffm].contra_field_index = f * spec->ffm_k;
                    n_ffm++;
                }
            }
        }
    }
    *n_lr_out = n_lr;
    *n_ffm_out = n_ffm;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Regressor: block chain LR -> [FFM -> Triangle] -> [head] -> Sigmoid  (regressor.rs:173-330)
 * ---------------------------------------------------------------------------------------- */
typedef struct { float w; float acc; } lr_cell; /* WeightAndOptimizerData, block_helpers.rs:23-28 */

typedef struct {
    uint32_t n_in, n_out, relu, init;
    float maxnorm;
    float *w;   /* [(n_in+1)*n_out], row-major per neuron, bias at n_in*n_out + j (block_neural.rs:83-86) */
    float *acc;
    float *in;  /* forward input copy  [n_in]  */
    float *out; /* forward output      [n_out] (post-activation when relu) */
    float *mask; /* relu mask          [n_out] */
} nn_layer;

struct fwo_regressor {
    fwo_model_desc d;
    fwo_opt opt_lr, opt_ffm, opt_nn;
    uint32_t lr_len, ffm_len, F, k, Fk;
    lr_cell *lr;
    float *ffm_w, *ffm_acc;
    uint32_t n_layers; /* hidden + final neuron, 0 if no head */
    nn_layer layers[FWO_MAX_NN_LAYERS + 1];
};

static inline float logistic(float t) { return 1.0f / (1.0f + expf(-t)); } /* block_loss_functions.rs:15-17 */

/* xoshiro/normal init is not reproducible without the Rust crates (parity unpinned); a simple
 * deterministic stand-in is used for Hu/Xavier so the head is not degenerate.  Parity runs
 * import identical weights on both sides. */
static float standin_uniform(uint64_t *s)
{
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (float)((*s >> 40) & 0xFFFFFF) / 16777216.0f;
}

fwo_regressor *fwo_regressor_new(const fwo_model_desc *desc)
{
    fwo_regressor *r = (fwo_regressor *)calloc(1, sizeof(*r));
    r->d = *desc;
    opt_init(&r->opt_lr, desc->optimizer, desc->learning_rate, desc->power_t, desc->init_acc_gradient);          /* block_lr.rs:63-65 */
    opt_init(&r->opt_ffm, desc->optimizer, desc->ffm_learning_rate, desc->ffm_power_t, desc->ffm_init_acc_gradient); /* block_ffm.rs:87-91 */
    opt_init(&r->opt_nn, desc->optimizer, desc->nn_learning_rate, desc->nn_power_t, desc->nn_init_acc_gradient);  /* block_neural.rs:108-109 */
    r->F = desc->ffm_num_fields; r->k = desc->ffm_k; r->Fk = r->F * r->k;

    /* BlockLR::allocate_and_init_weights, block_lr.rs:97-105 */
    r->lr_len = 1u << desc->bit_precision;
    r->lr = (lr_cell *)malloc(sizeof(lr_cell) * (size_t)r->lr_len);
    for (uint32_t i = 0; i < r->lr_len; i++) { r->lr[i].w = 0.0f; r->lr[i].acc = opt_initial_data(&r->opt_lr); }

    /* BlockFFM::allocate_and_init_weights, block_ffm.rs:784-829 */
    if (desc->ffm_k > 0) {
        r->ffm_len = (1u << desc->ffm_bit_precision) + r->Fk; /* block_ffm.rs:93-94 */
        r->ffm_w = (float *)malloc(sizeof(float) * (size_t)r->ffm_len);
        r->ffm_acc = (float *)malloc(sizeof(float) * (size_t)r->ffm_len);
        float init_acc = opt_initial_data(&r->opt_ffm);
        if (desc->ffm_init_width == 0.0f) {
            float one_over_k_root = 1.0f / sqrtf((float)desc->ffm_k) / 50.0f;
            for (uint32_t i = 0; i < r->ffm_len; i++) {
                r->ffm_w[i] = (1.0f * fwo_merand48((uint64_t)r->ffm_len + (uint64_t)i) - 0.5f) * one_over_k_root;
                r->ffm_acc[i] = init_acc;
            }
        } else {
            float zero_half_band_width = desc->ffm_init_width * desc->ffm_init_zero_band * 0.5f;
            float band_width = desc->ffm_init_width * (1.0f - desc->ffm_init_zero_band);
            for (uint32_t i = 0; i < r->ffm_len; i++) {
                float w = fwo_merand48((uint64_t)i) * band_width - band_width * 0.5f;
                if (w > 0.0f) w += zero_half_band_width; else w -= zero_half_band_width;
                w += desc->ffm_init_center;
                r->ffm_w[i] = w;
                r->ffm_acc[i] = init_acc;
            }
        }
    }

    /* head: topology "one" (regressor.rs:191-320) */
    if (desc->nn_num_layers > 0 && desc->graph == FWO_GRAPH_REGRESSOR) {
        uint32_t x_len = desc->num_combos + (desc->ffm_k > 0 ? r->F * (r->F + 1) / 2 : 0);
        uint32_t n_in = x_len;
        uint64_t seed = 12345;
        for (uint32_t l = 0; l <= desc->nn_num_layers; l++) {
            nn_layer *L = &r->layers[l];
            int final = (l == desc->nn_num_layers);
            L->n_in = final ? n_in + x_len : n_in; /* join [h, x] (regressor.rs:303-307) */
            L->n_out = final ? 1 : desc->nn_width[l];
            L->relu = final ? 0 : desc->nn_relu[l];
            L->init = final ? FWO_NN_INIT_ONE : desc->nn_init[l]; /* regressor.rs:308-315 */
            L->maxnorm = final ? 0.0f : desc->nn_maxnorm[l];
            size_t len = (size_t)(L->n_in + 1) * L->n_out;
            L->w = (float *)malloc(sizeof(float) * len);
            L->acc = (float *)malloc(sizeof(float) * len);
            L->in = (float *)malloc(sizeof(float) * L->n_in);
            L->out = (float *)malloc(sizeof(float) * L->n_out);
            L->mask = (float *)malloc(sizeof(float) * L->n_out);
            size_t bias_offset = (size_t)L->n_in * L->n_out;
            for (size_t i = 0; i < len; i++) { L->w[i] = 1.0f; L->acc[i] = opt_initial_data(&r->opt_nn); }
            if (L->init == FWO_NN_INIT_ZERO) for (size_t i = 0; i < len; i++) L->w[i] = 0.0f;
            if (L->init == FWO_NN_INIT_HU || L->init == FWO_NN_INIT_XAVIER) {
                /* stand-in (see above): uniform with the same variance as the reference's law */
                double sd = L->init == FWO_NN_INIT_HU ? sqrt(2.0 / L->n_in) : sqrt(2.0 / (double)bias_offset);
                double bound = sd * sqrt(3.0);
                for (size_t i = 0; i < bias_offset; i++)
                    L->w[i] = (float)((2.0 * standin_uniform(&seed) - 1.0) * bound);
            }
            for (uint32_t j = 0; j < L->n_out; j++) L->w[bias_offset + j] = 0.0f; /* block_neural.rs:409-412 */
            n_in = L->n_out;
        }
        r->n_layers = desc->nn_num_layers + 1;
    }
    return r;
}

void fwo_regressor_free(fwo_regressor *r)
{
    if (!r) return;
    free(r->lr); free(r->ffm_w); free(r->ffm_acc);
    for (uint32_t l = 0; l < r->n_layers; l++) {
        free(r->layers[l].w); free(r->layers[l].acc); free(r->layers[l].in);
        free(r->layers[l].out); free(r->layers[l].mask);
    }
    free(r);
}

uint32_t fwo_lr_len(const fwo_regressor *r) { return r->lr_len; }
float *fwo_lr_table(fwo_regressor *r) { return (float *)r->lr; }
uint32_t fwo_ffm_len(const fwo_regressor *r) { return r->ffm_len; }
float *fwo_ffm_weights(fwo_regressor *r) { return r->ffm_w; }
float *fwo_ffm_acc(fwo_regressor *r) { return r->ffm_acc; }
uint32_t fwo_nn_layer_count(const fwo_regressor *r) { return r->n_layers; }
uint32_t fwo_nn_layer_len(const fwo_regressor *r, uint32_t l) { return (r->layers[l].n_in + 1) * r->layers[l].n_out; }
float *fwo_nn_weights(fwo_regressor *r, uint32_t l) { return r->layers[l].w; }
float *fwo_nn_acc(fwo_regressor *r, uint32_t l) { return r->layers[l].acc; }
const float *fwo_lut(const fwo_regressor *r, int which)
{
    return which == 0 ? r->opt_lr.lut : which == 1 ? r->opt_ffm.lut : r->opt_nn.lut;
}

/* per-thread scratch (the reference keeps these on the stack / in the PortBuffer tape) */
typedef struct {
    float *contra;  /* F*F*k  (block_ffm.rs:153) */
    float *local;   /* n*F*k  (block_ffm.rs:294-312 local_data_ffm_values) */
    size_t local_cap;
    float *out;     /* F*F   FFM outputs / their gradients on the way back */
    float *x;       /* sigmoid / head input: [lr outs, triangle] and its gradient */
    float *errs;    /* head scratch */
    size_t x_cap;
    size_t contra_cap, out_cap;
} scratch_t;

static __thread scratch_t tls_scratch;

static scratch_t *get_scratch(const fwo_regressor *r, uint32_t n_ffm)
{
    scratch_t *s = &tls_scratch;
    size_t need_local = (size_t)n_ffm * r->Fk + 1;
    /* the scratch is per thread, shared by every regressor that thread drives: grow on demand */
    size_t need_contra = (size_t)r->F * r->Fk + 8, need_out = (size_t)r->F * r->F + 8;
    if (need_contra > s->contra_cap) {
        free(s->contra);
        s->contra_cap = need_contra;
        s->contra = (float *)malloc(sizeof(float) * need_contra);
    }
    if (need_out > s->out_cap) {
        free(s->out);
        s->out_cap = need_out;
        s->out = (float *)malloc(sizeof(float) * need_out);
    }
    if (need_local > s->local_cap) {
        free(s->local);
        s->local_cap = need_local * 2;
        s->local = (float *)malloc(sizeof(float) * s->local_cap);
    }
    size_t xl = (size_t)r->d.num_combos + (size_t)r->F * r->F + 16;
    size_t need_x = xl;
    for (uint32_t l = 0; l < r->n_layers; l++) if (r->layers[l].n_in + 16 > need_x) need_x = r->layers[l].n_in + 16;
    if (need_x > s->x_cap) {
        free(s->x); free(s->errs);
        s->x_cap = need_x;
        s->x = (float *)malloc(sizeof(float) * need_x);
        s->errs = (float *)malloc(sizeof(float) * need_x);
    }
    return s;
}

/* BlockSigmoid (block_loss_functions.rs:105-153): returns p, *g = general gradient */
static float sigmoid_block(const float *inputs, uint32_t n, float label, float importance, float *g)
{
    float wsum = 0.0f;
    for (uint32_t i = 0; i < n; i++) wsum += inputs[i]; /* .iter().sum(), :116-120 */
    float p;
    if (isnan(wsum)) { p = logistic(0.0f); *g = 0.0f; }
    else if (wsum < -50.0f) { p = logistic(-50.0f); *g = 0.0f; }
    else if (wsum > 50.0f) { p = logistic(50.0f); *g = 0.0f; }
    else { p = logistic(wsum); *g = -(label - p) * importance; }
    return p;
}

/* One BlockNeuronLayer forward (+ BlockRELU when the layer has one): block_neural.rs:196-222, block_relu.rs:79-99.
 * The layer's input must already be in L->in. */
static void layer_forward(nn_layer *L)
{
    size_t bias_offset = (size_t)L->n_in * L->n_out;
    for (uint32_t j = 0; j < L->n_out; j++) {
        /* output = bias; sgemv('T') adds W x (block_neural.rs:207-220).  MKL's internal
         * summation order is unspecified; sequential here (pinned to 5e-6 only). */
        float acc = 0.0f;
        const float *wj = L->w + (size_t)j * L->n_in;
        for (uint32_t i = 0; i < L->n_in; i++) acc += wj[i] * L->in[i];
        float y = L->w[bias_offset + j] + acc;
        if (L->relu) { /* block_relu.rs:88-97: w < 0 -> 0 (mask 0) else w (mask 1) */
            if (y < 0.0f) { L->out[j] = 0.0f; L->mask[j] = 0.0f; }
            else { L->out[j] = y; L->mask[j] = 1.0f; }
        } else {
            L->out[j] = y; L->mask[j] = 1.0f;
        }
    }
}

/* One BlockNeuronLayer backward (block_neural.rs:252-341): up[n_out] = gradient of the layer's outputs (after the
 * layer's own relu backward), out_err[n_in] receives the gradient of its inputs (from the pre-update weights). */
static void layer_backward(fwo_opt *opt, nn_layer *L, const float *up, float *out_err, uint64_t example_number)
{
    for (uint32_t i = 0; i < L->n_in; i++) out_err[i] = 0.0f;
    size_t bias_offset = (size_t)L->n_in * L->n_out;
    for (uint32_t j = 0; j < L->n_out; j++) {
        float general_gradient = up[j] * 1.0f; /* dropout_inv == 1 */
        if (general_gradient == 0.0f) continue;
        size_t j_offset = (size_t)j * L->n_in;
        for (uint32_t i = 0; i < L->n_in; i++) {
            float feature_value = L->in[i];
            float gradient = general_gradient * feature_value;
            float update = opt_update(opt, gradient, &L->acc[i + j_offset]);
            out_err[i] += L->w[i + j_offset] * general_gradient;
            L->w[i + j_offset] -= update;
        }
        {
            float gradient = general_gradient * 1.0f;
            float update = opt_update(opt, gradient, &L->acc[bias_offset + j]);
            L->w[bias_offset + j] -= update;
        }
        if (L->maxnorm != 0.0f && example_number % 10 == 0) { /* block_neural.rs:307-320 */
            float wsq = 0.000001f;
            for (uint32_t i = 0; i < L->n_in; i++) { float w = L->w[i + j_offset]; wsq += w * w; }
            float norm = sqrtf(wsq);
            if (norm > L->maxnorm) {
                float scaling = L->maxnorm / norm;
                for (uint32_t i = 0; i < L->n_in; i++) L->w[i + j_offset] *= scaling;
            }
        }
    }
}

/* Head forward (training and predict share it: block_neural.rs:196-222, block_relu.rs:79-99).
 * x: [x_len] input; returns the single output of the final neuron. */
static float head_forward(fwo_regressor *r, const float *x, uint32_t x_len)
{
    const float *in = x;
    uint32_t n_in = x_len;
    for (uint32_t l = 0; l < r->n_layers; l++) {
        nn_layer *L = &r->layers[l];
        int final = (l + 1 == r->n_layers);
        memcpy(L->in, in, sizeof(float) * n_in);
        if (final) memcpy(L->in + n_in, x, sizeof(float) * x_len); /* join [h, x] */
        layer_forward(L);
        in = L->out;
        n_in = L->n_out;
    }
    return r->layers[r->n_layers - 1].out[0];
}

/* Head backward (block_neural.rs:252-341, block_relu.rs:101-108, block_misc.rs:452-473).
 * g_out: gradient at the final neuron's output.  Writes d_x[x_len]. */
static void head_backward(fwo_regressor *r, float g_out, uint32_t x_len, float *d_x, float *errs, uint64_t example_number)
{
    float grad_scalar = g_out;
    float *up = &grad_scalar; /* upstream gradient vector of the current layer's outputs */
    float *direct = NULL;     /* gradient of the x copy that feeds the final neuron */
    static __thread float *gbuf[FWO_MAX_NN_LAYERS + 1];
    static __thread size_t gcap[FWO_MAX_NN_LAYERS + 1];
    for (int l = (int)r->n_layers - 1; l >= 0; l--) {
        nn_layer *L = &r->layers[l];
        if (gcap[l] < L->n_in + 1) { free(gbuf[l]); gcap[l] = L->n_in + 16; gbuf[l] = (float *)malloc(sizeof(float) * gcap[l]); }
        float *out_err = gbuf[l];
        layer_backward(&r->opt_nn, L, up, out_err, example_number);
        if (l == (int)r->n_layers - 1) {
            /* final neuron's inputs are [h, x]: split */
            uint32_t h_len = L->n_in - x_len;
            direct = errs;
            memcpy(direct, out_err + h_len, sizeof(float) * x_len);
        }
        if (l > 0) {
            /* the previous layer's relu backward: d_in = mask * d_out (block_relu.rs:101-108) */
            nn_layer *P = &r->layers[l - 1];
            for (uint32_t i = 0; i < P->n_out; i++) out_err[i] = P->mask[i] * out_err[i];
            up = out_err;
        } else {
            /* copy block backward: d_x = d(first output) + d(second output) (block_misc.rs:452-473) */
            for (uint32_t i = 0; i < x_len; i++) d_x[i] = out_err[i] + direct[i];
        }
    }
}

/* Test hook: ONE neuron layer (optionally followed by a relu) in isolation, driven the way the reference's own unit
 * tests drive it (block_neural.rs:507-581, block_relu.rs:156-173): constant input x, an observe block that returns
 * the outputs and feeds back d_out as the gradient.  outs: [n_steps][n_out] forward outputs of each step (every step
 * updates); d_in: [n_in] input gradient of the last step.  Pins layer_forward / layer_backward to those goldens. */
int fwo_test_neuron_layer(uint32_t optimizer, float lr, float power_t, float init_acc, uint32_t n_in, uint32_t n_out, uint32_t init,
                          uint32_t relu, const float *x, const float *d_out, uint32_t n_steps, float *outs, float *d_in)
{
    fwo_opt opt;
    opt_init(&opt, optimizer, lr, power_t, init_acc);
    nn_layer L;
    memset(&L, 0, sizeof(L));
    L.n_in = n_in; L.n_out = n_out; L.relu = relu; L.init = init;
    size_t len = (size_t)(n_in + 1) * n_out;
    L.w = (float *)malloc(sizeof(float) * len); L.acc = (float *)malloc(sizeof(float) * len);
    L.in = (float *)malloc(sizeof(float) * n_in); L.out = (float *)malloc(sizeof(float) * n_out); L.mask = (float *)malloc(sizeof(float) * n_out);
    float *up = (float *)malloc(sizeof(float) * n_out), *err = (float *)malloc(sizeof(float) * (n_in + 1));
    for (size_t i = 0; i < len; i++) { L.w[i] = init == FWO_NN_INIT_ZERO ? 0.0f : 1.0f; L.acc[i] = opt_initial_data(&opt); }
    for (uint32_t j = 0; j < n_out; j++) L.w[(size_t)n_in * n_out + j] = 0.0f; /* block_neural.rs:409-412 */
    for (uint32_t s = 0; s < n_steps; s++) {
        memcpy(L.in, x, sizeof(float) * n_in);
        layer_forward(&L);
        for (uint32_t j = 0; j < n_out; j++) { outs[(size_t)s * n_out + j] = L.out[j]; up[j] = L.mask[j] * d_out[j]; }
        layer_backward(&opt, &L, up, err, s);
    }
    if (d_in) memcpy(d_in, err, sizeof(float) * n_in);
    free(L.w); free(L.acc); free(L.in); free(L.out); free(L.mask); free(up); free(err);
    return 0;
}

/* LR forward, block_lr.rs:28-47 */
static void lr_forward(const fwo_regressor *r, const fwo_feature_buffer *fb, float *lr_out)
{
    for (uint32_t c = 0; c < r->d.num_combos; c++) lr_out[c] = 0.0f;
    for (uint32_t i = 0; i < fb->n_lr; i++) {
        const fwo_lr_feat *f = &fb->lr[i];
        lr_out[f->combo_index] += r->lr[f->hash].w * f->value;
    }
}

/* Triangle forward, block_misc.rs:862-884 */
static uint32_t triangle_forward(const float *sq, uint32_t F, float *tri)
{
    uint32_t o = 0;
    for (uint32_t i = 0; i < F; i++) {
        for (uint32_t j = 0; j < i; j++) tri[o++] = sq[i * F + j] * 2.0f;
        tri[o++] = sq[i * F + i];
    }
    return o;
}

/* The training-order chain (forward_backward of every block). */
float fwo_forward_backward(fwo_regressor *r, const fwo_feature_buffer *fb, int update)
{
    const uint32_t F = r->F, k = r->k, Fk = r->Fk;
    const int has_ffm = r->d.ffm_k > 0;
    const int block_only = r->d.graph == FWO_GRAPH_FFM_BLOCK_ONLY;
    scratch_t *s = get_scratch(r, fb->n_ffm);
    float *x = s->x;
    uint32_t n_lr_out = block_only ? 0 : r->d.num_combos;

    if (!block_only) lr_forward(r, fb, x); /* BlockLR::forward_backward, block_lr.rs:123-133 */

    float *out = s->out;
    float *contra = s->contra;
    float *local = s->local;
    if (has_ffm) { /* BlockFFM::forward_backward, block_ffm.rs:122-261 */
        const float *W = r->ffm_w;
        for (uint32_t i = 0; i < F * F; i++) out[i] = 0.0f;
        if (fb->n_ffm) __builtin_prefetch(&contra[fb->ffm[0].contra_field_index], 1, 3); /* block_ffm.rs:163 */
        uint32_t idx = 0;
        for (uint32_t field = 0; field < F; field++) { /* :165-217 */
            uint32_t field_k = field * k;
            if (idx >= fb->n_ffm || fb->ffm[idx].contra_field_index > field_k) {
                size_t off = field_k;
                for (uint32_t z = 0; z < F; z++) { for (uint32_t q = 0; q < k; q++) contra[off + q] = 0.0f; off += Fk; }
                continue;
            }
            int first = 1;
            while (idx < fb->n_ffm && fb->ffm[idx].contra_field_index == field_k) {
                /* the reference's software prefetches (block_ffm.rs:163, 184, 194, 205): hints only, the arithmetic
                 * below is untouched.  :184 fetches the NEXT feature's row head, :194/:205 the next k-block of this row. */
                if (idx + 1 < fb->n_ffm) __builtin_prefetch(&W[fb->ffm[idx + 1].hash], 0, 3);
                const fwo_ffm_feat *ft = &fb->ffm[idx];
                float v = ft->value;
                size_t fi = ft->hash, off = field_k;
                if (first) {
                    for (uint32_t z = 0; z < F; z++) { __builtin_prefetch(&W[fi + k], 0, 3); for (uint32_t q = 0; q < k; q++) contra[off + q] = W[fi + q] * v; off += Fk; fi += k; }
                    first = 0;
                } else {
                    for (uint32_t z = 0; z < F; z++) { __builtin_prefetch(&W[fi + k], 0, 3); for (uint32_t q = 0; q < k; q++) contra[off + q] += W[fi + q] * v; off += Fk; fi += k; }
                }
                idx++;
            }
        }
        size_t vo = 0;
        for (uint32_t e = 0; e < fb->n_ffm; e++) { /* :219-261 */
            const fwo_ffm_feat *ft = &fb->ffm[e];
            float v = ft->value;
            size_t fi = ft->hash;
            size_t cfi = ft->contra_field_index;
            size_t contra_offset = cfi * F;
            size_t contra_offset2 = contra_offset / k;
            size_t vv = 0;
            for (uint32_t z = 0; z < F; z++) {
                float correction = 0.0f;
                size_t vfi = fi + vv, vco = contra_offset + vv;
                if (vv == cfi) {
                    for (uint32_t q = 0; q < k; q++) {
                        float w = W[vfi + q];
                        float cw = contra[vco + q] - w * v;
                        float gradient = v * cw;
                        local[vo + q] = gradient;
                        correction += w * gradient;
                    }
                } else {
                    for (uint32_t q = 0; q < k; q++) {
                        float cw = contra[vco + q];
                        float gradient = v * cw;
                        local[vo + q] = gradient;
                        float w = W[vfi + q];
                        correction += w * gradient;
                    }
                }
                out[contra_offset2 + z] += correction * 0.5f;
                vv += k;
                vo += k;
            }
        }
    }

    /* downstream */
    float p, g;
    uint32_t x_len = n_lr_out;
    if (has_ffm && !block_only) x_len += triangle_forward(out, F, x + n_lr_out);
    if (block_only) {
        p = sigmoid_block(out, F * F, fb->label, fb->example_importance, &g);
        for (uint32_t i = 0; i < F * F; i++) out[i] = g; /* :149-151 */
    } else if (r->n_layers > 0) {
        float y = head_forward(r, x, x_len);
        p = sigmoid_block(&y, 1, fb->label, fb->example_importance, &g);
        if (update) head_backward(r, g, x_len, x, s->errs, fb->example_number);
    } else {
        p = sigmoid_block(x, x_len, fb->label, fb->example_importance, &g);
        for (uint32_t i = 0; i < x_len; i++) x[i] = g;
    }

    if (!update) return p;

    if (has_ffm) {
        if (!block_only) { /* Triangle backward, block_misc.rs:814-833 */
            const float *tri = x + n_lr_out;
            uint32_t o = 0;
            for (uint32_t i = 0; i < F; i++)
                for (uint32_t j = 0; j <= i; j++) { out[i * F + j] = tri[o]; out[j * F + i] = tri[o]; o++; }
        }
        /* FFM update, block_ffm.rs:265-288 */
        size_t li = 0;
        float *W = r->ffm_w, *A = r->ffm_acc;
        for (uint32_t e = 0; e < fb->n_ffm; e++) {
            const fwo_ffm_feat *ft = &fb->ffm[e];
            size_t fi = ft->hash;
            size_t co = ((size_t)ft->contra_field_index * F) / k;
            for (uint32_t z = 0; z < F; z++) {
                float general_gradient = out[co + z];
                for (uint32_t q = 0; q < k; q++) {
                    float gradient = general_gradient * local[li];
                    float upd = opt_update(&r->opt_ffm, gradient, &A[fi]);
                    W[fi] -= upd;
                    li++; fi++;
                }
            }
        }
    }
    if (!block_only) { /* LR update, block_lr.rs:135-151 */
        for (uint32_t i = 0; i < fb->n_lr; i++) {
            const fwo_lr_feat *f = &fb->lr[i];
            float gradient = x[f->combo_index] * f->value;
            float upd = opt_update(&r->opt_lr, gradient, &r->lr[f->hash].acc);
            r->lr[f->hash].w -= upd;
        }
    }
    return p;
}

/* SSE horizontal add as the reference writes it (block_ffm.rs:106-114):
 * r2 = r4 + movehl(r4) ; r1 = r2[0] + r2[1]  ==> (s0+s2) + (s1+s3) */
static inline float hadd4(const float s[4]) { return (s[0] + s[2]) + (s[1] + s[3]); }

/* Predict-order forward: BlockFFM::forward + prepare_contra_fields + calculate_interactions
 * (block_ffm.rs:316-440, 964-1104, 1107-1201), non-FMA variant (:936-947). */
static void ffm_forward_predict(const fwo_regressor *r, const fwo_feature_buffer *fb, float *out, float *contra)
{
    const uint32_t F = r->F, k = r->k, Fk = r->Fk;
    const float *W = r->ffm_w;
    for (uint32_t i = 0; i < F * F; i++) out[i] = 0.0f;
    uint32_t idx = 0;
    for (uint32_t field = 0; field < F; field++) {
        uint32_t field_k = field * k;
        size_t offset = (size_t)field_k * F;
        if (idx >= fb->n_ffm || fb->ffm[idx].contra_field_index > field_k) {
            for (uint32_t z = 0; z < Fk; z++) contra[offset + z] = 0.0f;
            continue;
        }
        uint32_t ffm_index = field * (F + 1);
        int first = 1;
        while (idx < fb->n_ffm && fb->ffm[idx].contra_field_index == field_k) {
            const fwo_ffm_feat *ft = &fb->ffm[idx];
            size_t fi = ft->hash;
            float v = ft->value;
            if (first) { /* :976-1029 */
                first = 0;
                if (v == 1.0f) for (uint32_t z = 0; z < Fk; z++) contra[offset + z] = W[fi + z];
                else for (uint32_t z = 0; z < Fk; z++) contra[offset + z] = W[fi + z] * v;
            } else if (v == 1.0f) { /* :1030-1057 */
                for (uint32_t z = 0; z < Fk; z++) contra[offset + z] = W[fi + z] + contra[offset + z];
            } else { /* :1058-1103 */
                for (uint32_t z = 0; z < Fk; z++) contra[offset + z] = W[fi + z] * v + contra[offset + z];
            }
            float correction = 0.0f; /* :416-424 */
            for (uint32_t q = 0; q < k; q++) { float w = W[fi + field_k + q]; correction += w * w; }
            out[ffm_index] -= correction * 0.5f * v * v;
            idx++;
        }
    }
    /* calculate_interactions :1107-1201 */
    const uint32_t LANES = 8;
    uint32_t k_end = k - k % LANES;
    for (uint32_t f1 = 0; f1 < F; f1++) {
        size_t f1_offset = (size_t)f1 * Fk;
        size_t f1_ffmk = (size_t)f1 * k;
        size_t f1_offset_ffmk = f1_offset + f1_ffmk;
        float cf = 0.0f;
        {
            const float *cp = contra + f1_offset_ffmk;
            if (k == LANES) {
                float s4[4];
                for (int l = 0; l < 4; l++) s4[l] = cp[l] * cp[l] + cp[4 + l] * cp[4 + l];
                cf = hadd4(s4);
            } else {
                for (uint32_t q = 0; q < k_end; q += LANES) {
                    float s4[4];
                    for (int l = 0; l < 4; l++) s4[l] = cp[q + l] * cp[q + l] + cp[q + 4 + l] * cp[q + 4 + l];
                    cf += hadd4(s4);
                }
                for (uint32_t q = k_end; q < k; q++) cf += contra[f1_offset_ffmk + q] * contra[f1_offset_ffmk + q];
            }
        }
        out[f1 * F + f1] += cf * 0.5f;
        size_t f2_offset_ffmk = f1_offset + f1_ffmk;
        for (uint32_t f2 = f1 + 1; f2 < F; f2++) {
            f2_offset_ffmk += Fk;
            f1_offset_ffmk += k;
            const float *c1 = contra + f1_offset_ffmk, *c2 = contra + f2_offset_ffmk;
            float c = 0.0f;
            if (k == LANES) {
                float s4[4];
                for (int l = 0; l < 4; l++) s4[l] = c1[l] * c2[l] + c1[4 + l] * c2[4 + l];
                c = hadd4(s4);
            } else {
                for (uint32_t q = 0; q < k_end; q += LANES) {
                    float s4[4];
                    for (int l = 0; l < 4; l++) s4[l] = c1[q + l] * c2[q + l] + c1[q + 4 + l] * c2[q + 4 + l];
                    c += hadd4(s4);
                }
                for (uint32_t q = k_end; q < k; q++) c += c1[q] * c2[q];
            }
            c *= 0.5f;
            out[f1 * F + f2] += c;
            out[f2 * F + f1] += c;
        }
    }
}

/* Regressor::predict, regressor.rs:381-395 */
float fwo_predict(fwo_regressor *r, const fwo_feature_buffer *fb)
{
    const int has_ffm = r->d.ffm_k > 0;
    const int block_only = r->d.graph == FWO_GRAPH_FFM_BLOCK_ONLY;
    scratch_t *s = get_scratch(r, fb->n_ffm);
    float *x = s->x;
    uint32_t n_lr_out = block_only ? 0 : r->d.num_combos;
    if (!block_only) lr_forward(r, fb, x);
    if (has_ffm) ffm_forward_predict(r, fb, s->out, s->contra);
    float g;
    if (block_only) return sigmoid_block(s->out, r->F * r->F, fb->label, fb->example_importance, &g);
    uint32_t x_len = n_lr_out;
    if (has_ffm) x_len += triangle_forward(s->out, r->F, x + n_lr_out);
    if (r->n_layers > 0) {
        float y = head_forward(r, x, x_len);
        return sigmoid_block(&y, 1, fb->label, fb->example_importance, &g);
    }
    return sigmoid_block(x, x_len, fb->label, fb->example_importance, &g);
}

/* Regressor::learn, regressor.rs:356-379 */
float fwo_learn(fwo_regressor *r, const fwo_feature_buffer *fb, int update)
{
    int upd = update && (fb->example_importance != 0.0f);
    if (!upd) return fwo_predict(r, fb);
    return fwo_forward_backward(r, fb, 1);
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers
 * ---------------------------------------------------------------------------------------- */
void fwo_learn_batch_sequential(fwo_regressor *r, const fwo_batch *b, float *preds, int update)
{
    uint32_t max_lr = 0, max_ffm = 0;
    for (uint32_t e = 0; e < b->n_examples; e++) {
        uint32_t a = b->lr_off[e + 1] - b->lr_off[e], c = b->ffm_off ? b->ffm_off[e + 1] - b->ffm_off[e] : 0;
        if (a > max_lr) max_lr = a;
        if (c > max_ffm) max_ffm = c;
    }
    fwo_lr_feat *lr = (fwo_lr_feat *)malloc(sizeof(fwo_lr_feat) * (max_lr + 1));
    fwo_ffm_feat *ffm = (fwo_ffm_feat *)malloc(sizeof(fwo_ffm_feat) * (max_ffm + 1));
    for (uint32_t e = 0; e < b->n_examples; e++) {
        fwo_feature_buffer fb;
        fb.label = b->labels[e];
        fb.example_importance = b->importance[e];
        fb.example_number = e;
        fb.n_lr = b->lr_off[e + 1] - b->lr_off[e];
        for (uint32_t i = 0; i < fb.n_lr; i++) {
            uint32_t j = b->lr_off[e] + i;
            lr[i].hash = b->lr_hash[j]; lr[i].value = b->lr_val[j]; lr[i].combo_index = b->lr_combo[j];
        }
        fb.lr = lr;
        fb.n_ffm = b->ffm_off ? b->ffm_off[e + 1] - b->ffm_off[e] : 0;
        for (uint32_t i = 0; i < fb.n_ffm; i++) {
            uint32_t j = b->ffm_off[e] + i;
            ffm[i].hash = b->ffm_hash[j]; ffm[i].value = b->ffm_val[j];
            ffm[i].contra_field_index = b->ffm_field[j] * r->d.ffm_k;
        }
        fb.ffm = ffm;
        float p = fwo_learn(r, &fb, update);
        if (preds) preds[e] = p;
    }
    free(lr); free(ffm);
}

typedef struct {
    fwo_regressor *r;
    const fwo_translate_spec *spec;
    const uint32_t *records;
    const uint64_t *rec_off;
    uint64_t n_records;
    uint64_t *next;
    float *preds;
} hog_arg;

#define HOG_CHUNK 256
static void *hog_worker(void *pv)
{
    hog_arg *a = (hog_arg *)pv;
    const uint32_t cap = 8192;
    fwo_lr_feat *lr = (fwo_lr_feat *)malloc(sizeof(fwo_lr_feat) * cap);
    fwo_ffm_feat *ffm = (fwo_ffm_feat *)malloc(sizeof(fwo_ffm_feat) * cap);
    for (;;) {
        /* the reference hands records over an mpsc channel (hogwild.rs:89-103); a shared
         * counter is the same work distribution without the Mutex<Receiver>. */
        uint64_t start = __atomic_fetch_add(a->next, HOG_CHUNK, __ATOMIC_RELAXED);
        if (start >= a->n_records) break;
        uint64_t end = start + HOG_CHUNK < a->n_records ? start + HOG_CHUNK : a->n_records;
        for (uint64_t e = start; e < end; e++) {
            fwo_feature_buffer fb;
            uint32_t n_lr = 0, n_ffm = 0;
            if (fwo_translate(a->spec, a->records + a->rec_off[e], lr, cap, &n_lr, ffm, cap, &n_ffm,
                              &fb.label, &fb.example_importance) != 0) continue;
            fb.example_number = e; fb.n_lr = n_lr; fb.lr = lr; fb.n_ffm = n_ffm; fb.ffm = ffm;
            float p = fwo_learn(a->r, &fb, 1);
            if (a->preds) a->preds[e] = p;
        }
    }
    free(lr); free(ffm);
    return NULL;
}

double fwo_hogwild_run(fwo_regressor *r, const fwo_translate_spec *spec, const uint32_t *records,
                       const uint64_t *rec_off, uint64_t n_records, uint32_t n_threads, float *preds)
{
    uint64_t next = 0;
    hog_arg a = { r, spec, records, rec_off, n_records, &next, preds };
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (n_threads <= 1) {
        hog_worker(&a);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
        for (uint32_t i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, hog_worker, &a);
        for (uint32_t i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------
 * Wave emulation (analysis tool, no reference counterpart): what a device that keeps `wave`
 * examples in flight computes.  All examples of a wave are scored against the same weight
 * snapshot; their updates are then applied
 *   mode 0: one after another, each slot's accumulator growing with every hit (what lock-free
 *           per-example atomics do -- "Hogwild on device");
 *   mode 1: aggregated per slot over the wave: G = sum g_i, acc += sum g_i^2, one step with G.
 * Used on the CPU to choose the device semantics and the safe amount of concurrency; regressor
 * graph without head only.
 * ---------------------------------------------------------------------------------------- */
void fwo_learn_records_wave(fwo_regressor *r, const fwo_translate_spec *spec, const uint32_t *records,
                            const uint64_t *rec_off, uint64_t n_records, uint32_t wave, int mode, float *preds)
{
    /* mode bits: low 4 bits = update mode; bits 8.. = ramp divisor (0 = none): the wave grows as
     * examples_seen / ramp_div, capped at `wave` (the concurrency ramp of the device path). */
    const uint32_t ramp_div = (uint32_t)mode >> 8;
    mode &= 15;
    const uint32_t F = r->F, k = r->k, Fk = r->Fk;
    const uint32_t cap = 4096;
    fwo_lr_feat *lr = (fwo_lr_feat *)malloc(sizeof(fwo_lr_feat) * cap * (size_t)wave);
    fwo_ffm_feat *ffm = (fwo_ffm_feat *)malloc(sizeof(fwo_ffm_feat) * cap * (size_t)wave);
    uint32_t *n_lr = (uint32_t *)malloc(4 * (size_t)wave), *n_ffm = (uint32_t *)malloc(4 * (size_t)wave);
    float *gs = (float *)malloc(4 * (size_t)wave);
    size_t *loc_off = (size_t *)malloc(sizeof(size_t) * ((size_t)wave + 1));
    size_t loc_cap = 0;
    float *loc = NULL;
    float *gsum_ffm = NULL, *g2_ffm = NULL, *gsum_lr = NULL, *g2_lr = NULL;
    uint32_t *touched = NULL;
    size_t touched_cap = 0;
    if (mode == 1) {
        gsum_ffm = (float *)calloc(r->ffm_len + 1, 4); g2_ffm = (float *)calloc(r->ffm_len + 1, 4);
        gsum_lr = (float *)calloc(r->lr_len, 4); g2_lr = (float *)calloc(r->lr_len, 4);
    }
    uint32_t m = 0;
    for (uint64_t a = 0; a < n_records; a += m) {
        uint32_t wv = wave;
        if (ramp_div) { uint64_t lim = a / ramp_div; if (lim < 1) lim = 1; if (lim < wv) wv = (uint32_t)lim; }
        m = (uint32_t)((a + wv <= n_records) ? wv : n_records - a);
        /* phase 1: forward of the whole wave on the snapshot */
        size_t need = 0;
        for (uint32_t i = 0; i < m; i++) {
            fwo_feature_buffer fb;
            fwo_lr_feat *l = lr + (size_t)i * cap; fwo_ffm_feat *f = ffm + (size_t)i * cap;
            if (fwo_translate(spec, records + rec_off[a + i], l, cap, &n_lr[i], f, cap, &n_ffm[i], &fb.label, &fb.example_importance) != 0) { n_lr[i] = n_ffm[i] = 0; }
            loc_off[i] = need; need += (size_t)n_ffm[i] * Fk;
        }
        loc_off[m] = need;
        if (need > loc_cap) { free(loc); loc_cap = need * 2 + 16; loc = (float *)malloc(4 * loc_cap); }
        for (uint32_t i = 0; i < m; i++) {
            fwo_feature_buffer fb;
            fb.label = (float)records[rec_off[a + i] + 1];
            memcpy(&fb.example_importance, &records[rec_off[a + i] + 2], 4);
            fb.example_number = a + i; fb.n_lr = n_lr[i]; fb.lr = lr + (size_t)i * cap; fb.n_ffm = n_ffm[i]; fb.ffm = ffm + (size_t)i * cap;
            float p = fwo_forward_backward(r, &fb, 0); /* training-order forward, no update */
            if (preds) preds[a + i] = p;
            float g;
            /* recompute g as the sigmoid block does */
            if (p != p) g = 0.0f; else g = -(fb.label - p) * fb.example_importance;
            /* clamp cases (|wsum| > 50) give g = 0 in the reference; p then equals logistic(+-50) */
            if (p <= 1.0f / (1.0f + expf(50.0f)) || p >= 1.0f / (1.0f + expf(-50.0f))) g = 0.0f;
            if (fb.example_importance == 0.0f) g = 0.0f;
            gs[i] = g;
            /* local gradients G_i (block_ffm.rs:219-261) were left in the thread-local scratch by forward_backward */
            memcpy(loc + loc_off[i], tls_scratch.local, 4 * (size_t)n_ffm[i] * Fk);
        }
        /* phase 2: updates */
        if (mode == 0) {
            for (uint32_t i = 0; i < m; i++) {
                float g = gs[i];
                if (g == 0.0f) continue;
                const float *G = loc + loc_off[i];
                size_t li = 0;
                for (uint32_t e = 0; e < n_ffm[i]; e++) {
                    size_t fi = ffm[(size_t)i * cap + e].hash;
                    for (uint32_t x = 0; x < Fk; x++) {
                        float gradient = g * G[li++];
                        float upd = opt_update(&r->opt_ffm, gradient, &r->ffm_acc[fi + x]);
                        r->ffm_w[fi + x] -= upd;
                    }
                }
                for (uint32_t e = 0; e < n_lr[i]; e++) {
                    const fwo_lr_feat *f = &lr[(size_t)i * cap + e];
                    float upd = opt_update(&r->opt_lr, g * f->value, &r->lr[f->hash].acc);
                    r->lr[f->hash].w -= upd;
                }
            }
        } else {
            size_t nt = 0;
            size_t max_touch = 0;
            for (uint32_t i = 0; i < m; i++) max_touch += (size_t)n_ffm[i] * Fk + n_lr[i];
            if (max_touch > touched_cap) { free(touched); touched_cap = max_touch * 2 + 16; touched = (uint32_t *)malloc(4 * touched_cap); }
            for (uint32_t i = 0; i < m; i++) {
                float g = gs[i];
                if (g == 0.0f) continue;
                const float *G = loc + loc_off[i];
                size_t li = 0;
                for (uint32_t e = 0; e < n_ffm[i]; e++) {
                    size_t fi = ffm[(size_t)i * cap + e].hash;
                    for (uint32_t x = 0; x < Fk; x++) {
                        float gradient = g * G[li++];
                        if (g2_ffm[fi + x] == 0.0f && gsum_ffm[fi + x] == 0.0f) touched[nt++] = (uint32_t)(fi + x);
                        gsum_ffm[fi + x] += gradient; g2_ffm[fi + x] += gradient * gradient;
                    }
                }
            }
            for (size_t t = 0; t < nt; t++) {
                uint32_t s = touched[t];
                if (g2_ffm[s] == 0.0f && gsum_ffm[s] == 0.0f) continue;
                float G = gsum_ffm[s], g2 = g2_ffm[s];
                gsum_ffm[s] = 0.0f; g2_ffm[s] = 0.0f;
                /* acc += sum g_i^2 ; step with the summed gradient */
                float new_acc = r->ffm_acc[s] + g2;
                r->ffm_acc[s] = new_acc;
                float step;
                if (r->opt_ffm.kind == FWO_OPT_ADAGRAD_LUT) step = G * r->opt_ffm.lut[f2bits(new_acc) >> 20];
                else if (r->opt_ffm.kind == FWO_OPT_ADAGRAD_FLEX) step = G * r->opt_ffm.lr * powf(new_acc, r->opt_ffm.minus_power_t);
                else step = G * r->opt_ffm.lr;
                r->ffm_w[s] -= step;
            }
            nt = 0;
            for (uint32_t i = 0; i < m; i++) {
                float g = gs[i];
                if (g == 0.0f) continue;
                for (uint32_t e = 0; e < n_lr[i]; e++) {
                    const fwo_lr_feat *f = &lr[(size_t)i * cap + e];
                    float gradient = g * f->value;
                    if (g2_lr[f->hash] == 0.0f && gsum_lr[f->hash] == 0.0f) touched[nt++] = f->hash;
                    gsum_lr[f->hash] += gradient; g2_lr[f->hash] += gradient * gradient;
                }
            }
            for (size_t t = 0; t < nt; t++) {
                uint32_t s = touched[t];
                if (g2_lr[s] == 0.0f && gsum_lr[s] == 0.0f) continue;
                float G = gsum_lr[s], g2 = g2_lr[s];
                gsum_lr[s] = 0.0f; g2_lr[s] = 0.0f;
                float new_acc = r->lr[s].acc + g2;
                r->lr[s].acc = new_acc;
                float step;
                if (r->opt_lr.kind == FWO_OPT_ADAGRAD_LUT) step = G * r->opt_lr.lut[f2bits(new_acc) >> 20];
                else if (r->opt_lr.kind == FWO_OPT_ADAGRAD_FLEX) step = G * r->opt_lr.lr * powf(new_acc, r->opt_lr.minus_power_t);
                else step = G * r->opt_lr.lr;
                r->lr[s].w -= step;
            }
        }
    }
    free(lr); free(ffm); free(n_lr); free(n_ffm); free(gs); free(loc_off); free(loc);
    free(gsum_ffm); free(g2_ffm); free(gsum_lr); free(g2_lr); free(touched);
    (void)F; (void)k;
}

/* ------------------------------------------------------------------------------------------
 * Head wave emulation (test tool, no reference counterpart): the batched semantics the device
 * gives a model with a dense head.  A sub-batch of `wave` examples is scored against ONE weight
 * snapshot; the dense layers then receive, per weight, G1 = sum_b g_bj x_bi and
 * G2 = sum_b (g_bj x_bi)^2 computed from the pre-update weights (block_neural.rs:266-305 per
 * example, summed), one optimizer step  acc += G2 ; w -= step(G1, acc);  the sparse LR / FFM
 * tables are updated example after example from the gradients the head's backward pass left
 * for each input (block_ffm.rs:265-288, block_lr.rs:135-151), using the local gradients of the
 * snapshot forward.  With wave == 1 this is fwo_learn, arithmetic included.
 * ---------------------------------------------------------------------------------------- */
void fwo_learn_records_head_wave(fwo_regressor *r, const fwo_translate_spec *spec, const uint32_t *records,
                                 const uint64_t *rec_off, uint64_t n_records, uint32_t wave, float *preds)
{
    if (r->n_layers == 0 || wave == 0) return;
    const uint32_t F = r->F, k = r->k, Fk = r->Fk, nl = r->n_layers;
    const uint32_t cap = 4096;
    const uint32_t n_lr_out = r->d.num_combos;
    const uint32_t x_len = n_lr_out + (r->d.ffm_k > 0 ? F * (F + 1) / 2 : 0);
    fwo_lr_feat *lr = (fwo_lr_feat *)malloc(sizeof(fwo_lr_feat) * cap * (size_t)wave);
    fwo_ffm_feat *ffm = (fwo_ffm_feat *)malloc(sizeof(fwo_ffm_feat) * cap * (size_t)wave);
    uint32_t *n_lr = (uint32_t *)malloc(4 * (size_t)wave), *n_ffm = (uint32_t *)malloc(4 * (size_t)wave);
    float *gs = (float *)malloc(4 * (size_t)wave);
    float *DX = (float *)malloc(4 * (size_t)wave * x_len);
    float *loc = (float *)malloc(4 * (size_t)wave * F * Fk + 16);
    float *in_s[FWO_MAX_NN_LAYERS + 1], *mask_s[FWO_MAX_NN_LAYERS + 1], *G1[FWO_MAX_NN_LAYERS + 1], *G2[FWO_MAX_NN_LAYERS + 1], *err[FWO_MAX_NN_LAYERS + 1];
    for (uint32_t l = 0; l < nl; l++) {
        nn_layer *L = &r->layers[l];
        size_t np = (size_t)(L->n_in + 1) * L->n_out;
        in_s[l] = (float *)malloc(4 * (size_t)wave * L->n_in);
        mask_s[l] = (float *)malloc(4 * (size_t)wave * L->n_out);
        G1[l] = (float *)calloc(np, 4); G2[l] = (float *)calloc(np, 4);
        err[l] = (float *)malloc(4 * ((size_t)L->n_in + 1));
    }
    uint32_t m = 0;
    for (uint64_t a = 0; a < n_records; a += m) {
        m = (uint32_t)((a + wave <= n_records) ? wave : n_records - a);
        /* phase 1: forward of every example on the snapshot */
        for (uint32_t i = 0; i < m; i++) {
            fwo_feature_buffer fb;
            fb.lr = lr + (size_t)i * cap; fb.ffm = ffm + (size_t)i * cap;
            if (fwo_translate(spec, records + rec_off[a + i], (fwo_lr_feat *)fb.lr, cap, &n_lr[i], (fwo_ffm_feat *)fb.ffm, cap, &n_ffm[i], &fb.label, &fb.example_importance) != 0) n_lr[i] = n_ffm[i] = 0;
            if (n_ffm[i] > F) { n_ffm[i] = 0; n_lr[i] = 0; } /* the tool covers single-valued fields (the fused device path) */
            fb.example_number = a + i; fb.n_lr = n_lr[i]; fb.n_ffm = n_ffm[i];
            float p = fwo_forward_backward(r, &fb, 0);
            if (preds) preds[a + i] = p;
            float g;
            { /* the sigmoid block's gradient for this example, from the final neuron's output */
                float y = r->layers[nl - 1].out[0];
                (void)sigmoid_block(&y, 1, fb.label, fb.example_importance, &g);
                if (fb.example_importance == 0.0f) g = 0.0f; /* regressor.rs:366-370 */
            }
            gs[i] = g;
            memcpy(loc + (size_t)i * F * Fk, tls_scratch.local, 4 * (size_t)n_ffm[i] * Fk);
            for (uint32_t l = 0; l < nl; l++) {
                nn_layer *L = &r->layers[l];
                memcpy(in_s[l] + (size_t)i * L->n_in, L->in, 4 * (size_t)L->n_in);
                memcpy(mask_s[l] + (size_t)i * L->n_out, L->mask, 4 * (size_t)L->n_out);
            }
        }
        /* phase 2: dense backward of every example on the snapshot, gradient sums per weight */
        for (uint32_t i = 0; i < m; i++) {
            float *dx = DX + (size_t)i * x_len;
            float g = gs[i];
            if (g == 0.0f) { for (uint32_t t = 0; t < x_len; t++) dx[t] = 0.0f; continue; }
            const float *up = &g;
            const float *direct = NULL;
            for (int l = (int)nl - 1; l >= 0; l--) {
                nn_layer *L = &r->layers[l];
                const float *in = in_s[l] + (size_t)i * L->n_in;
                float *oe = err[l];
                size_t bias_offset = (size_t)L->n_in * L->n_out;
                for (uint32_t t = 0; t < L->n_in; t++) oe[t] = 0.0f;
                for (uint32_t j = 0; j < L->n_out; j++) {
                    float gg = up[j];
                    if (gg == 0.0f) continue;
                    size_t jo = (size_t)j * L->n_in;
                    for (uint32_t t = 0; t < L->n_in; t++) {
                        float gradient = gg * in[t];
                        G1[l][jo + t] += gradient; G2[l][jo + t] += gradient * gradient;
                        oe[t] += L->w[jo + t] * gg;
                    }
                    G1[l][bias_offset + j] += gg; G2[l][bias_offset + j] += gg * gg;
                }
                if (l == (int)nl - 1) direct = oe + (L->n_in - x_len); /* final neuron's inputs are [h, x] */
                if (l > 0) {
                    nn_layer *P = &r->layers[l - 1];
                    const float *pm = mask_s[l - 1] + (size_t)i * P->n_out;
                    for (uint32_t t = 0; t < P->n_out; t++) oe[t] = pm[t] * oe[t]; /* block_relu.rs:101-108 */
                    up = oe;
                } else {
                    for (uint32_t t = 0; t < x_len; t++) dx[t] = oe[t] + direct[t]; /* block_misc.rs:452-473 */
                }
            }
        }
        /* phase 3: one optimizer step per dense weight */
        for (uint32_t l = 0; l < nl; l++) {
            nn_layer *L = &r->layers[l];
            size_t np = (size_t)(L->n_in + 1) * L->n_out;
            for (size_t t = 0; t < np; t++) {
                float g1 = G1[l][t], g2 = G2[l][t];
                if (g1 == 0.0f && g2 == 0.0f) continue;
                G1[l][t] = 0.0f; G2[l][t] = 0.0f;
                float step;
                if (r->opt_nn.kind == FWO_OPT_SGD) step = g1 * r->opt_nn.lr;
                else {
                    float new_acc = L->acc[t] + g2;
                    L->acc[t] = new_acc;
                    if (r->opt_nn.kind == FWO_OPT_ADAGRAD_LUT) step = g1 * r->opt_nn.lut[f2bits(new_acc) >> 20];
                    else { step = g1 * r->opt_nn.lr * powf(new_acc, r->opt_nn.minus_power_t); if (isnan(step) || isinf(step)) step = 0.0f; }
                }
                L->w[t] -= step;
            }
        }
        /* phase 4: sparse tables, example after example */
        for (uint32_t i = 0; i < m; i++) {
            if (gs[i] == 0.0f) continue;
            const float *dx = DX + (size_t)i * x_len;
            const float *G = loc + (size_t)i * F * Fk;
            const fwo_ffm_feat *fe = ffm + (size_t)i * cap;
            size_t li = 0;
            for (uint32_t e = 0; e < n_ffm[i]; e++) {
                size_t fi = fe[e].hash;
                uint32_t f = fe[e].contra_field_index / k;
                for (uint32_t z = 0; z < F; z++) {
                    uint32_t hi = f > z ? f : z, lo = f > z ? z : f;
                    float general_gradient = dx[n_lr_out + hi * (hi + 1) / 2 + lo]; /* block_misc.rs:814-833 */
                    for (uint32_t q = 0; q < k; q++) {
                        float gradient = general_gradient * G[li];
                        float upd = opt_update(&r->opt_ffm, gradient, &r->ffm_acc[fi]);
                        r->ffm_w[fi] -= upd;
                        li++; fi++;
                    }
                }
            }
            for (uint32_t e = 0; e < n_lr[i]; e++) {
                const fwo_lr_feat *f = &lr[(size_t)i * cap + e];
                float upd = opt_update(&r->opt_lr, dx[f->combo_index] * f->value, &r->lr[f->hash].acc);
                r->lr[f->hash].w -= upd;
            }
        }
    }
    for (uint32_t l = 0; l < nl; l++) { free(in_s[l]); free(mask_s[l]); free(G1[l]); free(G2[l]); free(err[l]); }
    free(lr); free(ffm); free(n_lr); free(n_ffm); free(gs); free(DX); free(loc);
}
