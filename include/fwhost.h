/*
 * fwhost.h -- host-side (CPU, C++ inside, C ABI outside) pieces either side of the GPU hot path:
 * hashing, the VW text parser, the .fwcache reader/writer and the regressor file format.  They restate the reference's host code
 * (parser.rs, cache.rs, persistence.rs, vwmap.rs) so that a maintainer can keep using the Rust
 * host unchanged; nothing here is on the GPU hot path and nothing here calls oracle/.
 */
#ifndef FWHOST_H
#define FWHOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fasthash murmur3::hash32_with_seed == MurmurHash3_x86_32 (call sites parser.rs:82-83, 382-385) */
uint32_t fwhost_murmur3_32(const void *key, size_t len, uint32_t seed);

/* ---- everything below returns 0 / a count on success and a negative value on failure with a message in err ----
 * Strings returned as char* are malloc'ed: release them with fwhost_free. */
void fwhost_free(void *p);

/* vw_namespace_map.csv (vwmap.rs:106-151) -> the JSON of VwNamespaceMapSource (vwmap.rs:39-51), the form the
 * reference stores in cache and regressor files. */
char *fwhost_vwmap_csv_to_json(const char *csv_text, char *err, size_t errcap);

/* ModelInstance::new_from_cmdline (model_instance.rs:296-495) for the flags of cmdline.rs that concern this path
 * (--keep --interactions --linear --ffm_field[_verbose] --ffm_k --ffm_bit_precision -b -l --power_t --ffm_learning_rate
 * --ffm_power_t --init_acc_gradient --ffm_init_acc_gradient --adaptive --sgd --noconstant --vwcompat --nn_layers --nn ...);
 * argv excludes the program name.  Returns the ModelInstance as the reference's JSON (persistence.rs:20-33). */
char *fwhost_model_instance_from_cmdline(int argc, const char *const *argv, const char *vwmap_json, char *err, size_t errcap);
char *fwhost_model_instance_normalize(const char *mi_json, char *err, size_t errcap);
/* update_hyperparameters_from_cmd (model_instance.rs:497-550) */
char *fwhost_model_instance_update_from_cmdline(const char *mi_json, int argc, const char *const *argv, char *err, size_t errcap);

/* VowpalParser (parser.rs:77-461).  parse_line: one text line (with its trailing newline, as read_until(0x0a)
 * delivers it) -> one record; returns its length in words, 0 for an empty line, -1 error, -2 "flush", -3 "hogwild_load". */
void *fwhost_parser_new(const char *vwmap_json, char *err, size_t errcap);
void fwhost_parser_free(void *parser);
int fwhost_parser_parse_line(void *parser, const char *line, size_t len, uint32_t *out, size_t cap, char *err, size_t errcap);
/* a whole text buffer, multi-threaded: records back to back in out, rec_off[n+1] word offsets; returns n */
int64_t fwhost_parser_parse_text(void *parser, const char *text, size_t len, uint32_t *out, uint64_t cap_words, uint32_t *rec_off,
                                 uint64_t cap_examples, int n_threads, uint64_t *n_words_out, char *err, size_t errcap);

/* .fwcache (cache.rs:12-26): "FWCA", u32 11, u64 + JSON vwmap, records.  A path that ends in "gz.fwcache" (the cache of a
 * *.gz input, cache.rs:68-71) holds the same bytes inside an LZ4 frame. */
int fwhost_cache_write(const char *path, const char *vwmap_json, const uint32_t *records, uint64_t n_words, char *err, size_t errcap);
int64_t fwhost_cache_read(const char *path, const char *expect_vwmap_json /* or NULL */, uint32_t **records_out, uint64_t *n_words_out,
                          uint32_t **rec_off_out, char **vwmap_json_out /* or NULL */, char *err, size_t errcap);

/* Regressor file (persistence.rs:55-97): "FWRE", u32 6, u64+JSON vwmap, u64+JSON ModelInstance, u64 total weight
 * count, then each block's payload in the byte layout fwgpu_export_block produces. */
int fwhost_regressor_write(const char *path, const char *vwmap_json, const char *mi_json, uint64_t total_weights,
                           const void *const *blocks, const uint64_t *block_bytes, uint32_t n_blocks, char *err, size_t errcap);
void *fwhost_regressor_open(const char *path, char *err, size_t errcap);
const char *fwhost_regressor_vwmap_json(void *reader);
const char *fwhost_regressor_mi_json(void *reader);
uint64_t fwhost_regressor_weights_len(void *reader);
int fwhost_regressor_read(void *reader, void *dst, uint64_t bytes);
int fwhost_regressor_skip(void *reader, uint64_t bytes);
/* what the writer stored: its optimizer (FWGPU_OPT_*; SGD = weights only, else accumulators follow, persistence.rs:163-172) and
 * whether the FFM block is quantised (ModelInstance.dequantize_weights, persistence.rs:144-161) */
uint32_t fwhost_regressor_optimizer(void *reader);
int fwhost_regressor_dequantize(void *reader);
/* quantization.rs:77-95 dequantize_ffm_weights: reads the 8-byte header and n 16-bit buckets, writes n f32 weights */
int fwhost_regressor_read_quantized(void *reader, float *dst, uint64_t n);
/* quantization.rs:19-75 quantize_ffm_weights (the writer behind --weight_quantization): dst receives the 8-byte header
 * {f32 increment, f32 min} and n 16-bit buckets, 8 + 2*n bytes; *mean_out (or NULL) the sampled mean the reference logs */
int fwhost_quantize_ffm_weights(const float *weights, uint64_t n, void *dst, float *mean_out);
/* The ModelInstance JSON a regressor file is written with: optimizer SGD for an inference regressor (main.rs:140-147,
 * persistence.rs:163-172), dequantize_weights true for a quantized one (main.rs:143-145).  Free with fwhost_free. */
char *fwhost_model_instance_for_save(const char *mi_json, int as_sgd, int quantized, char *err, size_t errcap);
void fwhost_regressor_close(void *reader);

/* ModelInstance JSON + vwmap JSON -> fwgpu_model_desc (struct fwgpu_model_desc of include/fwgpu.h, passed as void*). */
int fwhost_model_desc_from_json(const char *mi_json, const char *vwmap_json, int immutable, void *desc_out, void **keep, char *err, size_t errcap);
void fwhost_model_desc_free(void *keep);

/* The expf the device kernels use (glibc's algorithm restated; csrc/fwgpu_kernels.cuh expf_libm), on the host,
 * so that it can be compared with the C library's expf without a GPU.  Valid for |x| < 88. */
float fwhost_expf_libm(float x);
void fwhost_expf_libm_array(const float *in, float *out, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
