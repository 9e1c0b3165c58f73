/*
 * fwhost.h -- host-side (CPU, C++ inside, C ABI outside) pieces either side of the GPU hot path:
 * hashing, the VW text parser, the .fwcache reader/writer, the regressor file format and the
 * synthetic-data generator used by bench.py.  They restate the reference's host code
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

/*
 * Synthetic CTR-like data in the reference's record format (parser.rs:57-74), fixed width
 * (every namespace single-valued, value 1.0): [3+N, label, 1.0f, hash_0 .. hash_{N-1}].
 * Namespace j is named ns_names[j] (1 byte each, like benchmark/generate.py's A, B, C...); its
 * feature for example i is the string "<name><id>" with id drawn log-uniformly (Zipf ~ 1) from
 * [0, cardinality[j]) by a counter-based RNG keyed on (seed, i, j), hashed exactly as the parser
 * would hash the VW text line.  Labels are Bernoulli(sigmoid(planted additive + pairwise score)).
 * first_example lets callers generate disjoint shards / streams.  n_threads <= 0: all cores.
 * out must hold n_examples * (3 + n_namespaces) words.
 */
int fwhost_synth_records(uint32_t *out, uint64_t n_examples, uint64_t first_example, uint32_t n_namespaces,
                         const char *ns_names, const uint32_t *cardinality, uint64_t seed, int n_threads);

/* The VW text line that produces record i of the stream above (for parser round-trip tests).
 * Returns the number of bytes written (excluding the trailing NUL), or -1 if cap is too small. */
int fwhost_synth_line(char *dst, size_t cap, uint64_t example_index, uint32_t n_namespaces, const char *ns_names,
                      const uint32_t *cardinality, uint64_t seed);

/* The expf the device kernels use (glibc's algorithm restated; csrc/fwgpu_kernels.cuh expf_libm), on the host,
 * so that it can be compared with the C library's expf without a GPU.  Valid for |x| < 88. */
float fwhost_expf_libm(float x);
void fwhost_expf_libm_array(const float *in, float *out, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
