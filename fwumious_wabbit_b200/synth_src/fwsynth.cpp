// Synthetic record generator of bench.py and the tests -> libfwsynth.so.  NOT part of the product library (libfwgpu.so): the
// reference arm of bench.py generates its input without loading any product code.  Deterministic and shardable: every draw is a
// pure function of (seed, example index, namespace index), so the CPU baseline and the GPU arm of
// bench.py see the same stream without a multi-GB file.
#include "../csrc/host/murmur3.hpp"
#include <cstdint>
#include <cstddef>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

// log-uniform rank in [0, n): P(rank = r) ~ 1/(r+1)
inline uint32_t zipf_id(uint64_t h, uint32_t n)
{
    if (n <= 1) return 0;
    double u = u01(h);
    double r = std::exp(u * std::log((double)n + 1.0)) - 1.0;
    uint32_t id = (uint32_t)r;
    return id >= n ? n - 1 : id;
}

// planted score pieces: cheap hash-derived pseudo-normal values
inline double planted(uint64_t seed, uint32_t j, uint32_t id)
{
    uint64_t h = splitmix64(seed ^ (0xabcdefULL + j) * 0x100000001b3ULL ^ ((uint64_t)id << 20));
    return (u01(h) + u01(splitmix64(h)) + u01(splitmix64(h ^ 77)) - 1.5) * 0.8;
}
inline double planted_pair(uint64_t seed, uint32_t j, uint32_t a, uint32_t b)
{
    uint64_t h = splitmix64(seed * 31 + j * 1315423911ULL + ((uint64_t)a << 32 | b));
    return (u01(h) - 0.5) * 1.2;
}

struct Sample { uint32_t id; };

// bit 63 of the seed selects uniform ids instead of the log-uniform (Zipf ~ 1) law: a diagnostic stream without hot rows
inline void draw_ids(uint64_t seed, uint64_t i, uint32_t n_ns, const uint32_t *card, uint32_t *ids, uint32_t *label)
{
    double score = -0.4;
    const bool uniform = (seed >> 63) != 0;
    for (uint32_t j = 0; j < n_ns; j++) {
        uint64_t h = splitmix64(seed ^ splitmix64(i * 0x9e3779b97f4a7c15ULL + j));
        ids[j] = uniform ? (uint32_t)(h % card[j]) : zipf_id(h, card[j]);
        score += planted(seed, j, ids[j]) / std::sqrt((double)n_ns / 4.0);
    }
    for (uint32_t j = 0; j + 1 < n_ns; j += 2) score += planted_pair(seed, j, ids[j] % 64, ids[j + 1] % 64);
    double p = 1.0 / (1.0 + std::exp(-score));
    *label = u01(splitmix64(seed ^ 0x5bd1e995ULL ^ splitmix64(i + 0x1234567))) < p ? 1u : 0u;
}

inline int feature_name(char *buf, char ns, uint32_t id) { return std::snprintf(buf, 16, "%c%u", ns, id); }

} // namespace

extern "C" uint32_t fwsynth_murmur3_32(const void *key, size_t len, uint32_t seed) { return fwhost::murmur3_32(key, len, seed); }

extern "C" int fwsynth_records(uint32_t *out, uint64_t n_examples, uint64_t first_example, uint32_t n_ns, const char *ns_names,
                                    const uint32_t *card, uint64_t seed, int n_threads)
{
    if (!out || !ns_names || !card || n_ns == 0 || n_ns > 255) return -1;
    // vocabulary tables: hash of "<ns><id>" seeded with hash(ns) (parser.rs:82-83, 382-385), 31 bits.  Built once per
    // (namespace letter, cardinality) and kept for the life of the process: bench.py asks for many slices of one stream.
    static std::mutex cache_mu;
    static std::map<std::pair<char, uint32_t>, std::shared_ptr<std::vector<uint32_t>>> cache;
    std::vector<std::shared_ptr<std::vector<uint32_t>>> vocab(n_ns);
    unsigned hw = std::thread::hardware_concurrency();
    int nt = n_threads > 0 ? n_threads : (int)(hw ? hw : 1);
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        std::vector<uint32_t> todo;
        for (uint32_t j = 0; j < n_ns; j++) {
            auto it = cache.find({ns_names[j], card[j]});
            if (it != cache.end()) vocab[j] = it->second;
            else { vocab[j] = std::make_shared<std::vector<uint32_t>>(card[j]); todo.push_back(j); }
        }
        // one namespace = many ids: split every table over the threads (a few namespaces hold 1e7 ids)
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                char buf[16];
                for (uint32_t j : todo) {
                    const uint32_t ns_seed = fwhost::murmur3_32(&ns_names[j], 1, 0);
                    std::vector<uint32_t> &v = *vocab[j];
                    const uint64_t per = (card[j] + nt - 1) / nt, a = std::min<uint64_t>(card[j], per * t), b = std::min<uint64_t>(card[j], a + per);
                    for (uint64_t id = a; id < b; id++) {
                        int len = feature_name(buf, ns_names[j], (uint32_t)id);
                        v[id] = fwhost::murmur3_32(buf, (size_t)len, ns_seed) & 0x7fffffffu;
                    }
                }
            });
        for (auto &t : th) t.join();
        for (uint32_t j : todo) cache[{ns_names[j], card[j]}] = vocab[j];
    }
    const uint32_t rec_len = 3 + n_ns;
    const uint32_t one_bits = 0x3f800000u;
    auto work = [&](uint64_t a, uint64_t b) {
        std::vector<uint32_t> ids(n_ns);
        for (uint64_t e = a; e < b; e++) {
            uint32_t label;
            draw_ids(seed, first_example + e, n_ns, card, ids.data(), &label);
            uint32_t *r = out + e * rec_len;
            r[0] = rec_len; r[1] = label; r[2] = one_bits;
            for (uint32_t j = 0; j < n_ns; j++) r[3 + j] = (*vocab[j])[ids[j]];
        }
    };
    std::vector<std::thread> th;
    uint64_t per = (n_examples + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        uint64_t a = std::min<uint64_t>(n_examples, (uint64_t)t * per), b = std::min<uint64_t>(n_examples, a + per);
        if (a < b) th.emplace_back(work, a, b);
    }
    for (auto &t : th) t.join();
    return 0;
}

extern "C" int fwsynth_line(char *dst, size_t cap, uint64_t i, uint32_t n_ns, const char *ns_names, const uint32_t *card, uint64_t seed)
{
    std::vector<uint32_t> ids(n_ns);
    uint32_t label;
    draw_ids(seed, i, n_ns, card, ids.data(), &label);
    size_t pos = 0;
    int n = std::snprintf(dst, cap, "%s", label ? "1" : "-1");
    if (n < 0 || (size_t)n >= cap) return -1;
    pos = (size_t)n;
    for (uint32_t j = 0; j < n_ns; j++) {
        n = std::snprintf(dst + pos, cap - pos, " |%c %c%u", ns_names[j], ns_names[j], ids[j]);
        if (n < 0 || (size_t)n >= cap - pos) return -1;
        pos += (size_t)n;
    }
    return (int)pos;
}
