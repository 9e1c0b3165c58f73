// Host-side pieces either side of the GPU hot path (include/fwhost.h): VW text parser (parser.rs),
// .fwcache reader/writer (cache.rs), regressor file (persistence.rs), command-line -> ModelInstance
// (model_instance.rs:296-495).  CPU code, C ABI outside; nothing here calls oracle/.
#include "../../../include/fwhost.h"
#include "model.hpp"
#include "lz4frame.hpp"
#include "murmur3.hpp"

#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

using namespace fwhost;

namespace {

void set_err(char *err, size_t cap, const std::string &m) { if (err && cap) { snprintf(err, cap, "%s", m.c_str()); } }
char *dup_string(const std::string &s) { char *p = (char *)malloc(s.size() + 1); memcpy(p, s.data(), s.size()); p[s.size()] = 0; return p; }

constexpr uint32_t HEADER_LEN = 3, LABEL_OFFSET = 1, IMPORTANCE_OFFSET = 2;
constexpr uint32_t IS_NOT_SINGLE_MASK = 1u << 31, MASK31 = ~IS_NOT_SINGLE_MASK, NO_FEATURES = IS_NOT_SINGLE_MASK, NO_LABEL = 0xff, FLOAT32_ONE = 1065353216u;

// ---------------------------------------------------------------- parser (parser.rs:214-461)
struct NsInfo { uint32_t index; uint32_t seed; bool f32; };
struct Parser {
    VwMap vw;
    std::unordered_map<std::string, NsInfo> by_name; // the reference uses a radix tree (radix_tree.rs); any exact map is equivalent
    const NsInfo *by_byte[256] = {nullptr};          // the same entries for one-byte names (the common case), no hashing
    uint32_t n_ns = 0;
    const NsInfo *find(const char *name, size_t len) const
    {
        if (len == 1) return by_byte[(unsigned char)name[0]];
        const auto it = by_name.find(std::string(name, len));
        return it == by_name.end() ? nullptr : &it->second;
    }
};

bool rust_parse_f32(const char *s, size_t a, size_t b, float *out)
{
    // parse_float_or_error (parser.rs:110-139): "NONE" -> NaN, otherwise Rust's str::parse::<f32>()
    if (b - a == 4 && !memcmp(s + a, "NONE", 4)) { *out = NAN; return true; }
    if (b <= a || b - a > 63) return false;
    char tmp[64];
    memcpy(tmp, s + a, b - a);
    tmp[b - a] = 0;
    for (size_t i = 0; i < b - a; i++) {
        char c = tmp[i];
        bool ok = (c >= '0' && c <= '9') || c == '.' || c == '-' || c == '+' || c == 'e' || c == 'E' || strchr("infatyINFATY", c);
        if (!ok) return false;
    }
    char *end = nullptr;
    float v = strtof(tmp, &end);
    if (end == tmp || *end) return false;
    *out = v;
    return true;
}

// ---- VW text line -> record (behaviour of parser.rs:214-461, pinned by its known-answer tests; structure is this file's own) ----
// Three pieces: LineCursor walks the bytes of one line (the byte after the row -- the newline -- may be looked at, never
// consumed), RecordWriter owns the output encoding of parser.rs:57-74 (a namespace's single weight-1 feature lives in its
// header slot; anything else moves the namespace to (hash, f32) pairs behind the header), parse_line is the grammar
//     line      := label [importance] junk* namespace*          label decided by the FIRST byte: '1', '-', or '|' (no label)
//     namespace := '|' name [':' weight] feature*
//     feature   := name [':' weight]                             (an f32 namespace reads the value out of the name)
struct LineCursor {
    const char *p; size_t pos, end; // row = p[0, end); p[end] exists (newline)
    char peek() const { return p[pos]; }
    bool in_row() const { return pos < end; }
    void skip_spaces() { while (in_row() && p[pos] == ' ') pos++; }
    void skip_word() { while (in_row() && p[pos] != ' ') pos++; }
    void skip_until(char c) { while (in_row() && p[pos] != c) pos++; }
    // one space-delimited word starting at pos: [b, e), with `split` = its first ':' (or e when there is none)
    struct Word { size_t b, split, e; bool has_suffix() const { return split != e; } };
    Word word()
    {
        Word w;
        w.b = pos;
        while (in_row() && p[pos] != ' ' && p[pos] != ':') pos++;
        w.split = pos;
        skip_word();
        w.e = pos;
        return w;
    }
    std::string text(size_t a, size_t b) const { return std::string(p + a, b - a); }
};

class RecordWriter {
public:
    RecordWriter(uint32_t *out, size_t cap, uint32_t n_ns) : out_(out), cap_(cap), len_(n_ns + HEADER_LEN)
    {
        for (size_t i = 0; i < len_; i++) out_[i] = NO_FEATURES;
    }
    void set_label(uint32_t v) { out_[LABEL_OFFSET] = v; }
    void set_importance(float v) { memcpy(&out_[IMPORTANCE_OFFSET], &v, 4); }
    void open_namespace(uint32_t index) { slot_ = index + HEADER_LEN; count_ = 0; first_pair_ = len_; }
    // a categorical feature of weight exactly 1.0 that is the namespace's first: kept in the header slot (parser.rs:396-404)
    bool try_inline(uint32_t hash) { if (count_ != 0) return false; out_[slot_] = hash; count_ = 1; return true; }
    // any other feature: the namespace becomes a list of (hash, value bits) pairs; an inlined first feature is moved there first
    bool push_pair(uint32_t hash, float value)
    {
        if (len_ + 4 > cap_) return false;
        const uint32_t held = out_[slot_];
        if (count_ == 1 && (held & IS_NOT_SINGLE_MASK) == 0) { out_[len_++] = held; out_[len_++] = FLOAT32_ONE; }
        out_[len_++] = hash;
        memcpy(&out_[len_++], &value, 4);
        out_[slot_] = IS_NOT_SINGLE_MASK | (uint32_t)((first_pair_ << 16) + len_);
        count_++;
        return true;
    }
    int finish() { out_[0] = (uint32_t)len_; return (int)len_; }

private:
    uint32_t *out_; size_t cap_, len_, slot_ = HEADER_LEN, first_pair_ = 0; uint32_t count_ = 0;
};

// "flush" / "hogwild_load <file>" arrive on the same channel as examples (parser.rs:226-258)
int classify_command(const char *p, size_t size)
{
    if (size >= 5 && !memcmp(p, "flush", 5)) return -2;
    if (size >= strlen("hogwild_load ")) {
        LineCursor c{p, 0, size};
        size_t words = 0, first_len = 0;
        while (c.in_row()) {
            const size_t b = c.pos;
            c.skip_word();
            if (words++ == 0) first_len = c.pos - b;
            c.skip_spaces();
        }
        if (words == 2 && first_len == 12 && !memcmp(p, "hogwild_load", 12)) return -3;
    }
    return -1;
}

// returns record length in words; 0 = empty; -1 error; -2 flush; -3 hogwild_load
int parse_line(const Parser &P, const char *p, size_t size, uint32_t *out, size_t cap, std::string &err)
{
    if (size == 0) return 0;
    if (cap < P.n_ns + HEADER_LEN) { err = "record buffer too small"; return -1; }
    RecordWriter rec(out, cap, P.n_ns);
    LineCursor cur{p, 0, size - 1}; // the row ends before the line's last byte (its newline)
    const char first = p[0];
    if (first == '1') rec.set_label(1);
    else if (first == '-') rec.set_label(0);
    else if (first == '|') rec.set_label(NO_LABEL);
    else {
        const int cmd = classify_command(p, size);
        if (cmd == -1) err = "Cannot parse an example";
        return cmd;
    }
    // importance: the word after the label, unless a namespace starts there
    float importance = 1.0f;
    if (first != '|') {
        cur.skip_word();
        cur.skip_spaces();
        if (cur.peek() != '|') {
            const size_t b = cur.pos;
            cur.skip_word();
            if (!rust_parse_f32(p, b, cur.pos, &importance)) { err = "Failed parsing example importance: " + cur.text(b, cur.pos); return -1; }
            if (importance < 0.0f) { char msg[96]; snprintf(msg, sizeof(msg), "Example importance cannot be negative: %g! ", importance); err = msg; return -1; }
        }
    }
    rec.set_importance(importance);
    cur.skip_until('|'); // tags or anything else before the first namespace are ignored

    const NsInfo *ns = nullptr;
    float ns_weight = 1.0f;
    while (cur.in_row()) {
        cur.skip_spaces();
        const LineCursor::Word w = cur.word(); // may be empty when only spaces were left: the reference hashes it as a feature too
        if (p[w.b] == '|') {
            ns_weight = 1.0f;
            if (w.has_suffix() && !rust_parse_f32(p, w.split + 1, w.e, &ns_weight)) { err = "Failed parsing namespace weight: " + cur.text(w.split + 1, w.e); return -1; }
            ns = P.find(p + w.b + 1, w.split - (w.b + 1));
            if (!ns) { err = "Feature name was not predeclared in vw_namespace_map.csv: " + cur.text(w.b + 1, w.split); return -1; }
            rec.open_namespace(ns->index);
        } else {
            const uint32_t hash = murmur3_32(p + w.b, w.split - w.b, ns ? ns->seed : 0) & MASK31;
            float weight = 1.0f;
            if (w.has_suffix() && !rust_parse_f32(p, w.split + 1, w.e, &weight)) { err = "Failed parsing feature weight: " + cur.text(w.split + 1, w.e); return -1; }
            const bool is_f32 = ns && ns->f32;
            if (!(!is_f32 && ns_weight == 1.0f && weight == 1.0f && rec.try_inline(hash))) {
                float value = ns_weight * weight;
                if (is_f32) { // the number is the feature's name after the skipped prefix; an empty name is NaN (parser.rs:416-433)
                    const size_t vb = w.b + P.vw.namespace_skip_prefix;
                    value = NAN;
                    if (vb != w.split && (vb > w.split || !rust_parse_f32(p, vb, w.split, &value))) { err = "Failed parsing feature value to float (for float namespace): " + cur.text(w.b, w.split); return -1; }
                }
                if (!rec.push_pair(hash, value)) { err = "record too long"; return -1; }
                if (is_f32 && ns_weight * weight != 1.0f) { err = "Namespaces that are f32 can not have weight attached neither to namespace nor to a single feature (basically they can' use :weight syntax"; return -1; }
            }
        }
        cur.pos = w.e + 1;
    }
    return rec.finish();
}

// ---------------------------------------------------------------- files
constexpr uint32_t CACHE_VERSION = 11, REGRESSOR_VERSION = 6; // cache.rs:12-13, persistence.rs:17-18

struct RegReader { FILE *f = nullptr; std::string vwmap_json, mi_json; uint64_t weights_len = 0; uint32_t optimizer = 0; int dequantize = 0; };

bool read_exact(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n; }
bool read_blob(FILE *f, std::string &out)
{
    uint64_t len = 0;
    if (!read_exact(f, &len, 8) || len > (1ull << 32)) return false;
    out.resize(len);
    return len == 0 || read_exact(f, &out[0], len);
}

// ---------------------------------------------------------------- cmdline (cmdline.rs, model_instance.rs:296-495)
struct Args {
    std::map<std::string, std::vector<std::string>> multi;
    bool has(const std::string &k) const { return multi.count(k) > 0; }
    const std::string *one(const std::string &k) const { auto it = multi.find(k); return it == multi.end() || it->second.empty() ? nullptr : &it->second.back(); }
};
// flags that take a value / flags that do not (subset of cmdline.rs:9-322)
const char *VALUE_FLAGS[] = {"data", "predictions", "final_regressor", "initial_regressor", "keep", "interactions", "linear", "ffm_field", "ffm_field_verbose",
                             "ffm_k", "ffm_bit_precision", "bit_precision", "learning_rate", "ffm_learning_rate", "nn_learning_rate", "power_t", "ffm_power_t", "nn_power_t",
                             "init_acc_gradient", "ffm_init_acc_gradient", "nn_init_acc_gradient", "ffm_init_center", "ffm_init_width", "ffm_init_zero_band",
                             "ffm_initialization_type", "minimum_learning_rate", "link", "loss_function", "l2", "hash", "nn_layers", "nn_topology", "nn",
                             "predictions_after", "holdout_after", "hogwild_threads", "convert_inference_regressor", "transform", "prediction_model_delay",
                             "batch_size", "device", nullptr};
const char *BOOL_FLAGS[] = {"cache", "testonly", "save_resume", "adaptive", "sgd", "noconstant", "vwcompat", "hogwild_training", "quiet", "predictions_stdout",
                            "build_cache_without_training", "sequential", "invariant", "normalized", "weight_quantization", nullptr};
const std::pair<const char *, const char *> SHORT_FLAGS[] = {{"d", "data"}, {"p", "predictions"}, {"f", "final_regressor"}, {"i", "initial_regressor"}, {"b", "bit_precision"},
                                                             {"l", "learning_rate"}, {"c", "cache"}, {"t", "testonly"}, {"q", "interactions"}};

bool in_list(const char **l, const std::string &s) { for (; *l; l++) if (s == *l) return true; return false; }

Args parse_args(int argc, const char *const *argv)
{
    Args a;
    for (int i = 0; i < argc; i++) {
        std::string t = argv[i], name;
        if (t.rfind("--", 0) == 0) name = t.substr(2);
        else if (t.size() == 2 && t[0] == '-') { for (auto &sf : SHORT_FLAGS) if (t[1] == sf.first[0]) name = sf.second; if (name.empty()) throw std::runtime_error("Found argument '" + t + "' which wasn't expected"); }
        else throw std::runtime_error("Found argument '" + t + "' which wasn't expected, or isn't valid in this context");
        std::string inline_val;
        size_t eq = name.find('=');
        if (eq != std::string::npos) { inline_val = name.substr(eq + 1); name = name.substr(0, eq); }
        if (in_list(BOOL_FLAGS, name)) a.multi[name];
        else if (in_list(VALUE_FLAGS, name)) {
            if (eq != std::string::npos) a.multi[name].push_back(inline_val);
            else { if (i + 1 >= argc) throw std::runtime_error("The argument '--" + name + "' requires a value but none was supplied"); a.multi[name].push_back(argv[++i]); }
        } else throw std::runtime_error("Found argument '--" + name + "' which wasn't expected, or isn't valid in this context");
    }
    return a;
}

NsDesc ns_by_char(const VwMap &vw, char c)
{
    const VwEntry *e = vw.by_vwname(std::string(1, c));
    if (!e) throw std::runtime_error(std::string("Unknown namespace char in command line: ") + c);
    return NsDesc{e->index, e->f32};
}
NsDesc ns_by_verbose(const VwMap &vw, const std::string &s)
{
    const VwEntry *e = vw.by_verbose(s);
    if (!e) throw std::runtime_error("Unknown verbose namespace in command line: " + s);
    return NsDesc{e->index, e->f32};
}
float parse_float(const Args &a, const char *k, float dflt) { const std::string *v = a.one(k); return v ? strtof(v->c_str(), nullptr) : dflt; }

ModelInstanceH mi_from_args(const Args &a, const VwMap &vw)
{
    ModelInstanceH mi;
    const bool vwcompat = a.has("vwcompat");
    if (vwcompat) {
        mi.fastmath = false;
        mi.init_acc_gradient = 0.0f;
        if (!a.has("keep")) throw std::runtime_error("--vwcompat requires at least one --keep parameter, we do not implicitly take all features available");
        const std::string *h = a.one("hash");
        if (!h || *h != "all") throw std::runtime_error("--vwcompat requires use of --hash all");
        if (!a.has("sgd")) throw std::runtime_error("--vwcompat requires use of --sgd");
    }
    if (a.has("transform")) throw std::runtime_error("--transform namespaces are out of scope for the GPU path (feature_transform_*.rs)");
    auto combo_from_chars = [&](const std::string &s) {
        ComboDesc c;
        std::string names = s;
        size_t colon = s.find(':');
        if (colon != std::string::npos) {
            if (s.find(':', colon + 1) != std::string::npos) throw std::runtime_error("only one value parameter allowed (denoted with \":\"): \"" + s + "\"");
            c.weight = strtof(s.substr(colon + 1).c_str(), nullptr);
            names = s.substr(0, colon);
        }
        for (char ch : names) c.ns.push_back(ns_by_char(vw, ch));
        return c;
    };
    if (a.has("keep")) for (auto &s : a.multi.at("keep")) mi.feature_combo_descs.push_back(combo_from_chars(s));
    if (a.has("interactions")) for (auto &s : a.multi.at("interactions")) mi.feature_combo_descs.push_back(combo_from_chars(s));
    if (a.has("linear")) for (auto &s : a.multi.at("linear")) {
        ComboDesc c;
        std::string names = s;
        size_t colon = s.find(':');
        if (colon != std::string::npos) {
            if (s.find(':', colon + 1) != std::string::npos) throw std::runtime_error("Verbose features cannot have \":\" as part of their names: \"" + s + "\"");
            c.weight = strtof(s.substr(colon + 1).c_str(), nullptr);
            names = s.substr(0, colon);
        }
        size_t p0 = 0;
        for (;;) { size_t cm = names.find(',', p0); c.ns.push_back(ns_by_verbose(vw, names.substr(p0, cm == std::string::npos ? std::string::npos : cm - p0))); if (cm == std::string::npos) break; p0 = cm + 1; }
        mi.feature_combo_descs.push_back(c);
    }
    if (const std::string *v = a.one("ffm_k")) { mi.ffm_k = (uint32_t)strtoul(v->c_str(), nullptr, 10); if (mi.ffm_k > 128) throw std::runtime_error("Maximum ffm_k is: 128, passed: " + *v); }
    if (const std::string *v = a.one("ffm_initialization_type")) mi.ffm_initialization_type = *v;
    mi.ffm_init_center = parse_float(a, "ffm_init_center", mi.ffm_init_center);
    mi.ffm_init_width = parse_float(a, "ffm_init_width", mi.ffm_init_width);
    mi.ffm_init_zero_band = parse_float(a, "ffm_init_zero_band", mi.ffm_init_zero_band);
    if (a.has("ffm_field")) for (auto &s : a.multi.at("ffm_field")) { std::vector<NsDesc> f; for (char ch : s) f.push_back(ns_by_char(vw, ch)); mi.ffm_fields.push_back(f); }
    if (a.has("ffm_field_verbose")) for (auto &s : a.multi.at("ffm_field_verbose")) {
        if (s.find(':') != std::string::npos) throw std::runtime_error("Fields currently do not support passing a value via : \"" + s + "\"");
        std::vector<NsDesc> f;
        size_t p0 = 0;
        for (;;) { size_t cm = s.find(',', p0); f.push_back(ns_by_verbose(vw, s.substr(p0, cm == std::string::npos ? std::string::npos : cm - p0))); if (cm == std::string::npos) break; p0 = cm + 1; }
        mi.ffm_fields.push_back(f);
    }
    if (const std::string *v = a.one("ffm_bit_precision")) mi.ffm_bit_precision = (uint32_t)strtoul(v->c_str(), nullptr, 10);
    if (const std::string *v = a.one("bit_precision")) mi.bit_precision = (uint32_t)strtoul(v->c_str(), nullptr, 10);
    mi.learning_rate = parse_float(a, "learning_rate", mi.learning_rate);
    mi.init_acc_gradient = parse_float(a, "init_acc_gradient", mi.init_acc_gradient);
    mi.power_t = parse_float(a, "power_t", mi.power_t);
    mi.ffm_learning_rate = parse_float(a, "ffm_learning_rate", mi.learning_rate);         // defaults chain like model_instance.rs:418-428
    mi.ffm_init_acc_gradient = parse_float(a, "ffm_init_acc_gradient", mi.init_acc_gradient);
    mi.ffm_power_t = parse_float(a, "ffm_power_t", mi.power_t);
    mi.nn_learning_rate = parse_float(a, "nn_learning_rate", mi.ffm_learning_rate);
    mi.nn_init_acc_gradient = parse_float(a, "nn_init_acc_gradient", mi.ffm_init_acc_gradient);
    mi.nn_power_t = parse_float(a, "nn_power_t", mi.ffm_power_t);
    if (const std::string *v = a.one("nn_layers")) mi.nn_layers.resize(strtoul(v->c_str(), nullptr, 10));
    if (const std::string *v = a.one("nn_topology")) mi.nn_topology = *v;
    if (a.has("nn")) for (auto &s : a.multi.at("nn")) {
        size_t c1 = s.find(':'), c2 = c1 == std::string::npos ? c1 : s.find(':', c1 + 1);
        if (c1 == std::string::npos || c2 == std::string::npos || s.find(':', c2 + 1) != std::string::npos) throw std::runtime_error("--nn parameters have to be of form layer:parameter_name:parameter_value: " + s);
        size_t layer = strtoul(s.substr(0, c1).c_str(), nullptr, 10);
        if (layer >= mi.nn_layers.size()) throw std::runtime_error("--nn parameter addressing layer " + std::to_string(layer) + ", but we have only " + std::to_string(mi.nn_layers.size()) + " layers");
        mi.nn_layers[layer].emplace_back(s.substr(c1 + 1, c2 - c1 - 1), s.substr(c2 + 1));
    }
    if (const std::string *v = a.one("minimum_learning_rate")) mi.minimum_learning_rate = strtof(v->c_str(), nullptr);
    if (const std::string *v = a.one("link")) if (*v != "logistic") throw std::runtime_error("--link only supports 'logistic'");
    if (const std::string *v = a.one("loss_function")) if (*v != "logistic") throw std::runtime_error("--loss_function only supports 'logistic'");
    if (const std::string *v = a.one("l2")) if (std::fabs(strtof(v->c_str(), nullptr)) > 0.00000001f) throw std::runtime_error("--l2 can only be 0.0");
    if (a.has("noconstant")) mi.add_constant_feature = false;
    if (a.has("sgd")) mi.optimizer = FWGPU_OPT_SGD;
    if (a.has("adaptive")) mi.optimizer = FWGPU_OPT_ADAGRAD_FLEX;
    if (mi.optimizer == FWGPU_OPT_ADAGRAD_FLEX && mi.fastmath) mi.optimizer = FWGPU_OPT_ADAGRAD_LUT;
    return mi;
}

} // namespace

// ================================================================ C ABI
extern "C" uint32_t fwhost_murmur3_32(const void *key, size_t len, uint32_t seed) { return fwhost::murmur3_32(key, len, seed); }

extern "C" {

void fwhost_free(void *p) { free(p); }

char *fwhost_vwmap_csv_to_json(const char *csv, char *err, size_t errcap)
{
    try { return dup_string(json_to_string(vwmap_to_json(vwmap_from_csv(csv)))); }
    catch (const std::exception &e) { set_err(err, errcap, e.what()); return nullptr; }
}

char *fwhost_model_instance_from_cmdline(int argc, const char *const *argv, const char *vwmap_json, char *err, size_t errcap)
{
    try {
        VwMap vw = vwmap_from_json(json_parse(vwmap_json));
        return dup_string(json_to_string(mi_to_json(mi_from_args(parse_args(argc, argv), vw))));
    } catch (const std::exception &e) { set_err(err, errcap, e.what()); return nullptr; }
}

char *fwhost_model_instance_normalize(const char *mi_json, char *err, size_t errcap)
{
    try { return dup_string(json_to_string(mi_to_json(mi_from_json(json_parse(mi_json))))); }
    catch (const std::exception &e) { set_err(err, errcap, e.what()); return nullptr; }
}

// update_hyperparameters_from_cmd (model_instance.rs:497-550): -l / --ffm_learning_rate / --power_t / --ffm_power_t override a loaded model
char *fwhost_model_instance_update_from_cmdline(const char *mi_json, int argc, const char *const *argv, char *err, size_t errcap)
{
    try {
        ModelInstanceH mi = mi_from_json(json_parse(mi_json));
        Args a = parse_args(argc, argv);
        if (const std::string *v = a.one("learning_rate")) mi.learning_rate = strtof(v->c_str(), nullptr);
        if (const std::string *v = a.one("ffm_learning_rate")) mi.ffm_learning_rate = strtof(v->c_str(), nullptr);
        if (const std::string *v = a.one("power_t")) mi.power_t = strtof(v->c_str(), nullptr);
        if (const std::string *v = a.one("ffm_power_t")) mi.ffm_power_t = strtof(v->c_str(), nullptr);
        return dup_string(json_to_string(mi_to_json(mi)));
    } catch (const std::exception &e) { set_err(err, errcap, e.what()); return nullptr; }
}

// ---- parser
void *fwhost_parser_new(const char *vwmap_json, char *err, size_t errcap)
{
    try {
        Parser *P = new Parser();
        P->vw = vwmap_from_json(json_parse(vwmap_json));
        P->n_ns = P->vw.num_namespaces;
        for (auto &e : P->vw.entries) P->by_name[e.vwname] = NsInfo{e.index, murmur3_32(e.vwname.data(), e.vwname.size(), 0), e.f32}; // parser.rs:82-83
        for (auto &kv : P->by_name) if (kv.first.size() == 1) P->by_byte[(unsigned char)kv.first[0]] = &kv.second; // nodes of an unordered_map do not move
        return P;
    } catch (const std::exception &e) { set_err(err, errcap, e.what()); return nullptr; }
}
void fwhost_parser_free(void *p) { delete (Parser *)p; }

int fwhost_parser_parse_line(void *parser, const char *line, size_t len, uint32_t *out, size_t cap, char *err, size_t errcap)
{
    std::string e;
    int n = parse_line(*(Parser *)parser, line, len, out, cap, e);
    if (n == -1) set_err(err, errcap, e);
    return n;
}

// Whole buffer -> records back to back + offsets.  Lines are split first, then parsed by n_threads workers into
// per-thread slabs that are concatenated in order (the reference parses on one thread, main.rs:213-239).
int64_t fwhost_parser_parse_text(void *parser, const char *text, size_t len, uint32_t *out, uint64_t cap_words, uint32_t *rec_off, uint64_t cap_examples,
                                 int n_threads, uint64_t *n_words_out, char *err, size_t errcap)
{
    const Parser &P = *(Parser *)parser;
    std::vector<std::pair<size_t, size_t>> lines;
    for (size_t pos = 0; pos < len;) {
        const char *nl = (const char *)memchr(text + pos, '\n', len - pos);
        size_t end = nl ? (size_t)(nl - text) + 1 : len;
        lines.emplace_back(pos, end - pos);
        pos = end;
    }
    if (lines.size() > cap_examples) { set_err(err, errcap, "rec_off capacity too small"); return -1; }
    unsigned hw = std::thread::hardware_concurrency();
    int nt = n_threads > 0 ? n_threads : (int)(hw ? hw : 1);
    nt = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, lines.size() / 1024));
    std::vector<std::vector<uint32_t>> slabs(nt), lens(nt);
    std::vector<std::string> errs(nt);
    std::vector<int64_t> bad(nt, -1);
    size_t per = (lines.size() + nt - 1) / nt;
    auto work = [&](int t) {
        size_t a = std::min(lines.size(), (size_t)t * per), b = std::min(lines.size(), a + per);
        std::vector<uint32_t> tmp(1 << 16);
        if (b > a) { // one allocation per slab in the common case: a record word for every 6 bytes of text is typical of VW lines
            slabs[t].reserve((lines[b - 1].first + lines[b - 1].second - lines[a].first) / 6 + 64);
            lens[t].reserve(b - a);
        }
        for (size_t i = a; i < b; i++) {
            const char *lp = text + lines[i].first;
            size_t ll = lines[i].second;
            std::string e;
            // a last line without '\n' is parsed as the reference would see it after read_until(): pad a newline
            std::string padded;
            if (lp[ll - 1] != '\n') { padded.assign(lp, ll); padded.push_back('\n'); lp = padded.data(); ll = padded.size(); }
            int n = parse_line(P, lp, ll, tmp.data(), tmp.size(), e);
            if (n < 0) { bad[t] = (int64_t)i; errs[t] = n == -2 ? "flush command inside a data file" : n == -3 ? "hogwild_load command inside a data file" : e; return; }
            if (n == 0) continue;
            slabs[t].insert(slabs[t].end(), tmp.begin(), tmp.begin() + n);
            lens[t].push_back((uint32_t)n);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto &t : th) t.join();
    for (int t = 0; t < nt; t++) if (bad[t] >= 0) { set_err(err, errcap, errs[t] + " (line " + std::to_string(bad[t] + 1) + ")"); return -1; }
    uint64_t words = 0, n = 0;
    {   // rec_off holds u32 word offsets
        uint64_t total = 0;
        for (int t = 0; t < nt; t++) total += slabs[t].size();
        if (total > 0xffffffffull) { set_err(err, errcap, "input exceeds 2^32 record words (16 GiB): split it"); return -1; }
    }
    // every worker's slab goes to its place in `out`, and its record offsets to rec_off, again one thread per slab
    std::vector<uint64_t> word_base(nt), rec_base(nt);
    for (int t = 0; t < nt; t++) { word_base[t] = words; rec_base[t] = n; words += slabs[t].size(); n += lens[t].size(); }
    if (words > cap_words) { set_err(err, errcap, "record buffer too small"); return -1; }
    auto place = [&](int t) {
        if (!slabs[t].empty()) memcpy(out + word_base[t], slabs[t].data(), slabs[t].size() * 4);
        uint64_t w = word_base[t], i = rec_base[t];
        for (uint32_t l : lens[t]) { rec_off[i++] = (uint32_t)w; w += l; }
    };
    if (nt == 1) place(0);
    else {
        std::vector<std::thread> movers;
        for (int t = 0; t < nt; t++) movers.emplace_back(place, t);
        for (auto &m : movers) m.join();
    }
    rec_off[n] = (uint32_t)words;
    if (n_words_out) *n_words_out = words;
    return (int64_t)n;
}

// ---- .fwcache (cache.rs:12-26, 133-161, 187-232).  The cache of an input whose name ends in "gz" is the same byte stream
// inside an LZ4 frame (cache.rs:71, 89-125); lz4frame.hpp is the codec.
static bool cache_is_compressed(const std::string &path)
{
    const std::string suffix = "gz.fwcache"; // final_filename = input_filename + ".fwcache", gz = input_filename.ends_with("gz")
    return path.size() >= suffix.size() && path.compare(path.size() - suffix.size(), suffix.size(), suffix) == 0;
}

int fwhost_cache_write(const char *path, const char *vwmap_json, const uint32_t *records, uint64_t n_words, char *err, size_t errcap)
{
    try {
        std::string blob = json_to_string(vwmap_to_json(vwmap_from_json(json_parse(vwmap_json))));
        std::string tmp = std::string(path) + ".writing"; // cache.rs:69-70, 147-153: write then rename
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot create " + tmp);
        uint64_t len = blob.size();
        bool ok;
        if (cache_is_compressed(path)) {
            std::vector<uint8_t> image;
            image.reserve(16 + len + n_words * 4);
            auto put = [&](const void *p_, size_t n_) { image.insert(image.end(), (const uint8_t *)p_, (const uint8_t *)p_ + n_); };
            put("FWCA", 4); put(&CACHE_VERSION, 4); put(&len, 8); put(blob.data(), len); put(records, n_words * 4);
            const std::vector<uint8_t> z = lz4::encode_frame(image.data(), image.size());
            ok = fwrite(z.data(), 1, z.size(), f) == z.size();
        } else {
            ok = fwrite("FWCA", 1, 4, f) == 4 && fwrite(&CACHE_VERSION, 4, 1, f) == 1 && fwrite(&len, 8, 1, f) == 1 && fwrite(blob.data(), 1, len, f) == len &&
                 (n_words == 0 || fwrite(records, 4, n_words, f) == n_words);
        }
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw std::runtime_error("write failed");
        if (rename(tmp.c_str(), path)) throw std::runtime_error("rename failed");
        return 0;
    } catch (const std::exception &e) { set_err(err, errcap, e.what()); return -1; }
}

int64_t fwhost_cache_read(const char *path, const char *expect_vwmap_json, uint32_t **records_out, uint64_t *n_words_out, uint32_t **rec_off_out,
                          char **vwmap_json_out, char *err, size_t errcap)
{
    FILE *f = nullptr;
    try {
        f = fopen(path, "rb");
        if (!f) throw std::runtime_error(std::string("cannot open ") + path);
        // a plain cache is read straight into the buffer the caller gets; a compressed one is decoded first (cache.rs:89-125)
        const bool compressed = cache_is_compressed(path);
        std::vector<uint8_t> image;
        size_t pos = 0;
        if (compressed) {
            fseek(f, 0, SEEK_END);
            const long sz = ftell(f);
            fseek(f, 0, SEEK_SET);
            std::vector<uint8_t> packed(sz > 0 ? (size_t)sz : 0);
            if (!packed.empty() && !read_exact(f, packed.data(), packed.size())) throw std::runtime_error("short read");
            image = lz4::decode_frames(packed.data(), packed.size());
        }
        auto take = [&](void *dst, size_t n_) {
            if (!compressed) return read_exact(f, dst, n_);
            if (image.size() - pos < n_) return false;
            memcpy(dst, image.data() + pos, n_);
            pos += n_;
            return true;
        };
        char magic[4];
        uint32_t version = 0;
        if (!take(magic, 4) || memcmp(magic, "FWCA", 4)) throw std::runtime_error("Cache header does not begin with magic bytes FWFW"); // sic, cache.rs:167
        if (!take(&version, 4) || version != CACHE_VERSION) throw std::runtime_error("Cache file version of this binary: 11, version of the cache file: " + std::to_string(version));
        uint64_t blen = 0;
        if (!take(&blen, 8) || blen > (1ull << 32)) throw std::runtime_error("truncated cache header");
        std::string blob(blen, '\0');
        if (blen && !take(&blob[0], blen)) throw std::runtime_error("truncated cache header");
        VwMap in_file = vwmap_from_json(json_parse(blob));
        if (expect_vwmap_json && !(in_file == vwmap_from_json(json_parse(expect_vwmap_json)))) throw std::runtime_error("vw_namespace_map.csv and the one from cache file differ");
        uint64_t body_bytes;
        if (compressed) body_bytes = image.size() - pos;
        else {
            const long here = ftell(f);
            fseek(f, 0, SEEK_END);
            body_bytes = (uint64_t)(ftell(f) - here);
            fseek(f, here, SEEK_SET);
        }
        const uint64_t n_words = body_bytes / 4;
        if (n_words > 0xffffffffull) throw std::runtime_error("cache exceeds 2^32 record words (16 GiB): rec_off holds u32 word offsets");
        uint32_t *recs = (uint32_t *)malloc(std::max<uint64_t>(n_words, 1) * 4);
        if (!recs) throw std::runtime_error("out of memory reading the cache");
        if (n_words && compressed && !take(recs, n_words * 4)) { free(recs); throw std::runtime_error("short read"); }
        if (n_words && !compressed) { // straight from the file into place; large bodies by several readers at once (page faults and copies overlap)
            const uint64_t bytes = n_words * 4, base = (uint64_t)ftell(f);
            const int fd = fileno(f);
            const unsigned hw = std::thread::hardware_concurrency();
            const uint64_t readers = std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(hw ? hw : 1, 16), bytes >> 24)); // >= 16 MiB each
            const uint64_t per = ((bytes + readers - 1) / readers + 4095) & ~uint64_t(4095);
            std::vector<int> failed(readers, 0);
            auto pull = [&](uint64_t t) {
                uint64_t a = std::min(bytes, t * per);
                const uint64_t b = std::min(bytes, a + per);
                while (a < b) {
                    const ssize_t r = pread(fd, (char *)recs + a, (size_t)std::min<uint64_t>(b - a, 64u << 20), (off_t)(base + a));
                    if (r <= 0) { failed[t] = 1; return; }
                    a += (uint64_t)r;
                }
            };
            if (readers == 1) pull(0);
            else {
                std::vector<std::thread> th;
                for (uint64_t t = 0; t < readers; t++) th.emplace_back(pull, t);
                for (auto &x : th) x.join();
            }
            for (int bad : failed) if (bad) { free(recs); throw std::runtime_error("short read"); }
        }
        fclose(f);
        f = nullptr;
        { std::vector<uint8_t>().swap(image); }
        std::vector<uint32_t> offs;
        if (n_words) offs.reserve(n_words / std::max<uint32_t>(recs[0], HEADER_LEN) + 16); // exact when every record has the first one's length
        uint64_t w = 0;
        while (w < n_words) {
            uint32_t l = recs[w];
            if (l < HEADER_LEN || w + l > n_words) { free(recs); throw std::runtime_error("corrupt record length in cache"); }
            offs.push_back((uint32_t)w);
            w += l;
        }
        offs.push_back((uint32_t)w);
        uint32_t *ro = (uint32_t *)malloc(offs.size() * 4);
        memcpy(ro, offs.data(), offs.size() * 4);
        *records_out = recs; *n_words_out = n_words; *rec_off_out = ro;
        if (vwmap_json_out) *vwmap_json_out = dup_string(blob);
        return (int64_t)offs.size() - 1;
    } catch (const std::exception &e) {
        if (f) fclose(f);
        set_err(err, errcap, e.what());
        return -1;
    }
}

// ---- regressor file (persistence.rs:55-97; regressor.rs:426-442)
int fwhost_regressor_write(const char *path, const char *vwmap_json, const char *mi_json, uint64_t total_weights, const void *const *blocks,
                           const uint64_t *block_bytes, uint32_t n_blocks, char *err, size_t errcap)
{
    try {
        std::string vb = json_to_string(vwmap_to_json(vwmap_from_json(json_parse(vwmap_json))));
        std::string mb = json_to_string(mi_to_json(mi_from_json(json_parse(mi_json))));
        FILE *f = fopen(path, "wb");
        if (!f) throw std::runtime_error(std::string("Cannot open ") + path + " to save regressor to");
        uint64_t l1 = vb.size(), l2 = mb.size();
        bool ok = fwrite("FWRE", 1, 4, f) == 4 && fwrite(&REGRESSOR_VERSION, 4, 1, f) == 1 && fwrite(&l1, 8, 1, f) == 1 && fwrite(vb.data(), 1, l1, f) == l1 &&
                  fwrite(&l2, 8, 1, f) == 1 && fwrite(mb.data(), 1, l2, f) == l2 && fwrite(&total_weights, 8, 1, f) == 1;
        for (uint32_t i = 0; ok && i < n_blocks; i++) ok = block_bytes[i] == 0 || fwrite(blocks[i], 1, block_bytes[i], f) == block_bytes[i];
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw std::runtime_error("write failed");
        return 0;
    } catch (const std::exception &e) { set_err(err, errcap, e.what()); return -1; }
}

void *fwhost_regressor_open(const char *path, char *err, size_t errcap)
{
    RegReader *r = new RegReader();
    try {
        r->f = fopen(path, "rb");
        if (!r->f) throw std::runtime_error(std::string("cannot open ") + path);
        char magic[4];
        uint32_t version = 0;
        if (!read_exact(r->f, magic, 4) || memcmp(magic, "FWRE", 4)) throw std::runtime_error("Regressor header error: does not begin with magic bytes FWRE");
        if (!read_exact(r->f, &version, 4) || version != REGRESSOR_VERSION) throw std::runtime_error("Regressor file version of this binary: 6, version of the regressor file: " + std::to_string(version));
        if (!read_blob(r->f, r->vwmap_json) || !read_blob(r->f, r->mi_json) || !read_exact(r->f, &r->weights_len, 8)) throw std::runtime_error("truncated regressor header");
        json_parse(r->vwmap_json);
        {   // what the writer stored: the optimizer decides whether accumulators follow the weights, dequantize_weights
            // whether the FFM block is the 16-bit form of quantization.rs:41-95
            ModelInstanceH m = mi_from_json(json_parse(r->mi_json));
            r->optimizer = m.optimizer;
            r->dequantize = m.dequantize_weights > 0 ? 1 : 0;
        }
        return r;
    } catch (const std::exception &e) {
        if (r->f) fclose(r->f);
        delete r;
        set_err(err, errcap, e.what());
        return nullptr;
    }
}
const char *fwhost_regressor_vwmap_json(void *r) { return ((RegReader *)r)->vwmap_json.c_str(); }
const char *fwhost_regressor_mi_json(void *r) { return ((RegReader *)r)->mi_json.c_str(); }
uint64_t fwhost_regressor_weights_len(void *r) { return ((RegReader *)r)->weights_len; }
int fwhost_regressor_read(void *r, void *dst, uint64_t bytes) { return read_exact(((RegReader *)r)->f, dst, bytes) ? 0 : -1; }
uint32_t fwhost_regressor_optimizer(void *r) { return ((RegReader *)r)->optimizer; }
int fwhost_regressor_dequantize(void *r) { return ((RegReader *)r)->dequantize; }
// quantization.rs:77-95 dequantize_ffm_weights: 8-byte header {f32 increment, f32 min}, then one IEEE half per weight holding the
// bucket number; weight = min + bucket * increment
int fwhost_regressor_read_quantized(void *r, float *dst, uint64_t n)
{
    FILE *f = ((RegReader *)r)->f;
    float hdr[2];
    if (!read_exact(f, hdr, 8)) return -1;
    std::vector<uint16_t> buf(1 << 16);
    for (uint64_t done = 0; done < n;) {
        const uint64_t cnt = std::min<uint64_t>(buf.size(), n - done);
        if (!read_exact(f, buf.data(), cnt * 2)) return -1;
        for (uint64_t i = 0; i < cnt; i++) {
            const uint16_t h = buf[i];
            const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, ex = (h >> 10) & 0x1fu, man = h & 0x3ffu;
            uint32_t bits;
            if (ex == 0) {
                const float sub = (float)man * 5.9604644775390625e-08f; // man * 2^-24, exact
                memcpy(&bits, &sub, 4);
                bits |= sign;
            } else if (ex == 31) bits = sign | 0x7f800000u | (man << 13);
            else bits = sign | ((ex + 112u) << 23) | (man << 13);
            float v;
            memcpy(&v, &bits, 4);
            dst[done + i] = hdr[1] + v * hdr[0];
        }
        done += cnt;
    }
    return 0;
}
// quantization.rs:19-75 quantize_ffm_weights, the writer of the 16-bit form: min and max of the block rounded to 4 decimals, 65 025
// equal buckets between them, each weight's bucket number (round half away from zero) stored as an IEEE half (round to nearest even,
// what half::f16::from_f32 does).  dst receives 8 + 2·n bytes.  *mean_out is the reference's sampled mean (every 10th weight), the
// figure its "exploded weights" warning looks at.
static uint16_t f32_to_half_rne(float v)
{
    uint32_t x;
    memcpy(&x, &v, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u | ((x >> 13) & 0x3ffu) : 0u));
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                 // rounds to 65520 or more: infinity
    if (x < 0x33000001u) return (uint16_t)sign;                               // at most 2^-25: zero (the tie goes to even)
    uint32_t ex = x >> 23, man = x & 0x7fffffu, h;
    if (ex < 113) {                                                           // half subnormal: value = m · 2^-24
        man |= 0x800000u;
        const uint32_t shift = 126 - ex;                                      // 14 … 24
        h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1))) h++;
    } else {
        h = ((ex - 112) << 10) | (man >> 13);
        const uint32_t rem = man & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;                // a carry into the exponent is the right answer
    }
    return (uint16_t)(sign | h);
}
int fwhost_quantize_ffm_weights(const float *w, uint64_t n, void *dst, float *mean_out)
{
    if (!w || !dst || n == 0) return -1;
    float lo = w[0], hi = w[0], mean = 0.0f;
    uint64_t sampled = 0;
    for (uint64_t i = 0; i < n; i++) {
        hi = fmaxf(hi, w[i]);
        lo = fminf(lo, w[i]);
        if (i % 10 == 0) { sampled++; mean += w[i]; }
    }
    lo = roundf(lo * 10000.0f) / 10000.0f;
    hi = roundf(hi * 10000.0f) / 10000.0f;
    const float increment = (hi - lo) / 65025.0f;
    if (mean_out) *mean_out = mean / (float)sampled;
    uint8_t *out = (uint8_t *)dst;
    memcpy(out, &increment, 4);
    memcpy(out + 4, &lo, 4);
    uint16_t *q = (uint16_t *)(out + 8);
    for (uint64_t i = 0; i < n; i++) q[i] = f32_to_half_rne(roundf((w[i] - lo) / increment));
    return 0;
}
// The ModelInstance a regressor file is written with: an inference regressor says optimizer SGD (main.rs:140-147,
// persistence.rs:163-172), a quantized one dequantize_weights = true (main.rs:143-145).  Caller frees with fwhost_free.
char *fwhost_model_instance_for_save(const char *mi_json, int as_sgd, int quantized, char *err, size_t errcap)
{
    try {
        ModelInstanceH m = mi_from_json(json_parse(mi_json));
        if (as_sgd) m.optimizer = FWGPU_OPT_SGD;
        if (quantized) m.dequantize_weights = 1;
        return dup_string(json_to_string(mi_to_json(m)));
    } catch (const std::exception &e) { set_err(err, errcap, e.what()); return nullptr; }
}
int fwhost_regressor_skip(void *r, uint64_t bytes) { return fseek(((RegReader *)r)->f, (long)bytes, SEEK_CUR) == 0 ? 0 : -1; }
void fwhost_regressor_close(void *r) { RegReader *rr = (RegReader *)r; if (rr) { if (rr->f) fclose(rr->f); delete rr; } }

// ModelInstance + vwmap JSON -> fwgpu_model_desc.  keep must be released with fwhost_model_desc_free.
int fwhost_model_desc_from_json(const char *mi_json, const char *vwmap_json, int immutable, void *out_v, void **keep, char *err, size_t errcap)
{
    fwgpu_model_desc *out = (fwgpu_model_desc *)out_v;
    try {
        ModelInstanceH mi = mi_from_json(json_parse(mi_json));
        VwMap vw = vwmap_from_json(json_parse(vwmap_json));
        FlatDesc *f = new FlatDesc();
        mi_to_desc(mi, vw, immutable != 0, *f);
        *out = f->d;
        *keep = f;
        return 0;
    } catch (const std::exception &e) { set_err(err, errcap, e.what()); return -1; }
}
void fwhost_model_desc_free(void *keep) { delete (FlatDesc *)keep; }

} // extern "C"
