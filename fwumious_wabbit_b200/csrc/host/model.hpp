// Host-side model description: VwNamespaceMap (vwmap.rs) and ModelInstance (model_instance.rs), with the JSON
// forms the reference stores in its cache and regressor files, and the command-line subset of cmdline.rs /
// ModelInstance::new_from_cmdline (model_instance.rs:296-495) that concerns the LR/FFM path.
#pragma once
#include "../../../include/fwgpu.h"
#include "json.hpp"

#include <map>
#include <string>
#include <vector>

namespace fwhost {

struct VwEntry { std::string vwname, verbose; uint32_t index = 0; bool f32 = false; };
struct VwMap {
    uint32_t namespace_skip_prefix = 0;
    std::vector<VwEntry> entries;
    uint32_t num_namespaces = 0;

    void finish() { num_namespaces = 0; for (auto &e : entries) num_namespaces = std::max(num_namespaces, e.index); num_namespaces += 1; }
    const VwEntry *by_vwname(const std::string &n) const { for (auto &e : entries) if (e.vwname == n) return &e; return nullptr; }
    const VwEntry *by_verbose(const std::string &n) const { for (auto &e : entries) if (e.verbose == n) return &e; return nullptr; }
    bool operator==(const VwMap &o) const
    {
        if (namespace_skip_prefix != o.namespace_skip_prefix || entries.size() != o.entries.size()) return false;
        for (size_t i = 0; i < entries.size(); i++)
            if (entries[i].vwname != o.entries[i].vwname || entries[i].verbose != o.entries[i].verbose || entries[i].index != o.entries[i].index || entries[i].f32 != o.entries[i].f32) return false;
        return true;
    }
};

// vw_namespace_map.csv (vwmap.rs:106-151): "vwname,verbose[,f32]", "_namespace_skip_prefix,N"; index = line number
inline VwMap vwmap_from_csv(const std::string &csv)
{
    VwMap m;
    size_t pos = 0;
    uint32_t i = 0;
    while (pos < csv.size()) {
        size_t nl = csv.find('\n', pos);
        std::string line = csv.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = nl == std::string::npos ? csv.size() : nl + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue; // the csv crate skips empty lines
        std::vector<std::string> f;
        size_t a = 0;
        for (;;) { size_t c = line.find(',', a); f.push_back(line.substr(a, c == std::string::npos ? std::string::npos : c - a)); if (c == std::string::npos) break; a = c + 1; }
        if (f[0] == "_namespace_skip_prefix") {
            if (f.size() < 2) throw std::runtime_error("Couldn't parse _namespace_skip_prefix in vw_namespaces_map.csv");
            m.namespace_skip_prefix = (uint32_t)strtoul(f[1].c_str(), nullptr, 10);
            i++; // the reference enumerates csv records, the skip-prefix line consumes an index too (vwmap.rs:115-131)
            continue;
        }
        if (f.size() < 2) throw std::runtime_error("vw_namespace_map.csv: expected vwname,verbose");
        VwEntry e;
        e.vwname = f[0]; e.verbose = f[1]; e.index = i;
        if (f.size() > 2 && !f[2].empty()) {
            if (f[2] != "f32") throw std::runtime_error("Unknown type used for the feature in vw_namespace_map.csv: \"" + f[2] + "\". Only \"f32\" is possible.");
            e.f32 = true;
        }
        m.entries.push_back(e);
        i++;
    }
    m.finish();
    return m;
}

inline JValue vwmap_to_json(const VwMap &m) // VwNamespaceMapSource (vwmap.rs:39-51)
{
    JValue o = JValue::object();
    o.set("namespace_skip_prefix", JValue::integer(m.namespace_skip_prefix));
    JValue arr = JValue::array();
    for (auto &e : m.entries) {
        JValue je = JValue::object();
        je.set("namespace_vwname", JValue::string(e.vwname));
        je.set("namespace_verbose", JValue::string(e.verbose));
        je.set("namespace_index", JValue::integer(e.index));
        je.set("namespace_format", JValue::string(e.f32 ? "F32" : "Categorical"));
        arr.arr.push_back(je);
    }
    o.set("entries", arr);
    return o;
}
inline VwMap vwmap_from_json(const JValue &j)
{
    VwMap m;
    m.namespace_skip_prefix = (uint32_t)j.at("namespace_skip_prefix").as_num();
    for (auto &je : j.at("entries").arr) {
        VwEntry e;
        e.vwname = je.at("namespace_vwname").as_str();
        e.verbose = je.at("namespace_verbose").as_str();
        e.index = (uint32_t)je.at("namespace_index").as_num();
        e.f32 = je.at("namespace_format").as_str() == "F32";
        m.entries.push_back(e);
    }
    m.finish();
    return m;
}

struct NsDesc { uint32_t index = 0; bool f32 = false; }; // NamespaceDescriptor, Primitive only (vwmap.rs:22-27)
struct ComboDesc { std::vector<NsDesc> ns; float weight = 1.0f; };

struct ModelInstanceH { // model_instance.rs:47-97, defaults of new_empty :120-150
    float learning_rate = 0.5f, minimum_learning_rate = 0.0f, power_t = 0.5f;
    uint32_t bit_precision = 18;
    bool add_constant_feature = true;
    std::vector<ComboDesc> feature_combo_descs;
    std::vector<std::vector<NsDesc>> ffm_fields;
    uint32_t ffm_k = 0, ffm_bit_precision = 18;
    bool fastmath = true;
    std::string ffm_initialization_type = "default";
    float ffm_k_threshold = 0.0f, ffm_init_center = 0.0f, ffm_init_width = 0.0f, ffm_init_zero_band = 0.0f;
    float ffm_init_acc_gradient = 0.0f, init_acc_gradient = 1.0f, ffm_learning_rate = 0.5f, ffm_power_t = 0.5f;
    float nn_init_acc_gradient = 0.0f, nn_learning_rate = 0.02f, nn_power_t = 0.45f;
    std::vector<std::vector<std::pair<std::string, std::string>>> nn_layers;
    std::string nn_topology = "one";
    uint32_t optimizer = FWGPU_OPT_SGD;
    int dequantize_weights = 0; // 0 = Some(false), 1 = Some(true), -1 = None
};

inline JValue nsdesc_to_json(const NsDesc &d)
{
    JValue o = JValue::object();
    o.set("namespace_index", JValue::integer(d.index));
    o.set("namespace_type", JValue::string("Primitive"));
    o.set("namespace_format", JValue::string(d.f32 ? "F32" : "Categorical"));
    return o;
}
inline NsDesc nsdesc_from_json(const JValue &j)
{
    NsDesc d;
    d.index = (uint32_t)j.at("namespace_index").as_num();
    if (j.at("namespace_type").as_str() != "Primitive") throw std::runtime_error("transformed namespaces are out of scope for the GPU path (feature_transform_*.rs)");
    d.f32 = j.at("namespace_format").as_str() == "F32";
    return d;
}

inline const char *optimizer_name(uint32_t o) { return o == FWGPU_OPT_SGD ? "SGD" : o == FWGPU_OPT_ADAGRAD_FLEX ? "AdagradFlex" : "AdagradLUT"; }

inline JValue mi_to_json(const ModelInstanceH &m)
{
    JValue o = JValue::object();
    o.set("learning_rate", JValue::f32(m.learning_rate));
    o.set("minimum_learning_rate", JValue::f32(m.minimum_learning_rate));
    o.set("power_t", JValue::f32(m.power_t));
    o.set("bit_precision", JValue::integer(m.bit_precision));
    o.set("add_constant_feature", JValue::boolean(m.add_constant_feature));
    JValue combos = JValue::array();
    for (auto &c : m.feature_combo_descs) {
        JValue jc = JValue::object(), nd = JValue::array();
        for (auto &d : c.ns) nd.arr.push_back(nsdesc_to_json(d));
        jc.set("namespace_descriptors", nd);
        jc.set("weight", JValue::f32(c.weight));
        combos.arr.push_back(jc);
    }
    o.set("feature_combo_descs", combos);
    JValue fields = JValue::array();
    for (auto &f : m.ffm_fields) { JValue jf = JValue::array(); for (auto &d : f) jf.arr.push_back(nsdesc_to_json(d)); fields.arr.push_back(jf); }
    o.set("ffm_fields", fields);
    o.set("ffm_k", JValue::integer(m.ffm_k));
    o.set("ffm_bit_precision", JValue::integer(m.ffm_bit_precision));
    o.set("fastmath", JValue::boolean(m.fastmath));
    o.set("ffm_initialization_type", JValue::string(m.ffm_initialization_type));
    o.set("ffm_k_threshold", JValue::f32(m.ffm_k_threshold));
    o.set("ffm_init_center", JValue::f32(m.ffm_init_center));
    o.set("ffm_init_width", JValue::f32(m.ffm_init_width));
    o.set("ffm_init_zero_band", JValue::f32(m.ffm_init_zero_band));
    o.set("ffm_init_acc_gradient", JValue::f32(m.ffm_init_acc_gradient));
    o.set("init_acc_gradient", JValue::f32(m.init_acc_gradient));
    o.set("ffm_learning_rate", JValue::f32(m.ffm_learning_rate));
    o.set("ffm_power_t", JValue::f32(m.ffm_power_t));
    o.set("nn_init_acc_gradient", JValue::f32(m.nn_init_acc_gradient));
    o.set("nn_learning_rate", JValue::f32(m.nn_learning_rate));
    o.set("nn_power_t", JValue::f32(m.nn_power_t));
    JValue nn = JValue::object(), layers = JValue::array();
    for (auto &l : m.nn_layers) { JValue jl = JValue::object(); for (auto &kv : l) jl.set(kv.first, JValue::string(kv.second)); layers.arr.push_back(jl); }
    nn.set("layers", layers);
    nn.set("topology", JValue::string(m.nn_topology));
    o.set("nn_config", nn);
    o.set("optimizer", JValue::string(optimizer_name(m.optimizer)));
    JValue tn = JValue::object();
    tn.set("v", JValue::array());
    o.set("transform_namespaces", tn);
    o.set("dequantize_weights", m.dequantize_weights < 0 ? JValue::null() : JValue::boolean(m.dequantize_weights != 0));
    return o;
}

inline float jf(const JValue &j, const char *k, float dflt) { const JValue *v = j.get(k); return (v && v->t == JValue::Num) ? (float)v->num : dflt; }

inline ModelInstanceH mi_from_json(const JValue &j)
{
    ModelInstanceH m;
    m.learning_rate = (float)j.at("learning_rate").as_num();
    m.minimum_learning_rate = jf(j, "minimum_learning_rate", 0.0f); // #[serde(default)] fields may be missing
    m.power_t = (float)j.at("power_t").as_num();
    m.bit_precision = (uint32_t)j.at("bit_precision").as_num();
    m.add_constant_feature = j.at("add_constant_feature").as_bool();
    for (auto &jc : j.at("feature_combo_descs").arr) {
        ComboDesc c;
        for (auto &d : jc.at("namespace_descriptors").arr) c.ns.push_back(nsdesc_from_json(d));
        c.weight = (float)jc.at("weight").as_num();
        m.feature_combo_descs.push_back(c);
    }
    for (auto &jfld : j.at("ffm_fields").arr) { std::vector<NsDesc> f; for (auto &d : jfld.arr) f.push_back(nsdesc_from_json(d)); m.ffm_fields.push_back(f); }
    m.ffm_k = (uint32_t)jf(j, "ffm_k", 0);
    m.ffm_bit_precision = (uint32_t)jf(j, "ffm_bit_precision", 0);
    if (const JValue *v = j.get("fastmath")) m.fastmath = v->t == JValue::Bool ? v->b : false; else m.fastmath = false;
    m.ffm_initialization_type = j.at("ffm_initialization_type").as_str();
    m.ffm_k_threshold = jf(j, "ffm_k_threshold", 0); m.ffm_init_center = jf(j, "ffm_init_center", 0);
    m.ffm_init_width = jf(j, "ffm_init_width", 0); m.ffm_init_zero_band = jf(j, "ffm_init_zero_band", 0);
    m.ffm_init_acc_gradient = jf(j, "ffm_init_acc_gradient", 0); m.init_acc_gradient = jf(j, "init_acc_gradient", 0);
    m.ffm_learning_rate = jf(j, "ffm_learning_rate", 0); m.ffm_power_t = jf(j, "ffm_power_t", 0);
    m.nn_init_acc_gradient = jf(j, "nn_init_acc_gradient", 0); m.nn_learning_rate = jf(j, "nn_learning_rate", 0); m.nn_power_t = jf(j, "nn_power_t", 0);
    const JValue &nn = j.at("nn_config");
    for (auto &jl : nn.at("layers").arr) { std::vector<std::pair<std::string, std::string>> l; for (auto &kv : jl.obj) l.emplace_back(kv.first, kv.second.as_str()); m.nn_layers.push_back(l); }
    m.nn_topology = nn.at("topology").as_str();
    if (const JValue *v = j.get("optimizer")) {
        const std::string &s = v->as_str();
        m.optimizer = s == "SGD" ? FWGPU_OPT_SGD : s == "AdagradFlex" ? FWGPU_OPT_ADAGRAD_FLEX : s == "AdagradLUT" ? FWGPU_OPT_ADAGRAD_LUT : 99;
        if (m.optimizer == 99) throw std::runtime_error("unknown optimizer \"" + s + "\"");
    } else m.optimizer = FWGPU_OPT_ADAGRAD_FLEX; // default_optimizer_adagrad (model_instance.rs:108-110)
    if (const JValue *v = j.get("transform_namespaces")) { if (const JValue *vv = v->get("v")) if (!vv->arr.empty()) throw std::runtime_error("transformed namespaces are out of scope for the GPU path"); }
    if (const JValue *v = j.get("dequantize_weights")) m.dequantize_weights = v->t == JValue::Bool ? (v->b ? 1 : 0) : -1;
    return m;
}

// Flattened descriptor for fwgpu_create; the vectors own the arrays the desc points to.
struct FlatDesc {
    fwgpu_model_desc d{};
    std::vector<uint8_t> ns_is_f32;
    std::vector<uint32_t> combo_off, combo_ns, field_off, field_ns;
    std::vector<float> combo_weight;
};
inline void mi_to_desc(const ModelInstanceH &m, const VwMap &vw, bool immutable, FlatDesc &f)
{
    fwgpu_model_desc &d = f.d;
    memset(&d, 0, sizeof(d));
    d.learning_rate = m.learning_rate; d.power_t = m.power_t; d.init_acc_gradient = m.init_acc_gradient;
    d.ffm_learning_rate = m.ffm_learning_rate; d.ffm_power_t = m.ffm_power_t; d.ffm_init_acc_gradient = m.ffm_init_acc_gradient;
    d.nn_learning_rate = m.nn_learning_rate; d.nn_power_t = m.nn_power_t; d.nn_init_acc_gradient = m.nn_init_acc_gradient;
    d.bit_precision = m.bit_precision; d.ffm_bit_precision = m.ffm_bit_precision; d.ffm_k = m.ffm_k;
    d.ffm_num_fields = m.ffm_k ? (uint32_t)m.ffm_fields.size() : 0;
    d.num_combos = (uint32_t)m.feature_combo_descs.size() + (m.add_constant_feature ? 1 : 0);
    d.optimizer = m.optimizer; d.immutable = immutable ? 1 : 0;
    d.ffm_init_width = m.ffm_init_width; d.ffm_init_zero_band = m.ffm_init_zero_band; d.ffm_init_center = m.ffm_init_center;
    d.nn_num_layers = (uint32_t)m.nn_layers.size();
    if (d.nn_num_layers > FWGPU_MAX_NN_LAYERS) throw std::runtime_error("too many --nn_layers for the CUDA head");
    if (d.nn_num_layers && m.nn_topology != "one") throw std::runtime_error("only nn topology \"one\" is implemented by the CUDA head");
    for (uint32_t i = 0; i < d.nn_num_layers; i++) { // per-layer defaults: regressor.rs:217-251
        d.nn_width[i] = 20; d.nn_relu[i] = 0; d.nn_init[i] = FWGPU_NN_INIT_HU;
        for (auto &kv : m.nn_layers[i]) {
            if (kv.first == "width") d.nn_width[i] = (uint32_t)strtoul(kv.second.c_str(), nullptr, 10);
            else if (kv.first == "activation") {
                if (kv.second == "relu") d.nn_relu[i] = 1;
                else if (kv.second != "none") throw std::runtime_error("unknown nn activation type: \"" + kv.second + "\"");
            } else if (kv.first == "init") {
                if (kv.second == "xavier") d.nn_init[i] = FWGPU_NN_INIT_XAVIER; else if (kv.second == "hu") d.nn_init[i] = FWGPU_NN_INIT_HU;
                else if (kv.second == "one") d.nn_init[i] = FWGPU_NN_INIT_ONE; else if (kv.second == "zero") d.nn_init[i] = FWGPU_NN_INIT_ZERO;
                else throw std::runtime_error("unknown nn initialization type: \"" + kv.second + "\"");
            } else if (kv.first == "dropout" || kv.first == "maxnorm") {
                if (strtof(kv.second.c_str(), nullptr) != 0.0f) throw std::runtime_error("--nn " + kv.first + " is not implemented by the CUDA head");
            } else if (kv.first == "layernorm") {
                if (kv.second != "none") throw std::runtime_error("--nn layernorm is not implemented by the CUDA head");
            } else throw std::runtime_error("Unknown --nn parameter for layer number " + std::to_string(i) + " : " + kv.first);
        }
    }
    f.ns_is_f32.assign(vw.num_namespaces, 0);
    for (auto &e : vw.entries) if (e.index < vw.num_namespaces) f.ns_is_f32[e.index] = e.f32 ? 1 : 0;
    f.combo_off.assign(1, 0);
    for (auto &c : m.feature_combo_descs) { for (auto &n : c.ns) f.combo_ns.push_back(n.index); f.combo_off.push_back((uint32_t)f.combo_ns.size()); f.combo_weight.push_back(c.weight); }
    f.field_off.assign(1, 0);
    if (m.ffm_k) for (auto &fl : m.ffm_fields) { for (auto &n : fl) f.field_ns.push_back(n.index); f.field_off.push_back((uint32_t)f.field_ns.size()); }
    if (f.combo_ns.empty()) f.combo_ns.push_back(0);
    if (f.combo_weight.empty()) f.combo_weight.push_back(0.0f);
    if (f.field_ns.empty()) f.field_ns.push_back(0);
    d.n_namespaces = vw.num_namespaces; d.ns_is_f32 = f.ns_is_f32.data();
    d.n_combos = (uint32_t)m.feature_combo_descs.size(); d.combo_off = f.combo_off.data(); d.combo_ns = f.combo_ns.data(); d.combo_weight = f.combo_weight.data();
    d.add_constant = m.add_constant_feature ? 1 : 0; d.field_off = f.field_off.data(); d.field_ns = f.field_ns.data();
}

} // namespace fwhost
