// fwgpu -- command-line front end with the reference's flags (cmdline.rs) for the train / predict loop of
// main.rs:96-292, driving the CUDA hot path through the C ABI (include/fwgpu.h) and the host layer
// (include/fwhost.h).  Differences from `fw` are deliberate and small: examples go to the device in mini-batches
// (--batch_size, default 65536) and are trained Hogwild-style on the device; --sequential reproduces the
// reference's one-example-at-a-time semantics bit for bit (slow).
#include "../../include/fwgpu.h"
#include "../../include/fwhost.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

namespace {
[[noreturn]] void die(const std::string &m) { fprintf(stderr, "fwgpu: %s\n", m.c_str()); exit(1); }
std::string read_file(const std::string &p) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) die("cannot open " + p);
    std::string s; char buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}
bool exists(const std::string &p) { FILE *f = fopen(p.c_str(), "rb"); if (f) fclose(f); return f != nullptr; }
// text input by file extension, like buffer_handler.rs:8-36: .vw plain, .gz gzip (all members, MultiGzDecoder), .zst zstd
std::string read_input(const std::string &p) {
    const size_t dot = p.rfind('.');
    const std::string ext = dot == std::string::npos ? "" : p.substr(dot + 1);
    if (ext == "vw") return read_file(p);
    if (ext == "gz") {
        gzFile g = gzopen(p.c_str(), "rb");
        if (!g) die("Could not open the input file.");
        std::string s; std::vector<char> buf(1 << 20); int n;
        while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) s.append(buf.data(), (size_t)n);
        if (n < 0) { int e = 0; const char *m = gzerror(g, &e); die(std::string("gzip input: ") + (m ? m : "read error")); }
        gzclose(g);
        return s;
    }
    if (ext == "zst") die("zstd input is not supported by this build (no libzstd in the image); use .vw or .gz");
    die("Please specify a valid input format (.vw, .zst, .gz)"); // buffer_handler.rs:33-35
}
struct Flags {
    std::vector<std::pair<std::string, std::string>> kv;
    std::vector<std::string> raw;
    const char *get(const char *k) const { const char *r = nullptr; for (auto &p : kv) if (p.first == k) r = p.second.c_str(); return r; }
    bool has(const char *k) const { for (auto &p : kv) if (p.first == k) return true; return false; }
};
const char *VALUE[] = {"data", "predictions", "final_regressor", "initial_regressor", "predictions_after", "holdout_after", "convert_inference_regressor",
                       "batch_size", "device", "hogwild_threads", "prediction_model_delay", nullptr};
const char *BOOLS[] = {"cache", "testonly", "save_resume", "quiet", "predictions_stdout", "build_cache_without_training", "sequential", "hogwild_training", "weight_quantization", nullptr};
bool in(const char **l, const std::string &s) { for (; *l; l++) if (s == *l) return true; return false; }
std::string long_name(const std::string &t) {
    if (t.rfind("--", 0) == 0) return t.substr(2);
    if (t.size() == 2 && t[0] == '-') switch (t[1]) {
        case 'd': return "data"; case 'p': return "predictions"; case 'f': return "final_regressor"; case 'i': return "initial_regressor";
        case 'c': return "cache"; case 't': return "testonly"; case 'b': return "bit_precision"; case 'l': return "learning_rate"; case 'q': return "interactions";
    }
    return "";
}
void check(fwgpu_ctx *ctx, fwgpu_status st, const char *what) { if (st != FWGPU_OK) die(std::string(what) + ": " + fwgpu_last_error(ctx)); }
} // namespace

int main(int argc, char **argv)
{
    // The whole command line (short flags normalised to their long names) goes to the host layer, which knows the
    // flag table of cmdline.rs and rejects anything else; the driver flags are picked out here.
    Flags fl;
    std::vector<const char *> model_args;
    for (int i = 1; i < argc; i++) {
        std::string name = long_name(argv[i]);
        if (name.empty()) { fl.raw.push_back(argv[i]); continue; } // a value token
        fl.raw.push_back("--" + name);
        std::string key = name, val;
        size_t eq = name.find('=');
        if (eq != std::string::npos) { key = name.substr(0, eq); val = name.substr(eq + 1); }
        if (in(BOOLS, key)) fl.kv.emplace_back(key, "");
        else if (in(VALUE, key)) {
            if (eq == std::string::npos) { if (i + 1 >= argc) die("The argument '--" + key + "' requires a value but none was supplied"); val = argv[++i]; fl.raw.push_back(val); }
            fl.kv.emplace_back(key, val);
        }
    }
    for (auto &s : fl.raw) model_args.push_back(s.c_str());
    char err[1024] = {0};
    const bool testonly = fl.has("testonly"), quiet = fl.has("quiet");
    const char *final_regressor = fl.get("final_regressor");
    if (final_regressor && !fl.has("save_resume")) die("You need to use --save_resume with --final_regressor, for vowpal wabbit compatibility"); // main.rs:112-115
    const int device = fl.get("device") ? atoi(fl.get("device")) : 0;

    // ---- model: from a regressor file or from the command line (main.rs:153-172)
    std::string vwmap_json, mi_json;
    void *reader = nullptr;
    const bool convert = fl.has("convert_inference_regressor");
    if (const char *init = fl.get("initial_regressor")) {
        reader = fwhost_regressor_open(init, err, sizeof(err));
        if (!reader) die(err);
        vwmap_json = fwhost_regressor_vwmap_json(reader);
        char *u = fwhost_model_instance_update_from_cmdline(fwhost_regressor_mi_json(reader), (int)model_args.size(), model_args.data(), err, sizeof(err));
        if (!u) die(err);
        mi_json = u; fwhost_free(u);
    } else {
        if (convert) die("Convert mode requires --initial regressor");
        const char *data = fl.get("data");
        if (!data) die("--data expected");
        std::string dir = data; size_t slash = dir.find_last_of('/'); dir = slash == std::string::npos ? "." : dir.substr(0, slash);
        std::string csv_path = dir + "/vw_namespace_map.csv";
        if (!exists(csv_path)) die("Could not find vw_namespace_map.csv in input dataset directory of \"" + csv_path + "\"");
        char *vj = fwhost_vwmap_csv_to_json(read_file(csv_path).c_str(), err, sizeof(err));
        if (!vj) die(err);
        vwmap_json = vj; fwhost_free(vj);
        char *mj = fwhost_model_instance_from_cmdline((int)model_args.size(), model_args.data(), vwmap_json.c_str(), err, sizeof(err));
        if (!mj) die(err);
        mi_json = mj; fwhost_free(mj);
    }
    const bool immutable = testonly || convert;
    // what the file holds: accumulators unless its ModelInstance says SGD
    const bool file_sgd = reader && fwhost_regressor_optimizer(reader) == FWGPU_OPT_SGD;
    // persistence.rs:144-161: a file written with --weight_quantization holds the FFM block as 16-bit buckets (never when converting)
    const bool file_quantized = reader && fwhost_regressor_dequantize(reader) && !convert;
    fwgpu_model_desc desc; void *keep = nullptr;
    if (fwhost_model_desc_from_json(mi_json.c_str(), vwmap_json.c_str(), immutable ? 1 : 0, &desc, &keep, err, sizeof(err))) die(err);
    if (fl.has("sequential")) desc.hogwild_ramp_div = 0x7fffffffu;
    fwgpu_ctx *ctx = nullptr;
    if (fwgpu_create(&desc, device, &ctx) != FWGPU_OK) die(std::string("fwgpu_create: ") + fwgpu_last_error(nullptr));
    // blocks with weights in execution order (regressor.rs:426-442): LR, FFM, the head's neuron layers
    std::vector<int> blocks{FWGPU_BLOCK_LR};
    if (desc.ffm_k > 0) blocks.push_back(FWGPU_BLOCK_FFM);
    if (desc.nn_num_layers > 0) for (uint32_t l = 0; l <= desc.nn_num_layers; l++) blocks.push_back(FWGPU_BLOCK_NN0 + (int)l);
    const int n_blocks = (int)blocks.size();
    if (reader) { // overwrite_weights_from_buf (regressor.rs:444-469)
        uint64_t expect = 0;
        for (int b = 0; b < n_blocks; b++) { uint64_t n, by; check(ctx, fwgpu_block_len(ctx, blocks[b], &n, &by), "block_len"); expect += n; }
        if (fwhost_regressor_weights_len(reader) != expect) die("Lenghts of weights array in regressor file differ: got " + std::to_string(fwhost_regressor_weights_len(reader)) + ", expected " + std::to_string(expect));
        for (int b = 0; b < n_blocks; b++) {
            uint64_t n, by; fwgpu_block_len(ctx, blocks[b], &n, &by);
            const bool want_state = !file_sgd && !immutable;
            std::vector<float> buf((size_t)n * (file_sgd ? 1 : 2));
            if (file_quantized && blocks[b] == FWGPU_BLOCK_FFM) {
                if (!file_sgd) die("a quantized regressor file must be an inference (SGD) regressor");
                if (fwhost_regressor_read_quantized(reader, buf.data(), n)) die("truncated regressor file");
            } else if (fwhost_regressor_read(reader, buf.data(), buf.size() * 4)) die("truncated regressor file");
            if (!file_sgd && !want_state) { // drop the accumulators (read_weights_from_buf_into_forward_only)
                if (blocks[b] == FWGPU_BLOCK_LR) { for (uint64_t i = 0; i < n; i++) buf[i] = buf[2 * i]; }
                buf.resize(n);
            }
            check(ctx, fwgpu_import_block(ctx, blocks[b], buf.data(), buf.size() * 4, want_state ? 1 : 0), "import_block");
        }
        fwhost_regressor_close(reader);
    }
    // --weight_quantization (main.rs:109, block_ffm.rs:835-848): the FFM weights go to the file as 16-bit buckets.  Only the
    // conversion to an inference regressor also records that in the ModelInstance (main.rs:143-145), so that is the use the
    // reference suggests; with --final_regressor the reference writes the buckets under an unchanged ModelInstance and so do we.
    const bool quantize = fl.has("weight_quantization");
    auto save = [&](const char *path, bool as_sgd, bool mark_quantized) {
        std::vector<std::vector<float>> payload(n_blocks);
        std::vector<const void *> ptrs; std::vector<uint64_t> sizes; uint64_t total = 0;
        std::vector<uint8_t> buckets;
        for (int b = 0; b < n_blocks; b++) {
            uint64_t n, by; fwgpu_block_len(ctx, blocks[b], &n, &by);
            payload[b].resize(by / 4);
            check(ctx, fwgpu_export_block(ctx, blocks[b], payload[b].data(), by), "export_block");
            total += n;
            if (quantize && blocks[b] == FWGPU_BLOCK_FFM) { // weights first, accumulators (if any) after them: block_ffm.rs:835-848
                float mean = 0.0f;
                buckets.resize(8 + 2 * (size_t)n);
                if (fwhost_quantize_ffm_weights(payload[b].data(), n, buckets.data(), &mean)) die("cannot quantize an empty FFM block");
                if (std::fabs(mean) > 10.0f) fprintf(stderr, "fwgpu: warning: mean FFM weight %g, the weights look exploded (quantization.rs:45-47)\n", mean);
                ptrs.push_back(buckets.data()); sizes.push_back(buckets.size());
                if (by > n * 4) { ptrs.push_back(payload[b].data() + n); sizes.push_back(by - n * 4); }
            } else { ptrs.push_back(payload[b].data()); sizes.push_back(by); }
        }
        char *mj = fwhost_model_instance_for_save(mi_json.c_str(), as_sgd ? 1 : 0, mark_quantized ? 1 : 0, err, sizeof(err));
        if (!mj) die(err);
        const int rc = fwhost_regressor_write(path, vwmap_json.c_str(), mj, total, ptrs.data(), sizes.data(), (uint32_t)ptrs.size(), err, sizeof(err));
        fwhost_free(mj);
        if (rc) die(err);
    };
    if (convert) { save(fl.get("convert_inference_regressor"), true, quantize); fwgpu_destroy(ctx); return 0; }

    // ---- input: cache or text (main.rs:173-239, cache.rs:68-131)
    const char *data = fl.get("data");
    if (!data) die("--data expected");
    const std::string cache_path = std::string(data) + ".fwcache";
    uint32_t *records = nullptr, *rec_off = nullptr; uint64_t n_words = 0; int64_t n_examples = -1;
    if (fl.has("cache") && exists(cache_path)) {
        n_examples = fwhost_cache_read(cache_path.c_str(), vwmap_json.c_str(), &records, &n_words, &rec_off, nullptr, err, sizeof(err));
        if (n_examples < 0 && !quiet) fprintf(stderr, "fwgpu: couldn't use the existing cache file: %s\n", err); // cache.rs:99-105: fall back to text
    }
    if (n_examples < 0) {
        std::string text = read_input(data);
        void *parser = fwhost_parser_new(vwmap_json.c_str(), err, sizeof(err));
        if (!parser) die(err);
        uint64_t lines = std::count(text.begin(), text.end(), '\n') + 1;
        uint64_t cap = text.size() + (desc.n_namespaces + 4) * lines + 16;
        records = (uint32_t *)malloc(cap * 4); rec_off = (uint32_t *)malloc((lines + 1) * 4);
        n_examples = fwhost_parser_parse_text(parser, text.data(), text.size(), records, cap, rec_off, lines, 0, &n_words, err, sizeof(err));
        fwhost_parser_free(parser);
        if (n_examples < 0) die(err);
        if (fl.has("cache") && fwhost_cache_write(cache_path.c_str(), vwmap_json.c_str(), records, n_words, err, sizeof(err))) die(err);
    }
    if (fl.has("build_cache_without_training")) { if (!quiet) fprintf(stderr, "fwgpu: cache written, %lld rows\n", (long long)n_examples); return 0; }

    // ---- the loop (main.rs:213-270), in mini-batches
    FILE *pf = fl.get("predictions") ? fopen(fl.get("predictions"), "w") : nullptr;
    if (fl.get("predictions") && !pf) die(std::string("cannot create ") + fl.get("predictions"));
    const uint64_t predictions_after = fl.get("predictions_after") ? strtoull(fl.get("predictions_after"), nullptr, 10) : 0;
    const uint64_t holdout_after = fl.get("holdout_after") ? strtoull(fl.get("holdout_after"), nullptr, 10) : UINT64_MAX;
    const uint64_t batch = fl.get("batch_size") ? strtoull(fl.get("batch_size"), nullptr, 10) : 65536;
    float *preds = nullptr;
    if (fwgpu_host_alloc((void **)&preds, std::max<uint64_t>(batch, 1) * 4) != FWGPU_OK) die("pinned allocation failed");
    auto t0 = std::chrono::steady_clock::now();
    // --prediction_model_delay D (main.rs:200-258): example i is scored by a model that has learned examples 1 .. i-D-1; the
    // example D places back is learned right after.  In mini-batch form: score [done, done+cnt) with cnt <= D, then learn
    // [done-D, done+cnt-D): every example sees at least the reference's delay, the first of each batch exactly it.  The last
    // D examples are never learned, as in the reference.
    const uint64_t delay = fl.get("prediction_model_delay") ? strtoull(fl.get("prediction_model_delay"), nullptr, 10) : 0;
    for (uint64_t done = 0; delay && done < (uint64_t)n_examples;) {
        const uint64_t cnt = std::min<uint64_t>(std::min<uint64_t>(batch, delay), (uint64_t)n_examples - done);
        check(ctx, fwgpu_learn_records(ctx, records, n_words, rec_off + done, (uint32_t)cnt, preds, 0), "learn_records");
        if (!testonly && done + cnt > delay) { // examples (done - D, done + cnt - D] in 1-based numbering
            const uint64_t a = done > delay ? done - delay : 0, b = done + cnt - delay;
            check(ctx, fwgpu_learn_records(ctx, records, n_words, rec_off + a, (uint32_t)(b - a), nullptr, 1), "learn_records");
        }
        check(ctx, fwgpu_sync(ctx), "sync");
        for (uint64_t i = 0; i < cnt; i++) {
            if (done + i + 1 > predictions_after) {
                if (fl.has("predictions_stdout")) printf("%.6f\n", preds[i]);
                if (pf) fprintf(pf, "%.6f\n", preds[i]);
            }
        }
        done += cnt;
    }
    for (uint64_t done = 0; !delay && done < (uint64_t)n_examples;) {
        uint64_t cnt = std::min<uint64_t>(batch, (uint64_t)n_examples - done);
        // example numbers are 1-based in the reference: update while example_num < holdout_after (main.rs:241-244)
        bool update = !testonly;
        if (update && holdout_after != UINT64_MAX) {
            const uint64_t first_num = done + 1;
            if (first_num >= holdout_after) update = false;
            else cnt = std::min<uint64_t>(cnt, holdout_after - first_num);
        }
        // rec_off holds absolute word offsets into `records`: pass the unshifted base together with the offset window
        check(ctx, fwgpu_learn_records(ctx, records, n_words, rec_off + done, (uint32_t)cnt, preds, update ? 1 : 0), "learn_records");
        check(ctx, fwgpu_sync(ctx), "sync");
        for (uint64_t i = 0; i < cnt; i++) {
            if (done + i + 1 > predictions_after) {
                if (fl.has("predictions_stdout")) printf("%.6f\n", preds[i]);
                if (pf) fprintf(pf, "%.6f\n", preds[i]); // main.rs:260-269 "{:.6}"
            }
        }
        done += cnt;
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!quiet) fprintf(stderr, "fwgpu: Elapsed: %.2fs rows: %lld (%.0f rows/s)\n", secs, (long long)n_examples, n_examples / std::max(secs, 1e-9));
    if (pf) fclose(pf);
    if (final_regressor) save(final_regressor, /*as_sgd=*/immutable, /*mark_quantized=*/false); // an immutable ctx holds weights only: the file says SGD (persistence.rs:163-172)
    fwgpu_host_free(preds);
    fwhost_free(records); fwhost_free(rec_off);
    fwhost_model_desc_free(keep);
    fwgpu_destroy(ctx);
    return 0;
}
