"""Python face of the C++ host layer (include/fwhost.h): vw_namespace_map, VW text parser, .fwcache,
regressor files, command line -> ModelInstance.  Names follow the reference modules they restate
(vwmap.rs, parser.rs, cache.rs, persistence.rs, model_instance.rs)."""
import ctypes as C
import json
import os

import numpy as np

from . import _lib
from .model_instance import ModelInstance, Optimizer

_ERR = 1024


def _L():
    L = _lib.lib()
    if getattr(L, "_host_bound", False):
        return L
    vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
    L.fwhost_free.argtypes = [vp]
    L.fwhost_free.restype = None
    for name, args in {
        "fwhost_vwmap_csv_to_json": [cp, cp, sz],
        "fwhost_model_instance_from_cmdline": [C.c_int, C.POINTER(cp), cp, cp, sz],
        "fwhost_model_instance_normalize": [cp, cp, sz],
        "fwhost_model_instance_update_from_cmdline": [cp, C.c_int, C.POINTER(cp), cp, sz],
    }.items():
        f = getattr(L, name)
        f.argtypes, f.restype = args, vp
    L.fwhost_parser_new.argtypes, L.fwhost_parser_new.restype = [cp, cp, sz], vp
    L.fwhost_parser_free.argtypes, L.fwhost_parser_free.restype = [vp], None
    L.fwhost_parser_parse_line.argtypes, L.fwhost_parser_parse_line.restype = [vp, cp, sz, vp, sz, cp, sz], C.c_int
    L.fwhost_parser_parse_text.argtypes = [vp, cp, sz, vp, C.c_uint64, vp, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), cp, sz]
    L.fwhost_parser_parse_text.restype = C.c_int64
    L.fwhost_cache_write.argtypes, L.fwhost_cache_write.restype = [cp, cp, vp, C.c_uint64, cp, sz], C.c_int
    L.fwhost_cache_read.argtypes = [cp, cp, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp), cp, sz]
    L.fwhost_cache_read.restype = C.c_int64
    L.fwhost_regressor_write.argtypes = [cp, cp, cp, C.c_uint64, C.POINTER(vp), C.POINTER(C.c_uint64), C.c_uint32, cp, sz]
    L.fwhost_regressor_write.restype = C.c_int
    L.fwhost_regressor_open.argtypes, L.fwhost_regressor_open.restype = [cp, cp, sz], vp
    L.fwhost_regressor_vwmap_json.argtypes, L.fwhost_regressor_vwmap_json.restype = [vp], cp
    L.fwhost_regressor_mi_json.argtypes, L.fwhost_regressor_mi_json.restype = [vp], cp
    L.fwhost_regressor_weights_len.argtypes, L.fwhost_regressor_weights_len.restype = [vp], C.c_uint64
    L.fwhost_regressor_read.argtypes, L.fwhost_regressor_read.restype = [vp, vp, C.c_uint64], C.c_int
    L.fwhost_regressor_skip.argtypes, L.fwhost_regressor_skip.restype = [vp, C.c_uint64], C.c_int
    L.fwhost_regressor_optimizer.argtypes, L.fwhost_regressor_optimizer.restype = [vp], C.c_uint32
    L.fwhost_regressor_dequantize.argtypes, L.fwhost_regressor_dequantize.restype = [vp], C.c_int
    L.fwhost_regressor_read_quantized.argtypes, L.fwhost_regressor_read_quantized.restype = [vp, vp, C.c_uint64], C.c_int
    L.fwhost_quantize_ffm_weights.argtypes, L.fwhost_quantize_ffm_weights.restype = [vp, C.c_uint64, vp, vp], C.c_int
    L.fwhost_model_instance_for_save.argtypes, L.fwhost_model_instance_for_save.restype = [cp, C.c_int, C.c_int, cp, sz], vp
    L.fwhost_regressor_close.argtypes, L.fwhost_regressor_close.restype = [vp], None
    L._host_bound = True
    return L


def _take(ptr, err):
    if not ptr:
        raise ValueError(err.value.decode())
    s = C.string_at(ptr).decode()
    _L().fwhost_free(ptr)
    return s


# ---------------------------------------------------------------- vwmap.rs
class VwNamespaceMap:
    """vw_namespace_map.csv (vwmap.rs:91-151); `source_json` is VwNamespaceMapSource as the reference serialises it."""

    def __init__(self, source_json: str):
        self.source_json = source_json
        self.source = json.loads(source_json)
        self.num_namespaces = max([e["namespace_index"] for e in self.source["entries"]] + [0]) + 1

    @staticmethod
    def new(csv_text: str):
        err = C.create_string_buffer(_ERR)
        return VwNamespaceMap(_take(_L().fwhost_vwmap_csv_to_json(csv_text.encode(), err, _ERR), err))

    @staticmethod
    def new_from_csv_filepath(path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"Could not find vw_namespace_map.csv in input dataset directory of {path!r}")
        return VwNamespaceMap.new(open(path).read())

    def ns_is_f32(self):
        out = [0] * self.num_namespaces
        for e in self.source["entries"]:
            out[e["namespace_index"]] = 1 if e["namespace_format"] == "F32" else 0
        return out


# ---------------------------------------------------------------- model_instance.rs
def model_instance_json_from_cmdline(argv, vw: VwNamespaceMap) -> str:
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    err = C.create_string_buffer(_ERR)
    return _take(_L().fwhost_model_instance_from_cmdline(len(argv), arr, vw.source_json.encode(), err, _ERR), err)


def model_instance_from_json(mi_json: str, vw: VwNamespaceMap = None) -> ModelInstance:
    j = json.loads(mi_json)
    mi = ModelInstance()
    for k in ("learning_rate", "minimum_learning_rate", "power_t", "bit_precision", "add_constant_feature", "ffm_k",
              "ffm_bit_precision", "fastmath", "ffm_initialization_type", "ffm_k_threshold", "ffm_init_center",
              "ffm_init_width", "ffm_init_zero_band", "ffm_init_acc_gradient", "init_acc_gradient", "ffm_learning_rate",
              "ffm_power_t", "nn_init_acc_gradient", "nn_learning_rate", "nn_power_t"):
        if k in j:
            setattr(mi, k, j[k])
    mi.feature_combo_descs = [([d["namespace_index"] for d in c["namespace_descriptors"]], c["weight"]) for c in j["feature_combo_descs"]]
    mi.ffm_fields = [[d["namespace_index"] for d in f] for f in j["ffm_fields"]]
    mi.nn_layers = [dict(l) for l in j["nn_config"]["layers"]]
    mi.nn_topology = j["nn_config"]["topology"]
    mi.optimizer = {"SGD": Optimizer.SGD, "AdagradFlex": Optimizer.AdagradFlex, "AdagradLUT": Optimizer.AdagradLUT}[j.get("optimizer", "AdagradFlex")]
    if vw is not None:
        mi.num_namespaces = vw.num_namespaces
        mi.ns_is_f32 = vw.ns_is_f32()
    return mi


def model_instance_to_json(mi: ModelInstance, vw: VwNamespaceMap) -> str:
    f32 = vw.ns_is_f32()

    def nd(i):
        return {"namespace_index": i, "namespace_type": "Primitive", "namespace_format": "F32" if f32[i] else "Categorical"}

    j = {
        "learning_rate": mi.learning_rate, "minimum_learning_rate": mi.minimum_learning_rate, "power_t": mi.power_t,
        "bit_precision": mi.bit_precision, "add_constant_feature": mi.add_constant_feature,
        "feature_combo_descs": [{"namespace_descriptors": [nd(i) for i in c[0]], "weight": c[1]} for c in mi.feature_combo_descs],
        "ffm_fields": [[nd(i) for i in f] for f in mi.ffm_fields], "ffm_k": mi.ffm_k, "ffm_bit_precision": mi.ffm_bit_precision,
        "fastmath": mi.fastmath, "ffm_initialization_type": mi.ffm_initialization_type, "ffm_k_threshold": mi.ffm_k_threshold,
        "ffm_init_center": mi.ffm_init_center, "ffm_init_width": mi.ffm_init_width, "ffm_init_zero_band": mi.ffm_init_zero_band,
        "ffm_init_acc_gradient": mi.ffm_init_acc_gradient, "init_acc_gradient": mi.init_acc_gradient,
        "ffm_learning_rate": mi.ffm_learning_rate, "ffm_power_t": mi.ffm_power_t, "nn_init_acc_gradient": mi.nn_init_acc_gradient,
        "nn_learning_rate": mi.nn_learning_rate, "nn_power_t": mi.nn_power_t,
        "nn_config": {"layers": [dict(l) for l in mi.nn_layers], "topology": mi.nn_topology},
        "optimizer": Optimizer.names[mi.optimizer], "transform_namespaces": {"v": []}, "dequantize_weights": False,
    }
    err = C.create_string_buffer(_ERR)
    return _take(_L().fwhost_model_instance_normalize(json.dumps(j).encode(), err, _ERR), err)  # the reference's field order and float formatting


def new_model_instance_from_cmdline(argv, vw: VwNamespaceMap) -> ModelInstance:
    """ModelInstance::new_from_cmdline (model_instance.rs:296-495)."""
    return model_instance_from_json(model_instance_json_from_cmdline(argv, vw), vw)


# ---------------------------------------------------------------- parser.rs
class FlushCommand(Exception):
    pass


class HogwildLoadCommand(Exception):
    pass


class VowpalParser:
    def __init__(self, vw: VwNamespaceMap):
        err = C.create_string_buffer(_ERR)
        self.h = _L().fwhost_parser_new(vw.source_json.encode(), err, _ERR)
        if not self.h:
            raise ValueError(err.value.decode())
        self.vw = vw

    def __del__(self):
        try:
            if self.h:
                _L().fwhost_parser_free(self.h)
                self.h = None
        except Exception:
            pass

    def next_vowpal(self, line):
        """One text line -> the u32 record (parser.rs:158-169); empty input -> empty record (EOF)."""
        if isinstance(line, str):
            line = line.encode()
        out = np.zeros(8192, dtype=np.uint32)
        err = C.create_string_buffer(_ERR)
        n = _L().fwhost_parser_parse_line(self.h, line, len(line), out.ctypes.data_as(C.c_void_p), out.size, err, _ERR)
        if n == -2:
            raise FlushCommand()
        if n == -3:
            raise HogwildLoadCommand()
        if n < 0:
            raise ValueError(err.value.decode())
        return out[:n].copy()

    def parse_text(self, text, threads=0):
        """A whole .vw buffer -> (records, rec_off)."""
        if isinstance(text, str):
            text = text.encode()
        n_lines = text.count(b"\n") + 1
        cap_words = len(text) + (self.vw.num_namespaces + 4) * n_lines + 16  # every feature costs >= 2 bytes of text and <= 2 words
        out = np.empty(cap_words, dtype=np.uint32)
        off = np.empty(n_lines + 1, dtype=np.uint32)
        nw = C.c_uint64(0)
        err = C.create_string_buffer(_ERR)
        n = _L().fwhost_parser_parse_text(self.h, text, len(text), out.ctypes.data_as(C.c_void_p), out.size,
                                          off.ctypes.data_as(C.c_void_p), off.size - 1, threads, C.byref(nw), err, _ERR)
        if n < 0:
            raise ValueError(err.value.decode())
        return out[: nw.value].copy(), off[: n + 1].copy()


# ---------------------------------------------------------------- cache.rs
def cache_write(path, vw: VwNamespaceMap, records):
    records = np.ascontiguousarray(records, dtype=np.uint32)
    err = C.create_string_buffer(_ERR)
    if _L().fwhost_cache_write(path.encode(), vw.source_json.encode(), records.ctypes.data_as(C.c_void_p), records.size, err, _ERR) != 0:
        raise IOError(err.value.decode())


def cache_read(path, vw: VwNamespaceMap = None):
    """Returns (records, rec_off, vwmap_json).  With `vw`, a namespace map that differs from the file's is an error
    (the reference then rebuilds the cache, cache.rs:99-105, 178-182)."""
    recs, offs, blob = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nw = C.c_uint64(0)
    err = C.create_string_buffer(_ERR)
    n = _L().fwhost_cache_read(path.encode(), vw.source_json.encode() if vw else None, C.byref(recs), C.byref(nw), C.byref(offs),
                               C.byref(blob), err, _ERR)
    if n < 0:
        raise IOError(err.value.decode())
    r = np.ctypeslib.as_array(C.cast(recs, C.POINTER(C.c_uint32)), shape=(max(nw.value, 1),))[: nw.value].copy()
    o = np.ctypeslib.as_array(C.cast(offs, C.POINTER(C.c_uint32)), shape=(n + 1,)).copy()
    js = C.string_at(blob).decode()
    for p in (recs, offs, blob):
        _L().fwhost_free(p)
    return r, o, js


# ---------------------------------------------------------------- persistence.rs
def _block_order(mi):
    """Blocks with weights in execution order (regressor.rs:426-442): LR, FFM, then every neuron layer of the head
    (hidden layers, final neuron); triangle / join / copy / relu / sigmoid hold nothing."""
    nn = [_lib.BLOCK_NN0 + l for l in range(len(mi.nn_layers) + 1)] if mi.nn_layers else []
    return [_lib.BLOCK_LR] + ([_lib.BLOCK_FFM] if mi.ffm_k > 0 else []) + nn


def quantize_ffm_weights(weights):
    """quantization.rs:41-75: the 8-byte header {increment, min} and one 16-bit bucket per weight, as bytes (uint8 array)."""
    w = np.ascontiguousarray(weights, dtype=np.float32)
    out = np.empty(8 + 2 * w.size, dtype=np.uint8)
    if _L().fwhost_quantize_ffm_weights(w.ctypes.data_as(C.c_void_p), w.size, out.ctypes.data_as(C.c_void_p), None) != 0:
        raise ValueError("cannot quantize an empty FFM block")
    return out


def save_regressor_to_filename(filename, mi: ModelInstance, vw: VwNamespaceMap, re, quantize_weights=False):
    """persistence.rs:76-92: header, vwmap JSON, ModelInstance JSON, total weight count, block payloads.
    quantize_weights (the reference's --weight_quantization, main.rs:109,143-147) writes the FFM weights as 16-bit buckets
    (block_ffm.rs:835-848); on an immutable regressor - the conversion to an inference regressor, the suggested use - the
    ModelInstance in the file also says dequantize_weights = true, so that loaders know."""
    blocks, total = [], 0
    order = _block_order(mi)
    for b in order:
        n, _ = re.block_len(b)
        total += n
        payload = re.export_block(b)
        if quantize_weights and b == _lib.BLOCK_FFM:
            flat = payload.reshape(-1).view(np.float32)
            blocks.append(quantize_ffm_weights(flat[:n]))
            if flat.size > n:  # accumulators follow the weights unchanged
                blocks.append(np.ascontiguousarray(flat[n:]))
        else:
            blocks.append(payload)
    err = C.create_string_buffer(_ERR)
    # an inference regressor is written with mi.optimizer = SGD (main.rs:140-147)
    p = _L().fwhost_model_instance_for_save(model_instance_to_json(mi, vw).encode(), 1 if re.immutable else 0,
                                            1 if (quantize_weights and re.immutable) else 0, err, _ERR)
    if not p:
        raise ValueError(err.value.decode())
    mi_json = C.string_at(p)
    _L().fwhost_free(p)
    ptrs = (C.c_void_p * len(blocks))(*[b.ctypes.data_as(C.c_void_p) for b in blocks])
    sizes = (C.c_uint64 * len(blocks))(*[b.nbytes for b in blocks])
    if _L().fwhost_regressor_write(filename.encode(), vw.source_json.encode(), mi_json, total, ptrs, sizes, len(blocks), err, _ERR) != 0:
        raise IOError(err.value.decode())


def new_regressor_from_filename(filename, immutable=False, cmd_arguments=None, device=0):
    """persistence.rs:127-174: returns (mi, vw, regressor).  immutable=True builds the forward-only regressor and
    skips the optimizer state while reading (block_lr.rs:277-292, block_ffm.rs:879-899)."""
    from .regressor import Regressor

    err = C.create_string_buffer(_ERR)
    L = _L()
    r = L.fwhost_regressor_open(filename.encode(), err, _ERR)
    if not r:
        raise IOError(err.value.decode())
    try:
        vw = VwNamespaceMap(L.fwhost_regressor_vwmap_json(r).decode())
        mi_json = L.fwhost_regressor_mi_json(r).decode()
        if cmd_arguments:
            arr = (C.c_char_p * len(cmd_arguments))(*[a.encode() for a in cmd_arguments])
            mi_json = _take(L.fwhost_model_instance_update_from_cmdline(mi_json.encode(), len(cmd_arguments), arr, err, _ERR), err)
        mi = model_instance_from_json(mi_json, vw)
        file_has_state = L.fwhost_regressor_optimizer(r) != Optimizer.SGD  # what the WRITER stored: accumulators unless it was SGD
        # persistence.rs:144-161: --weight_quantization files hold the FFM block as 8-byte header + one half per weight
        quantized = bool(L.fwhost_regressor_dequantize(r))
        if quantized and file_has_state:
            raise IOError("a quantized regressor file must be an inference (SGD) regressor")
        re = Regressor(mi, device=device, immutable=immutable)
        expected = sum(re.block_len(b)[0] for b in _block_order(mi))
        got = L.fwhost_regressor_weights_len(r)
        if got != expected:
            raise IOError(f"Lenghts of weights array in regressor file differ: got {got}, expected {expected}")  # sic, regressor.rs:458-462
        want_state = file_has_state and not immutable
        for b in _block_order(mi):
            n, _ = re.block_len(b)
            if b == _lib.BLOCK_LR:
                file_bytes = n * (8 if file_has_state else 4)
                buf = np.empty(file_bytes // 4, dtype=np.float32)
                if L.fwhost_regressor_read(r, buf.ctypes.data_as(C.c_void_p), file_bytes) != 0:
                    raise IOError("truncated regressor file")
                if file_has_state and not want_state:
                    buf = buf.reshape(n, 2)[:, 0].copy()
                re.import_block(b, buf, with_optimizer_state=want_state)
            else:
                w = np.empty(n, dtype=np.float32)
                if quantized and b == _lib.BLOCK_FFM:
                    if L.fwhost_regressor_read_quantized(r, w.ctypes.data_as(C.c_void_p), n) != 0:
                        raise IOError("truncated regressor file")
                elif L.fwhost_regressor_read(r, w.ctypes.data_as(C.c_void_p), n * 4) != 0:
                    raise IOError("truncated regressor file")
                if want_state:
                    acc = np.empty(n, dtype=np.float32)
                    if L.fwhost_regressor_read(r, acc.ctypes.data_as(C.c_void_p), n * 4) != 0:
                        raise IOError("truncated regressor file")
                    re.import_block(b, np.concatenate([w, acc]), True)
                else:
                    if file_has_state:
                        L.fwhost_regressor_skip(r, n * 4)
                    re.import_block(b, w, False)
        return mi, vw, re
    finally:
        L.fwhost_regressor_close(r)
