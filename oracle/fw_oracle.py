"""ctypes binding of the CPU oracle (oracle/fw_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under fwumious_wabbit_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfworacle.so")
_SO_NATIVE = os.path.join(_HERE, "_build", "libfworacle_native.so")

OPT_SGD, OPT_ADAGRAD_FLEX, OPT_ADAGRAD_LUT = 0, 1, 2
GRAPH_REGRESSOR, GRAPH_FFM_BLOCK_ONLY = 0, 1
NN_INIT_XAVIER, NN_INIT_HU, NN_INIT_ONE, NN_INIT_ZERO = 0, 1, 2, 3
MAX_NN_LAYERS = 8
NO_FEATURES = 0x80000000


def build(force=False):
    src = [os.path.join(_HERE, "fw_oracle.c"), os.path.join(_HERE, "fw_oracle.h")]
    if force or not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in src
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class LrFeat(C.Structure):
    _fields_ = [("hash", C.c_uint32), ("value", C.c_float), ("combo_index", C.c_uint32)]


class FfmFeat(C.Structure):
    _fields_ = [("hash", C.c_uint32), ("value", C.c_float), ("contra_field_index", C.c_uint32)]


class FeatureBuffer(C.Structure):
    _fields_ = [
        ("label", C.c_float),
        ("example_importance", C.c_float),
        ("example_number", C.c_uint64),
        ("n_lr", C.c_uint32),
        ("lr", C.POINTER(LrFeat)),
        ("n_ffm", C.c_uint32),
        ("ffm", C.POINTER(FfmFeat)),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("learning_rate", C.c_float), ("power_t", C.c_float), ("init_acc_gradient", C.c_float),
        ("ffm_learning_rate", C.c_float), ("ffm_power_t", C.c_float), ("ffm_init_acc_gradient", C.c_float),
        ("nn_learning_rate", C.c_float), ("nn_power_t", C.c_float), ("nn_init_acc_gradient", C.c_float),
        ("bit_precision", C.c_uint32), ("ffm_bit_precision", C.c_uint32), ("ffm_k", C.c_uint32),
        ("ffm_num_fields", C.c_uint32), ("num_combos", C.c_uint32), ("optimizer", C.c_uint32),
        ("graph", C.c_uint32),
        ("ffm_init_width", C.c_float), ("ffm_init_zero_band", C.c_float), ("ffm_init_center", C.c_float),
        ("nn_num_layers", C.c_uint32),
        ("nn_width", C.c_uint32 * MAX_NN_LAYERS), ("nn_relu", C.c_uint32 * MAX_NN_LAYERS),
        ("nn_init", C.c_uint32 * MAX_NN_LAYERS), ("nn_maxnorm", C.c_float * MAX_NN_LAYERS),
    ]


class TranslateSpec(C.Structure):
    _fields_ = [
        ("n_namespaces", C.c_uint32), ("ns_is_f32", C.POINTER(C.c_uint8)),
        ("n_combos", C.c_uint32), ("combo_off", C.POINTER(C.c_uint32)), ("combo_ns", C.POINTER(C.c_uint32)),
        ("combo_weight", C.POINTER(C.c_float)), ("add_constant", C.c_uint32),
        ("n_fields", C.c_uint32), ("field_off", C.POINTER(C.c_uint32)), ("field_ns", C.POINTER(C.c_uint32)),
        ("bit_precision", C.c_uint32), ("ffm_bit_precision", C.c_uint32), ("ffm_k", C.c_uint32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_examples", C.c_uint32), ("labels", C.POINTER(C.c_float)), ("importance", C.POINTER(C.c_float)),
        ("lr_off", C.POINTER(C.c_uint32)), ("lr_hash", C.POINTER(C.c_uint32)), ("lr_val", C.POINTER(C.c_float)),
        ("lr_combo", C.POINTER(C.c_uint32)),
        ("ffm_off", C.POINTER(C.c_uint32)), ("ffm_hash", C.POINTER(C.c_uint32)), ("ffm_val", C.POINTER(C.c_float)),
        ("ffm_field", C.POINTER(C.c_uint32)),
    ]


_lib = None
_use_native = False


def use_native_build():
    """bench.py's CPU legs: load a copy compiled -march=native ON THIS MACHINE (same source, same -ffp-contract=off) instead
    of the portable x86-64-v3 build.  Must be called before the first lib(); falls back to the portable build when there
    is no compiler on the box.  Returns the flags string that was used."""
    global _use_native
    if _lib is not None:
        return "x86-64-v3 (library already loaded)"
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "native"])  # always rebuilt: a copy made on another machine may not run here
        _use_native = os.path.exists(_SO_NATIVE)
    except Exception:
        _use_native = False
    return "-O3 -march=native" if _use_native else "-O3 -march=x86-64-v3 (no compiler on this box)"


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not _use_native:
        build()
    L = C.CDLL(_SO_NATIVE if _use_native else _SO)
    L.fwo_murmur3_32.restype = C.c_uint32
    L.fwo_murmur3_32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    L.fwo_lut_build.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.fwo_opt_update.restype = C.c_float
    L.fwo_opt_update.argtypes = [C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]
    L.fwo_merand48.restype = C.c_float
    L.fwo_merand48.argtypes = [C.c_uint64]
    L.fwo_parse_line.restype = C.c_int
    L.fwo_parse_line.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_char_p,
                                 C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t, C.c_char_p, C.c_size_t]
    L.fwo_translate.restype = C.c_int
    L.fwo_translate.argtypes = [C.POINTER(TranslateSpec), C.POINTER(C.c_uint32), C.POINTER(LrFeat), C.c_uint32,
                                C.POINTER(C.c_uint32), C.POINTER(FfmFeat), C.c_uint32, C.POINTER(C.c_uint32),
                                C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.fwo_lr_hash_mask.restype = C.c_uint32
    L.fwo_lr_hash_mask.argtypes = [C.c_uint32]
    L.fwo_ffm_hash_mask.restype = C.c_uint32
    L.fwo_ffm_hash_mask.argtypes = [C.c_uint32, C.c_uint32]
    L.fwo_regressor_new.restype = C.c_void_p
    L.fwo_regressor_new.argtypes = [C.POINTER(ModelDesc)]
    L.fwo_regressor_free.argtypes = [C.c_void_p]
    for name in ("fwo_learn", "fwo_forward_backward"):
        f = getattr(L, name)
        f.restype = C.c_float
        f.argtypes = [C.c_void_p, C.POINTER(FeatureBuffer), C.c_int]
    L.fwo_predict.restype = C.c_float
    L.fwo_predict.argtypes = [C.c_void_p, C.POINTER(FeatureBuffer)]
    for name in ("fwo_lr_len", "fwo_ffm_len", "fwo_nn_layer_count"):
        f = getattr(L, name)
        f.restype = C.c_uint32
        f.argtypes = [C.c_void_p]
    L.fwo_nn_layer_len.restype = C.c_uint32
    L.fwo_nn_layer_len.argtypes = [C.c_void_p, C.c_uint32]
    for name in ("fwo_lr_table", "fwo_ffm_weights", "fwo_ffm_acc"):
        f = getattr(L, name)
        f.restype = C.POINTER(C.c_float)
        f.argtypes = [C.c_void_p]
    for name in ("fwo_nn_weights", "fwo_nn_acc"):
        f = getattr(L, name)
        f.restype = C.POINTER(C.c_float)
        f.argtypes = [C.c_void_p, C.c_uint32]
    L.fwo_lut.restype = C.POINTER(C.c_float)
    L.fwo_lut.argtypes = [C.c_void_p, C.c_int]
    L.fwo_test_neuron_layer.restype = C.c_int
    L.fwo_test_neuron_layer.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.fwo_learn_batch_sequential.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(C.c_float), C.c_int]
    L.fwo_hogwild_run.restype = C.c_double
    L.fwo_hogwild_run.argtypes = [C.c_void_p, C.POINTER(TranslateSpec), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                  C.c_uint64, C.c_uint32, C.POINTER(C.c_float)]
    L.fwo_learn_records_wave.argtypes = [C.c_void_p, C.POINTER(TranslateSpec), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                         C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_float)]
    L.fwo_learn_records_wave.restype = None
    L.fwo_learn_records_head_wave.argtypes = [C.c_void_p, C.POINTER(TranslateSpec), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                              C.c_uint64, C.c_uint32, C.POINTER(C.c_float)]
    L.fwo_learn_records_head_wave.restype = None
    _lib = L
    return L


def murmur3_32(data: bytes, seed: int = 0) -> int:
    return lib().fwo_murmur3_32(data, len(data), seed)


def lut_build(lr, power_t, init_acc):
    out = np.zeros(2048, dtype=np.float32)
    lib().fwo_lut_build(lr, power_t, init_acc, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def neuron_layer_test(optimizer, lr, power_t, init_acc, n_in, n_out, init, relu, x, d_out, n_steps):
    """One neuron layer (+ relu) in isolation (block_neural.rs:507-581, block_relu.rs:156-173): returns
    (outs[n_steps, n_out], d_in[n_in] of the last step)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    d_out = np.ascontiguousarray(d_out, dtype=np.float32)
    outs = np.zeros((n_steps, n_out), np.float32)
    d_in = np.zeros(n_in, np.float32)
    lib().fwo_test_neuron_layer(optimizer, lr, power_t, init_acc, n_in, n_out, init, 1 if relu else 0, x.ctypes.data, d_out.ctypes.data,
                                n_steps, outs.ctypes.data, d_in.ctypes.data)
    return outs, d_in


def merand48(seed: int) -> float:
    return lib().fwo_merand48(seed)


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_desc(**kw):
    """ModelInstance::new_empty() defaults (model_instance.rs:120-150) + overrides."""
    d = ModelDesc()
    d.learning_rate, d.power_t, d.init_acc_gradient = 0.5, 0.5, 1.0
    d.ffm_learning_rate, d.ffm_power_t, d.ffm_init_acc_gradient = 0.5, 0.5, 0.0
    d.nn_learning_rate, d.nn_power_t, d.nn_init_acc_gradient = 0.02, 0.45, 0.0
    d.bit_precision, d.ffm_bit_precision, d.ffm_k, d.ffm_num_fields = 18, 18, 0, 0
    d.num_combos = 1  # no combos + constant
    d.optimizer = OPT_SGD
    d.graph = GRAPH_REGRESSOR
    nn_layers = kw.pop("nn_layers", None)
    for k, v in kw.items():
        if not hasattr(d, k):
            raise AttributeError(k)
        setattr(d, k, v)
    if nn_layers:
        d.nn_num_layers = len(nn_layers)
        for i, layer in enumerate(nn_layers):
            d.nn_width[i] = int(layer.get("width", 20))
            d.nn_relu[i] = 1 if layer.get("activation", "none") == "relu" else 0
            d.nn_init[i] = {"xavier": 0, "hu": 1, "one": 2, "zero": 3}[layer.get("init", "hu")]
            d.nn_maxnorm[i] = float(layer.get("maxnorm", 0.0))
    return d


class Spec:
    """Python owner of a fwo_translate_spec (keeps the numpy arrays alive)."""

    def __init__(self, n_namespaces, combos, combo_weights=None, add_constant=True, fields=(), bit_precision=18,
                 ffm_bit_precision=18, ffm_k=0, ns_is_f32=None):
        self.n_namespaces = n_namespaces
        self.combos = [list(c) for c in combos]
        self.fields = [list(f) for f in fields]
        self.combo_off = np.zeros(len(self.combos) + 1, dtype=np.uint32)
        self.combo_off[1:] = np.cumsum([len(c) for c in self.combos])
        self.combo_ns = np.array([n for c in self.combos for n in c] + [0], dtype=np.uint32)
        self.combo_weight = np.array(list(combo_weights) if combo_weights is not None else [1.0] * len(self.combos),
                                     dtype=np.float32)
        if self.combo_weight.size == 0:
            self.combo_weight = np.zeros(1, dtype=np.float32)
        self.field_off = np.zeros(len(self.fields) + 1, dtype=np.uint32)
        self.field_off[1:] = np.cumsum([len(f) for f in self.fields])
        self.field_ns = np.array([n for f in self.fields for n in f] + [0], dtype=np.uint32)
        self.ns_is_f32 = np.array(ns_is_f32 if ns_is_f32 is not None else [0] * n_namespaces, dtype=np.uint8)
        s = TranslateSpec()
        s.n_namespaces = n_namespaces
        s.ns_is_f32 = self.ns_is_f32.ctypes.data_as(C.POINTER(C.c_uint8))
        s.n_combos = len(self.combos)
        s.combo_off, s.combo_ns, s.combo_weight = _u32p(self.combo_off), _u32p(self.combo_ns), _f32p(self.combo_weight)
        s.add_constant = 1 if add_constant else 0
        s.n_fields = len(self.fields)
        s.field_off, s.field_ns = _u32p(self.field_off), _u32p(self.field_ns)
        s.bit_precision, s.ffm_bit_precision, s.ffm_k = bit_precision, ffm_bit_precision, ffm_k
        self.c = s
        self.add_constant = bool(add_constant)
        self.bit_precision, self.ffm_bit_precision, self.ffm_k = bit_precision, ffm_bit_precision, ffm_k

    def translate(self, record):
        rec = np.ascontiguousarray(record, dtype=np.uint32)
        cap = 8192
        lr = (LrFeat * cap)()
        ffm = (FfmFeat * cap)()
        n_lr, n_ffm = C.c_uint32(0), C.c_uint32(0)
        label, imp = C.c_float(0), C.c_float(0)
        rc = lib().fwo_translate(C.byref(self.c), _u32p(rec), lr, cap, C.byref(n_lr), ffm, cap, C.byref(n_ffm),
                                 C.byref(label), C.byref(imp))
        if rc != 0:
            raise RuntimeError("fwo_translate: capacity exceeded")
        lr_l = [(lr[i].hash, lr[i].value, lr[i].combo_index) for i in range(n_lr.value)]
        ffm_l = [(ffm[i].hash, ffm[i].value, ffm[i].contra_field_index) for i in range(n_ffm.value)]
        return label.value, imp.value, lr_l, ffm_l


class Parser:
    """VowpalParser (parser.rs:77-108): namespace names by index."""

    def __init__(self, ns_names, ns_is_f32=None, namespace_skip_prefix=0):
        self.names = [n.encode() if isinstance(n, str) else n for n in ns_names]
        self.arr = (C.c_char_p * len(self.names))(*self.names)
        self.is_f32 = np.array(ns_is_f32 if ns_is_f32 is not None else [0] * len(self.names), dtype=np.uint8)
        self.skip = namespace_skip_prefix

    def parse(self, line):
        if isinstance(line, str):
            line = line.encode()
        out = np.zeros(4096, dtype=np.uint32)
        err = C.create_string_buffer(512)
        n = lib().fwo_parse_line(self.arr, self.is_f32.ctypes.data_as(C.POINTER(C.c_uint8)), len(self.names), self.skip,
                                 line, len(line), _u32p(out), out.size, err, 512)
        if n == -2:
            raise FlushCommand()
        if n == -3:
            raise HogwildLoadCommand()
        if n < 0:
            raise ValueError(err.value.decode())
        return out[:n].copy()


class FlushCommand(Exception):
    pass


class HogwildLoadCommand(Exception):
    pass


def feature_buffer(lr=(), ffm=(), label=0.0, importance=1.0, example_number=0):
    """lr: [(hash, value, combo_index)], ffm: [(hash, value, contra_field_index)]"""
    fb = FeatureBuffer()
    fb.label, fb.example_importance, fb.example_number = label, importance, example_number
    lr_arr = (LrFeat * max(1, len(lr)))(*[LrFeat(*t) for t in lr])
    ffm_arr = (FfmFeat * max(1, len(ffm)))(*[FfmFeat(*t) for t in ffm])
    fb.n_lr, fb.lr = len(lr), C.cast(lr_arr, C.POINTER(LrFeat))
    fb.n_ffm, fb.ffm = len(ffm), C.cast(ffm_arr, C.POINTER(FfmFeat))
    fb._keep = (lr_arr, ffm_arr)
    return fb


class Regressor:
    def __init__(self, desc=None, **kw):
        self.desc = desc if desc is not None else make_desc(**kw)
        self.h = lib().fwo_regressor_new(C.byref(self.desc))

    def __del__(self):
        try:
            if self.h:
                lib().fwo_regressor_free(self.h)
                self.h = None
        except Exception:
            pass

    def learn(self, fb, update=True):
        return lib().fwo_learn(self.h, C.byref(fb), 1 if update else 0)

    def predict(self, fb):
        return lib().fwo_predict(self.h, C.byref(fb))

    def forward_backward(self, fb, update=True):
        return lib().fwo_forward_backward(self.h, C.byref(fb), 1 if update else 0)

    # numpy views onto the oracle's tables (no copies)
    @property
    def lr_table(self):
        n = lib().fwo_lr_len(self.h)
        return np.ctypeslib.as_array(lib().fwo_lr_table(self.h), shape=(n, 2))

    @property
    def ffm_weights(self):
        n = lib().fwo_ffm_len(self.h)
        return np.ctypeslib.as_array(lib().fwo_ffm_weights(self.h), shape=(n,))

    @property
    def ffm_acc(self):
        n = lib().fwo_ffm_len(self.h)
        return np.ctypeslib.as_array(lib().fwo_ffm_acc(self.h), shape=(n,))

    @property
    def nn_layer_count(self):
        return lib().fwo_nn_layer_count(self.h)

    def nn_weights(self, layer):
        n = lib().fwo_nn_layer_len(self.h, layer)
        return np.ctypeslib.as_array(lib().fwo_nn_weights(self.h, layer), shape=(n,))

    def nn_acc(self, layer):
        n = lib().fwo_nn_layer_len(self.h, layer)
        return np.ctypeslib.as_array(lib().fwo_nn_acc(self.h, layer), shape=(n,))

    def lut(self, which):
        return np.ctypeslib.as_array(lib().fwo_lut(self.h, which), shape=(2048,)).copy()

    def learn_batch(self, batch, update=True):
        """batch: dict of numpy arrays with the fwgpu_batch layout; sequential semantics."""
        b = Batch()
        n = int(batch["labels"].shape[0])
        b.n_examples = n
        b.labels, b.importance = _f32p(batch["labels"]), _f32p(batch["importance"])
        b.lr_off, b.lr_hash, b.lr_val, b.lr_combo = (_u32p(batch["lr_off"]), _u32p(batch["lr_hash"]),
                                                     _f32p(batch["lr_val"]), _u32p(batch["lr_combo"]))
        b.ffm_off, b.ffm_hash, b.ffm_val, b.ffm_field = (_u32p(batch["ffm_off"]), _u32p(batch["ffm_hash"]),
                                                         _f32p(batch["ffm_val"]), _u32p(batch["ffm_field"]))
        preds = np.zeros(n, dtype=np.float32)
        lib().fwo_learn_batch_sequential(self.h, C.byref(b), _f32p(preds), 1 if update else 0)
        return preds

    def learn_wave(self, spec, records, rec_off, wave, mode):
        """Analysis tool: emulate `wave` examples in flight (see fwo_learn_records_wave)."""
        records = np.ascontiguousarray(records, dtype=np.uint32)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        n = rec_off.shape[0] - 1
        preds = np.zeros(n, dtype=np.float32)
        lib().fwo_learn_records_wave(self.h, C.byref(spec.c), _u32p(records), rec_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     n, wave, mode, _f32p(preds))
        return preds

    def learn_head_wave(self, spec, records, rec_off, wave):
        """Test tool: the device's batched semantics for a model with a dense head (see fwo_learn_records_head_wave)."""
        records = np.ascontiguousarray(records, dtype=np.uint32)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        n = rec_off.shape[0] - 1
        preds = np.zeros(n, dtype=np.float32)
        lib().fwo_learn_records_head_wave(self.h, C.byref(spec.c), _u32p(records), rec_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                          n, wave, _f32p(preds))
        return preds

    def hogwild(self, spec, records, rec_off, n_threads, want_preds=False):
        records = np.ascontiguousarray(records, dtype=np.uint32)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        n = rec_off.shape[0] - 1
        preds = np.zeros(n, dtype=np.float32) if want_preds else None
        secs = lib().fwo_hogwild_run(self.h, C.byref(spec.c), _u32p(records),
                                     rec_off.ctypes.data_as(C.POINTER(C.c_uint64)), n, n_threads,
                                     _f32p(preds) if want_preds else None)
        return secs, preds
