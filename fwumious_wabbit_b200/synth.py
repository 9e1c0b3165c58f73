"""Synthetic workloads of BASELINE.json's configs: records in the reference's cache format
(parser.rs:57-74) produced by the C++ generator in synth_src/fwsynth.cpp (libfwsynth.so -- its own
library, so that bench.py's reference arm generates its input without loading libfwgpu.so), plus the
matching ModelInstance (the flags SURVEY.md section 8d lists for each config)."""
import ctypes as C
import string

import numpy as np

import os

from .model_instance import ModelInstance, Optimizer

NS_LETTERS = string.ascii_uppercase + string.ascii_lowercase
_SYNTH_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfwsynth.so")
_synth = None


def _host():
    global _synth
    if _synth is None:
        if not os.path.exists(_SYNTH_PATH):
            raise ImportError(f"{_SYNTH_PATH} is missing: build it with `python -m fwumious_wabbit_b200.build`")
        L = C.CDLL(_SYNTH_PATH)
        L.fwsynth_murmur3_32.restype = C.c_uint32
        L.fwsynth_murmur3_32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
        L.fwsynth_records.restype = C.c_int
        L.fwsynth_records.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_char_p, C.c_void_p, C.c_uint64, C.c_int]
        L.fwsynth_line.restype = C.c_int
        L.fwsynth_line.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_char_p, C.c_void_p, C.c_uint64]
        _synth = L
    return _synth


def murmur3_32(data: bytes, seed: int = 0) -> int:
    return _host().fwsynth_murmur3_32(data, len(data), seed)


class Workload:
    """name, ModelInstance, namespace letters and cardinalities."""

    def __init__(self, name, mi, ns_names, cardinality, description):
        self.name, self.mi, self.ns_names = name, mi, ns_names
        self.cardinality = np.asarray(cardinality, dtype=np.uint32)
        self.description = description

    @property
    def n_namespaces(self):
        return len(self.ns_names)

    @property
    def record_len(self):
        return 3 + self.n_namespaces

    def records(self, n_examples, first=0, seed=1, out=None, threads=0, uniform=False):
        """(n_examples, 3 + N) uint32 array of fixed-width records.  uniform=True draws ids uniformly instead of
        Zipf-like (a diagnostic stream without hot rows)."""
        if uniform:
            seed |= 1 << 63
        if out is None:
            out = np.empty((n_examples, self.record_len), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.flags.c_contiguous and out.size >= n_examples * self.record_len
        rc = _host().fwsynth_records(out.ctypes.data_as(C.c_void_p), n_examples, first, self.n_namespaces,
                                          self.ns_names.encode(), self.cardinality.ctypes.data_as(C.c_void_p), seed, threads)
        if rc != 0:
            raise RuntimeError("fwsynth_records failed")
        return out

    def line(self, i, seed=1) -> str:
        buf = C.create_string_buffer(64 + 24 * self.n_namespaces)
        n = _host().fwsynth_line(buf, len(buf), i, self.n_namespaces, self.ns_names.encode(),
                                      self.cardinality.ctypes.data_as(C.c_void_p), seed)
        if n < 0:
            raise RuntimeError("fwsynth_line failed")
        return buf.value.decode()

    # algorithmic bytes per example, SURVEY.md section 8(d)
    def algorithmic_bytes_per_example(self, train=True):
        mi = self.mi
        F = len(mi.ffm_fields) if mi.ffm_k else 0
        n_lr = mi.num_combos
        inp = 4 * (3 + self.n_namespaces)
        if train and mi.nn_layers:
            # two passes around the head: the rows are read twice (4 B/slot more), the head input X is written and read
            # and its gradient dX written and read (4 floats per input); head weights are per sub-batch, not per example
            x_len = n_lr + F * (F + 1) // 2
            return 20 * F * F * mi.ffm_k + 16 * n_lr + inp + 4 + 16 * x_len
        if train:
            return 16 * F * F * mi.ffm_k + 16 * n_lr + inp + 4
        return 4 * F * F * mi.ffm_k + 4 * n_lr + inp + 4


def _mi(n_ns, *, ffm_k, ffm_bits, bits, interactions=(), lr=0.1, ffm_lr=0.05, power_t=0.5, ffm_init_acc=0.0):
    mi = ModelInstance()
    mi.num_namespaces = n_ns
    mi.feature_combo_descs = [([j], 1.0) for j in range(n_ns)] + [(list(c), 1.0) for c in interactions]
    mi.add_constant_feature = True
    mi.bit_precision = bits
    mi.learning_rate, mi.power_t = lr, power_t
    mi.ffm_k, mi.ffm_bit_precision = ffm_k, ffm_bits
    mi.ffm_fields = [[j] for j in range(n_ns)] if ffm_k else []
    mi.ffm_learning_rate, mi.ffm_power_t = ffm_lr, power_t
    mi.optimizer = Optimizer.AdagradLUT  # --adaptive with fastmath (model_instance.rs:481-492)
    mi.init_acc_gradient, mi.ffm_init_acc_gradient = 1.0, ffm_init_acc
    return mi


def workload(name: str) -> Workload:
    """BASELINE.json configs: c1 (LR only), c2 (FFM k=4, 8 fields), c3 (FFM k=8, 39 fields, Criteo shape),
    c4 (c3 with ffm_bit_precision 28), c5 (c3 + 2x256 ReLU head)."""
    if name == "c1":
        # benchmark/generate.py shape with 6 random namespaces -> A..H; --interactions AB, -b 18, power_t 0
        mi = _mi(8, ffm_k=0, ffm_bits=18, bits=18, interactions=[(0, 1)], lr=0.1, power_t=0.0)
        return Workload("c1", mi, NS_LETTERS[:8], [1000] * 8,
                        "LR-only, 8 namespaces A..H + interaction AB, bit_precision=18, AdagradLUT power_t=0")
    if name == "c2":
        mi = _mi(8, ffm_k=4, ffm_bits=20, bits=18)
        return Workload("c2", mi, NS_LETTERS[:8], [100000] * 8,
                        "FFM k=4, 8 fields (one namespace each, 1e5 Zipf ids), ffm_bit_precision=20, -b 18")
    if name == "c5":
        w = workload("c3")
        # --nn_layers 2 --nn 0:width:256 --nn 0:activation:relu --nn 1:width:256 --nn 1:activation:relu;
        # nn_learning_rate / nn_power_t / nn_init_acc_gradient default to the ffm values (model_instance.rs:418-428)
        w.mi.nn_layers = [{"width": "256", "activation": "relu"}, {"width": "256", "activation": "relu"}]
        w.mi.nn_learning_rate, w.mi.nn_power_t, w.mi.nn_init_acc_gradient = w.mi.ffm_learning_rate, w.mi.ffm_power_t, w.mi.ffm_init_acc_gradient
        w.name = "c5"
        w.description = w.description.replace("FFM k=8", "Deep FFM k=8 + 2x256 ReLU head (topology one)")
        return w
    if name in ("c3", "c4"):
        card = [100] * 13 + [10 ** (3 + (j % 5)) for j in range(26)]  # 13 binned numeric + 26 categorical 1e3..1e7
        # -l 0.05 --ffm_learning_rate 0.02 --ffm_init_acc_gradient 0.1: with 39 fields the reference's defaults
        # (ffm_init_acc_gradient 0) make the sequential learner itself diverge on this stream (logloss above the prior)
        mi = _mi(39, ffm_k=8, ffm_bits=24 if name == "c3" else 28, bits=24, lr=0.05, ffm_lr=0.02, ffm_init_acc=0.1)
        return Workload(name, mi, NS_LETTERS[:39], card,
                        f"FFM k=8, 39 fields (Criteo shape: 13 low-card + 26 high-card), ffm_bit_precision={mi.ffm_bit_precision}, -b 24, "
                        "-l 0.05 --ffm_learning_rate 0.02 --ffm_init_acc_gradient 0.1")
    raise KeyError(name)
