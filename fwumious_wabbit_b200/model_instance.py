"""ModelInstance: the hyper-parameters the hot path reads (reference: model_instance.rs:47-150).

Field names and defaults follow ModelInstance::new_empty (model_instance.rs:120-150).  Namespaces
are identified by their namespace_index (vwmap.rs:22-27); feature_combo_descs is a list of
(namespace index list, weight) and ffm_fields a list of namespace index lists, the same shapes the
reference builds from --keep/--interactions/--linear and --ffm_field (model_instance.rs:296-495).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _lib


class Optimizer:
    SGD = _lib.OPT_SGD
    AdagradFlex = _lib.OPT_ADAGRAD_FLEX
    AdagradLUT = _lib.OPT_ADAGRAD_LUT
    names = {0: "SGD", 1: "AdagradFlex", 2: "AdagradLUT"}


@dataclass
class ModelInstance:
    learning_rate: float = 0.5
    minimum_learning_rate: float = 0.0
    power_t: float = 0.5
    bit_precision: int = 18
    add_constant_feature: bool = True
    feature_combo_descs: List[Tuple[List[int], float]] = field(default_factory=list)
    ffm_fields: List[List[int]] = field(default_factory=list)
    ffm_k: int = 0
    ffm_bit_precision: int = 18
    fastmath: bool = True
    ffm_initialization_type: str = "default"
    ffm_k_threshold: float = 0.0
    ffm_init_center: float = 0.0
    ffm_init_width: float = 0.0
    ffm_init_zero_band: float = 0.0
    ffm_init_acc_gradient: float = 0.0
    init_acc_gradient: float = 1.0
    ffm_learning_rate: float = 0.5
    ffm_power_t: float = 0.5
    nn_init_acc_gradient: float = 0.0
    nn_learning_rate: float = 0.02
    nn_power_t: float = 0.45
    nn_layers: List[dict] = field(default_factory=list)
    nn_topology: str = "one"
    optimizer: int = Optimizer.SGD
    # not in the reference: namespace table the translate spec refers to
    num_namespaces: int = 0
    ns_is_f32: Optional[List[int]] = None
    max_ffm_per_example: int = 0
    max_lr_per_example: int = 0
    hogwild_ramp_div: int = 0  # 0 = default (32), 0xffffffff = no concurrency ramp
    hogwild_max_inflight: int = 0  # 0 = automatic (16 for constant-step models, unlimited otherwise)

    @staticmethod
    def new_empty():
        return ModelInstance()

    @property
    def num_combos(self):
        return len(self.feature_combo_descs) + (1 if self.add_constant_feature else 0)

    def to_desc(self, immutable=False):
        """Flatten into a fwgpu_model_desc; returns (desc, keepalive)."""
        d = _lib.ModelDesc()
        d.learning_rate, d.power_t, d.init_acc_gradient = self.learning_rate, self.power_t, self.init_acc_gradient
        d.ffm_learning_rate, d.ffm_power_t, d.ffm_init_acc_gradient = (self.ffm_learning_rate, self.ffm_power_t,
                                                                      self.ffm_init_acc_gradient)
        d.nn_learning_rate, d.nn_power_t, d.nn_init_acc_gradient = (self.nn_learning_rate, self.nn_power_t,
                                                                   self.nn_init_acc_gradient)
        d.bit_precision, d.ffm_bit_precision, d.ffm_k = self.bit_precision, self.ffm_bit_precision, self.ffm_k
        d.ffm_num_fields = len(self.ffm_fields) if self.ffm_k > 0 else 0
        d.num_combos = self.num_combos
        d.optimizer = self.optimizer
        d.immutable = 1 if immutable else 0
        d.ffm_init_width, d.ffm_init_zero_band, d.ffm_init_center = (self.ffm_init_width, self.ffm_init_zero_band,
                                                                    self.ffm_init_center)
        d.nn_num_layers = len(self.nn_layers)
        for i, layer in enumerate(self.nn_layers[: _lib.MAX_NN_LAYERS]):
            d.nn_width[i] = int(layer.get("width", 20))
            d.nn_relu[i] = 1 if layer.get("activation", "none") == "relu" else 0
            d.nn_init[i] = {"xavier": 0, "hu": 1, "one": 2, "zero": 3}[layer.get("init", "hu")]
            unsupported = {k_: v for k_, v in layer.items() if k_ in ("dropout", "maxnorm") and float(v) != 0.0}
            if layer.get("layernorm", "none") != "none" or unsupported:
                raise ValueError(f"--nn {i}: dropout / maxnorm / layernorm are not implemented by the CUDA head")
        if self.nn_layers and self.nn_topology != "one":
            raise ValueError("only nn topology \"one\" is implemented by the CUDA head")
        nns = self.num_namespaces
        for ns_list, _ in self.feature_combo_descs:
            nns = max(nns, max(ns_list) + 1 if ns_list else 0)
        for f in self.ffm_fields:
            nns = max(nns, max(f) + 1 if f else 0)
        is_f32 = np.zeros(max(nns, 1), dtype=np.uint8)
        if self.ns_is_f32 is not None:
            is_f32[: len(self.ns_is_f32)] = self.ns_is_f32
        combo_off = np.zeros(len(self.feature_combo_descs) + 1, dtype=np.uint32)
        combo_off[1:] = np.cumsum([len(c[0]) for c in self.feature_combo_descs])
        combo_ns = np.array([n for c in self.feature_combo_descs for n in c[0]] + [0], dtype=np.uint32)
        combo_w = np.array([c[1] for c in self.feature_combo_descs] + [0.0], dtype=np.float32)
        fields = self.ffm_fields if self.ffm_k > 0 else []
        field_off = np.zeros(len(fields) + 1, dtype=np.uint32)
        field_off[1:] = np.cumsum([len(f) for f in fields])
        field_ns = np.array([n for f in fields for n in f] + [0], dtype=np.uint32)
        d.n_namespaces = nns
        d.ns_is_f32 = is_f32.ctypes.data_as(_lib.u8p)
        d.n_combos = len(self.feature_combo_descs)
        d.combo_off = combo_off.ctypes.data_as(_lib.u32p)
        d.combo_ns = combo_ns.ctypes.data_as(_lib.u32p)
        d.combo_weight = combo_w.ctypes.data_as(_lib.f32p)
        d.add_constant = 1 if self.add_constant_feature else 0
        d.field_off = field_off.ctypes.data_as(_lib.u32p)
        d.field_ns = field_ns.ctypes.data_as(_lib.u32p)
        d.max_ffm_per_example = self.max_ffm_per_example
        d.max_lr_per_example = self.max_lr_per_example
        d.hogwild_ramp_div = self.hogwild_ramp_div
        d.hogwild_max_inflight = self.hogwild_max_inflight
        return d, (is_f32, combo_off, combo_ns, combo_w, field_off, field_ns)
