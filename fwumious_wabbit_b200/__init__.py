"""fwumious_wabbit_b200 -- B200-native (sm_100a) LR / FFM learn-predict hot path behind the
reference's operator interface.  The compute lives in libfwgpu.so (csrc/, C ABI in include/fwgpu.h);
this package is the thin host-side mirror used by tests, bench.py and the CLI."""
from . import _lib
from .feature_buffer import CsrBatch, FeatureBuffer, HashAndValue, HashAndValueAndSeq
from .model_instance import ModelInstance, Optimizer
from .regressor import Dataset, Regressor

__all__ = ["ModelInstance", "Optimizer", "Regressor", "Dataset", "FeatureBuffer", "HashAndValue",
           "HashAndValueAndSeq", "CsrBatch", "_lib"]
