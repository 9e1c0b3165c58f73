"""FeatureBuffer / HashAndValue / HashAndValueAndSeq (reference: feature_buffer.rs:10-31) and the
CSR mini-batch the C ABI takes (include/fwgpu.h fwgpu_batch)."""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib


@dataclass
class HashAndValue:
    hash: int
    value: float
    combo_index: int


@dataclass
class HashAndValueAndSeq:
    hash: int
    value: float
    contra_field_index: int  # field index * ffm_k, as in the reference


@dataclass
class FeatureBuffer:
    label: float = 0.0
    example_importance: float = 1.0
    example_number: int = 0
    lr_buffer: List[HashAndValue] = field(default_factory=list)
    ffm_buffer: List[HashAndValueAndSeq] = field(default_factory=list)


class CsrBatch:
    """Owns the numpy arrays of one fwgpu_batch."""

    def __init__(self, labels, importance, lr_off, lr_hash, lr_val, lr_combo, ffm_off, ffm_hash, ffm_val, ffm_field):
        a = np.ascontiguousarray
        self.labels, self.importance = a(labels, np.float32), a(importance, np.float32)
        self.lr_off, self.lr_hash = a(lr_off, np.uint32), a(lr_hash, np.uint32)
        self.lr_val, self.lr_combo = a(lr_val, np.float32), a(lr_combo, np.uint32)
        self.ffm_off, self.ffm_hash = a(ffm_off, np.uint32), a(ffm_hash, np.uint32)
        self.ffm_val, self.ffm_field = a(ffm_val, np.float32), a(ffm_field, np.uint32)
        self.n = int(self.labels.shape[0])
        assert self.lr_off.shape[0] == self.n + 1 and self.ffm_off.shape[0] == self.n + 1

    @staticmethod
    def from_feature_buffers(fbs, ffm_k):
        n = len(fbs)
        labels = np.array([fb.label for fb in fbs], dtype=np.float32)
        imp = np.array([fb.example_importance for fb in fbs], dtype=np.float32)
        lr_off = np.zeros(n + 1, dtype=np.uint32)
        ffm_off = np.zeros(n + 1, dtype=np.uint32)
        lr_off[1:] = np.cumsum([len(fb.lr_buffer) for fb in fbs])
        ffm_off[1:] = np.cumsum([len(fb.ffm_buffer) for fb in fbs])
        lr = [e for fb in fbs for e in fb.lr_buffer]
        ffm = [e for fb in fbs for e in fb.ffm_buffer]
        k = max(ffm_k, 1)
        return CsrBatch(
            labels, imp, lr_off,
            np.array([e.hash for e in lr], dtype=np.uint32), np.array([e.value for e in lr], dtype=np.float32),
            np.array([e.combo_index for e in lr], dtype=np.uint32), ffm_off,
            np.array([e.hash for e in ffm], dtype=np.uint32), np.array([e.value for e in ffm], dtype=np.float32),
            np.array([e.contra_field_index // k for e in ffm], dtype=np.uint32),
        )

    def as_dict(self):
        return dict(labels=self.labels, importance=self.importance, lr_off=self.lr_off, lr_hash=self.lr_hash,
                    lr_val=self.lr_val, lr_combo=self.lr_combo, ffm_off=self.ffm_off, ffm_hash=self.ffm_hash,
                    ffm_val=self.ffm_val, ffm_field=self.ffm_field)

    def slice(self, a, b):
        l0, l1 = int(self.lr_off[a]), int(self.lr_off[b])
        f0, f1 = int(self.ffm_off[a]), int(self.ffm_off[b])
        return CsrBatch(self.labels[a:b], self.importance[a:b], self.lr_off[a:b + 1] - l0, self.lr_hash[l0:l1],
                        self.lr_val[l0:l1], self.lr_combo[l0:l1], self.ffm_off[a:b + 1] - f0, self.ffm_hash[f0:f1],
                        self.ffm_val[f0:f1], self.ffm_field[f0:f1])

    def c_struct(self):
        b = _lib.Batch()
        b.n_examples = self.n
        p = lambda arr, t: arr.ctypes.data_as(t)
        b.labels, b.importance = p(self.labels, _lib.f32p), p(self.importance, _lib.f32p)
        b.lr_off, b.lr_hash = p(self.lr_off, _lib.u32p), p(self.lr_hash, _lib.u32p)
        b.lr_val, b.lr_combo = p(self.lr_val, _lib.f32p), p(self.lr_combo, _lib.u32p)
        b.ffm_off, b.ffm_hash = p(self.ffm_off, _lib.u32p), p(self.ffm_hash, _lib.u32p)
        b.ffm_val, b.ffm_field = p(self.ffm_val, _lib.f32p), p(self.ffm_field, _lib.u32p)
        return b
