"""Regressor: the reference's operator interface for the hot path (regressor.rs:142-534) on top of
the C ABI.  learn()/predict() take one FeatureBuffer like Regressor::learn/predict; the *_batch
and *_records calls are the mini-batch forms the GPU is built for."""
import ctypes as C

import numpy as np

from . import _lib
from .feature_buffer import CsrBatch, FeatureBuffer
from .model_instance import ModelInstance, Optimizer


def _vp(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


class Dataset:
    def __init__(self, reg, handle, n_examples):
        self.reg, self.handle, self.n_examples = reg, handle, n_examples

    def free(self):
        if self.handle:
            _lib.lib().fwgpu_dataset_free(self.reg.h, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Regressor:
    def __init__(self, mi: ModelInstance, device: int = 0, immutable: bool = False, shard=None):
        """shard = (rank, world, rendezvous_prefix): one model whose tables are hash-range-sharded over `world` GPUs
        (one process per GPU, collective constructor; fwgpu_create_sharded)."""
        self.L = _lib.lib()
        self.mi = mi
        self.immutable = immutable
        desc, keep = mi.to_desc(immutable=immutable)
        h = C.c_void_p()
        if shard is not None:
            rank, world, prefix = shard
            st = self.L.fwgpu_create_sharded(C.byref(desc), device, rank, world, str(prefix).encode(), 0, C.byref(h))
        else:
            st = self.L.fwgpu_create(C.byref(desc), device, C.byref(h))
        if st != 0:
            raise _lib.FwgpuError(st, self.L.fwgpu_last_error(None).decode())
        self.h = h
        self.device = device

    # regressor.rs:343-345
    def get_name(self):
        opt = Optimizer.SGD if self.immutable else self.mi.optimizer
        return f'Regressor with optimizer "{Optimizer.names[opt]}"'

    def close(self):
        if getattr(self, "h", None):
            self.L.fwgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != 0:
            raise _lib.FwgpuError(st, self.L.fwgpu_last_error(self.h).decode())

    def sync(self):
        self._check(self.L.fwgpu_sync(self.h))

    def shard_barrier(self):
        self._check(self.L.fwgpu_shard_barrier(self.h))

    def shard_info(self):
        r, w, f, n = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.fwgpu_shard_info(self.h, C.byref(r), C.byref(w), C.byref(f), C.byref(n)))
        return r.value, w.value, f.value, n.value

    # ---- Regressor::learn / predict (regressor.rs:356-395), one example ----
    def learn(self, fb: FeatureBuffer, update: bool = True) -> float:
        if update and self.immutable:
            # regressor.rs:362-365 panics; here the C ABI reports FWGPU_ERR_IMMUTABLE
            pass
        return float(self.learn_batch(CsrBatch.from_feature_buffers([fb], self.mi.ffm_k), update)[0])

    def predict(self, fb: FeatureBuffer) -> float:
        return float(self.predict_batch(CsrBatch.from_feature_buffers([fb], self.mi.ffm_k))[0])

    # ---- mini-batch forms ----
    def learn_batch(self, batch: CsrBatch, update: bool = True, out=None, sync=True):
        preds = out if out is not None else np.empty(batch.n, dtype=np.float32)
        b = batch.c_struct()
        self._check(self.L.fwgpu_learn_batch(self.h, C.byref(b), _vp(preds), 1 if update else 0))
        if sync:
            self.sync()
        return preds

    def predict_batch(self, batch: CsrBatch, out=None, sync=True):
        preds = out if out is not None else np.empty(batch.n, dtype=np.float32)
        b = batch.c_struct()
        self._check(self.L.fwgpu_predict_batch(self.h, C.byref(b), _vp(preds)))
        if sync:
            self.sync()
        return preds

    def learn_records(self, records, rec_off=None, n_examples=None, update=True, out=None, want_preds=True, sync=True):
        """translate + learn on raw parser/cache records (main.rs:240-256)."""
        records = np.ascontiguousarray(records, dtype=np.uint32) if not isinstance(records, np.ndarray) else records
        assert records.dtype == np.uint32 and records.flags.c_contiguous
        if rec_off is not None:
            rec_off = np.ascontiguousarray(rec_off, dtype=np.uint32)
            n = rec_off.shape[0] - 1
        else:
            n = int(n_examples)
        preds = out if out is not None else (np.empty(n, dtype=np.float32) if want_preds else None)
        self._check(self.L.fwgpu_learn_records(self.h, _vp(records), records.size, _vp(rec_off), n, _vp(preds),
                                               1 if update else 0))
        if sync:
            self.sync()
        return preds

    def translate_records(self, records, rec_off=None, n_examples=None, lr_cap=None, ffm_cap=None):
        records = np.ascontiguousarray(records, dtype=np.uint32)
        if rec_off is not None:
            rec_off = np.ascontiguousarray(rec_off, dtype=np.uint32)
            n = rec_off.shape[0] - 1
        else:
            n = int(n_examples)
        lr_cap = lr_cap or max(1, 4 * records.size + n)
        ffm_cap = ffm_cap or max(1, 4 * records.size + n)
        labels, imp = np.empty(n, np.float32), np.empty(n, np.float32)
        lr_off, ffm_off = np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32)
        lr_hash, lr_val, lr_combo = np.empty(lr_cap, np.uint32), np.empty(lr_cap, np.float32), np.empty(lr_cap, np.uint32)
        ffm_hash, ffm_val, ffm_field = np.empty(ffm_cap, np.uint32), np.empty(ffm_cap, np.float32), np.empty(ffm_cap, np.uint32)
        self._check(self.L.fwgpu_translate_records(self.h, _vp(records), records.size, _vp(rec_off), n, _vp(labels), _vp(imp),
                                                   _vp(lr_off), _vp(lr_hash), _vp(lr_val), _vp(lr_combo), lr_cap,
                                                   _vp(ffm_off), _vp(ffm_hash), _vp(ffm_val), _vp(ffm_field), ffm_cap))
        nl, nf = int(lr_off[n]), int(ffm_off[n])
        return CsrBatch(labels, imp, lr_off, lr_hash[:nl], lr_val[:nl], lr_combo[:nl], ffm_off, ffm_hash[:nf],
                        ffm_val[:nf], ffm_field[:nf])

    # ---- records resident in HBM ----
    def upload_dataset(self, records, rec_off=None, n_examples=None):
        records = np.ascontiguousarray(records, dtype=np.uint32)
        if rec_off is not None:
            rec_off = np.ascontiguousarray(rec_off, dtype=np.uint32)
            n = rec_off.shape[0] - 1
        else:
            n = int(n_examples)
        h = C.c_void_p()
        self._check(self.L.fwgpu_dataset_upload(self.h, _vp(records), records.size, _vp(rec_off), n, C.byref(h)))
        return Dataset(self, h, n)

    def learn_dataset(self, ds: Dataset, first=0, count=None, update=True, out=None, sync=True):
        count = ds.n_examples - first if count is None else count
        self._check(self.L.fwgpu_dataset_learn(self.h, ds.handle, first, count, _vp(out), 1 if update else 0))
        if sync:
            self.sync()
        return out

    # ---- weights (regressor.rs:426-469) ----
    def block_len(self, block):
        n, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.fwgpu_block_len(self.h, block, C.byref(n), C.byref(b)))
        return n.value, b.value

    def export_block(self, block) -> np.ndarray:
        n, nbytes = self.block_len(block)
        out = np.empty(nbytes // 4, dtype=np.float32)
        if nbytes:
            self._check(self.L.fwgpu_export_block(self.h, block, _vp(out), nbytes))
        return out

    def import_block(self, block, payload, with_optimizer_state=True):
        payload = np.ascontiguousarray(payload, dtype=np.float32)
        self._check(self.L.fwgpu_import_block(self.h, block, _vp(payload), payload.nbytes, 1 if with_optimizer_state else 0))

    def lut(self, which):
        out = np.empty(_lib.LUT_SIZE, dtype=np.float32)
        self._check(self.L.fwgpu_get_lut(self.h, which, _vp(out)))
        return out

    # convenience views used by tests
    def get_lr_table(self):
        """(len, 2) array of {w, acc}; SGD/immutable ctxs return (len, 1)."""
        n, nbytes = self.block_len(_lib.BLOCK_LR)
        return self.export_block(_lib.BLOCK_LR).reshape(n, nbytes // (4 * n))

    def get_ffm(self):
        n, nbytes = self.block_len(_lib.BLOCK_FFM)
        raw = self.export_block(_lib.BLOCK_FFM)
        if nbytes == n * 8:
            return raw[:n], raw[n:]
        return raw[:n], None

    def set_ffm(self, w, acc=None):
        if acc is None:
            self.import_block(_lib.BLOCK_FFM, w, with_optimizer_state=False)
        else:
            self.import_block(_lib.BLOCK_FFM, np.concatenate([w, acc]).astype(np.float32), True)

    # dense head: layer l < len(mi.nn_layers) is hidden layer l, l == len(mi.nn_layers) the final neuron
    def nn_layer_count(self):
        return len(self.mi.nn_layers) + 1 if self.mi.nn_layers else 0

    def get_nn(self, layer):
        """(weights, accumulators or None) of a head layer: (n_in + 1) * n_out floats, biases last (block_neural.rs:83-86)."""
        n, nbytes = self.block_len(_lib.BLOCK_NN0 + layer)
        raw = self.export_block(_lib.BLOCK_NN0 + layer)
        return (raw[:n], raw[n:]) if nbytes == n * 8 else (raw[:n], None)

    def set_nn(self, layer, w, acc=None):
        if acc is None:
            self.import_block(_lib.BLOCK_NN0 + layer, w, with_optimizer_state=False)
        else:
            self.import_block(_lib.BLOCK_NN0 + layer, np.concatenate([w, acc]).astype(np.float32), True)

    def set_lr_table(self, table):
        table = np.ascontiguousarray(table, dtype=np.float32)
        self.import_block(_lib.BLOCK_LR, table.reshape(-1), with_optimizer_state=(table.ndim == 2 and table.shape[1] == 2))

    def debug_logistic(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._check(self.L.fwgpu_debug_logistic(self.h, _vp(x), _vp(out), x.size))
        return out

    def set_examples_seen(self, n):
        self._check(self.L.fwgpu_set_examples_seen(self.h, n))

    def examples_seen(self):
        return int(self.L.fwgpu_get_examples_seen(self.h))

    # ---- measurement ----
    def set_profiling(self, on=True):
        self._check(self.L.fwgpu_set_profiling(self.h, 1 if on else 0))

    def kernel_time(self, kind=0):
        ms, n = C.c_double(0), C.c_uint64(0)
        self._check(self.L.fwgpu_kernel_time(self.h, kind, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def path_counts(self):
        """{fixed, fixed_cta, general: launches per learn-kernel family; general_examples: examples the general kernel handled}."""
        out = (C.c_uint64 * 4)()
        self._check(self.L.fwgpu_debug_path_counts(self.h, out))
        return {"fixed": int(out[0]), "fixed_cta": int(out[1]), "general": int(out[2]), "general_examples": int(out[3])}

    def launch_count(self):
        return int(self.L.fwgpu_launch_count(self.h))

    def stream_ptr(self):
        return int(self.L.fwgpu_stream(self.h) or 0)
