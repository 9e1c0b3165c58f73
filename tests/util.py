"""Shared helpers: build the oracle regressor / translate spec from the product's ModelInstance,
random CSR batches, logloss."""
import numpy as np

from oracle import fw_oracle as fo


def oracle_regressor(mi, graph=fo.GRAPH_REGRESSOR):
    return fo.Regressor(
        learning_rate=mi.learning_rate, power_t=mi.power_t, init_acc_gradient=mi.init_acc_gradient,
        ffm_learning_rate=mi.ffm_learning_rate, ffm_power_t=mi.ffm_power_t, ffm_init_acc_gradient=mi.ffm_init_acc_gradient,
        nn_learning_rate=mi.nn_learning_rate, nn_power_t=mi.nn_power_t, nn_init_acc_gradient=mi.nn_init_acc_gradient,
        bit_precision=mi.bit_precision, ffm_bit_precision=mi.ffm_bit_precision, ffm_k=mi.ffm_k,
        ffm_num_fields=len(mi.ffm_fields) if mi.ffm_k else 0, num_combos=mi.num_combos, optimizer=mi.optimizer,
        graph=graph, ffm_init_width=mi.ffm_init_width, ffm_init_zero_band=mi.ffm_init_zero_band,
        ffm_init_center=mi.ffm_init_center, nn_layers=mi.nn_layers or None)


def oracle_spec(mi, n_namespaces=None):
    nns = n_namespaces or mi.num_namespaces
    for ns_list, _ in mi.feature_combo_descs:
        nns = max(nns, max(ns_list) + 1)
    for f in mi.ffm_fields:
        nns = max(nns, (max(f) + 1) if f else 0)
    return fo.Spec(nns, [c[0] for c in mi.feature_combo_descs], [c[1] for c in mi.feature_combo_descs],
                   add_constant=mi.add_constant_feature, fields=mi.ffm_fields if mi.ffm_k else [],
                   bit_precision=mi.bit_precision, ffm_bit_precision=mi.ffm_bit_precision, ffm_k=mi.ffm_k,
                   ns_is_f32=mi.ns_is_f32)


def oracle_translate_batch(spec, records, rec_off=None, fixed_len=None):
    """Translate a batch with the oracle; returns the dict layout of fwgpu_batch."""
    recs = np.ascontiguousarray(records, dtype=np.uint32).reshape(-1)
    if rec_off is None:
        n = recs.size // fixed_len
        rec_off = np.arange(n + 1, dtype=np.uint64) * fixed_len
    n = len(rec_off) - 1
    labels, imp = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lr_off, ffm_off = np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32)
    lr_h, lr_v, lr_c, f_h, f_v, f_f = [], [], [], [], [], []
    k = max(spec.ffm_k, 1)
    for i in range(n):
        lab, im, lr, ffm = spec.translate(recs[int(rec_off[i]):int(rec_off[i + 1])])
        labels[i], imp[i] = lab, im
        for h, v, c in lr:
            lr_h.append(h); lr_v.append(v); lr_c.append(c)
        for h, v, cf in ffm:
            f_h.append(h); f_v.append(v); f_f.append(cf // k)
        lr_off[i + 1], ffm_off[i + 1] = len(lr_h), len(f_h)
    return dict(labels=labels, importance=imp, lr_off=lr_off, lr_hash=np.array(lr_h, np.uint32),
                lr_val=np.array(lr_v, np.float32), lr_combo=np.array(lr_c, np.uint32), ffm_off=ffm_off,
                ffm_hash=np.array(f_h, np.uint32), ffm_val=np.array(f_v, np.float32), ffm_field=np.array(f_f, np.uint32))


def logloss(preds, labels):
    p = np.clip(np.asarray(preds, dtype=np.float64), 1e-7, 1 - 1e-7)
    y = np.asarray(labels, dtype=np.float64)
    return float(-np.mean(y * np.log(p) + (1 - y) * np.log(1 - p)))


def random_csr(rng, n, mi, multi_valued=False, empty_prob=0.0, value_one=True):
    """Random translated examples directly in CSR form (hashes already masked)."""
    F = len(mi.ffm_fields) if mi.ffm_k else 0
    lr_mask = (1 << mi.bit_precision) - 1
    kp = 1
    while kp < max(mi.ffm_k, 1):
        kp <<= 1
    ffm_mask = (((1 << mi.ffm_bit_precision) - 1) ^ (kp - 1)) if mi.ffm_k else 0
    labels = (rng.random(n) < 0.4).astype(np.float32)
    imp = np.ones(n, np.float32)
    lr_off, ffm_off = [0], [0]
    lr_h, lr_v, lr_c, f_h, f_v, f_f = [], [], [], [], [], []
    ncomb = len(mi.feature_combo_descs)
    for _ in range(n):
        for c in range(ncomb):
            if rng.random() < empty_prob:
                continue
            lr_h.append(int(rng.integers(0, 1 << 31)) & lr_mask)
            lr_v.append(1.0 if value_one else float(np.float32(rng.uniform(0.5, 2.0))))
            lr_c.append(c)
        if mi.add_constant_feature:
            lr_h.append(11650396 & lr_mask); lr_v.append(1.0); lr_c.append(ncomb)
        for f in range(F):
            if rng.random() < empty_prob:
                continue
            m = int(rng.integers(1, 4)) if multi_valued else 1
            for _j in range(m):
                f_h.append(int(rng.integers(0, 1 << 31)) & ffm_mask)
                f_v.append(1.0 if value_one else float(np.float32(rng.uniform(0.5, 2.0))))
                f_f.append(f)
        lr_off.append(len(lr_h)); ffm_off.append(len(f_h))
    return dict(labels=labels, importance=imp, lr_off=np.array(lr_off, np.uint32), lr_hash=np.array(lr_h, np.uint32),
                lr_val=np.array(lr_v, np.float32), lr_combo=np.array(lr_c, np.uint32), ffm_off=np.array(ffm_off, np.uint32),
                ffm_hash=np.array(f_h, np.uint32), ffm_val=np.array(f_v, np.float32), ffm_field=np.array(f_f, np.uint32))


def csr_from_dict(d):
    from fwumious_wabbit_b200 import CsrBatch

    return CsrBatch(d["labels"], d["importance"], d["lr_off"], d["lr_hash"], d["lr_val"], d["lr_combo"], d["ffm_off"],
                    d["ffm_hash"], d["ffm_val"], d["ffm_field"])


def sync_tables_from_oracle(gpu_reg, ora):
    """Load the oracle's current tables into the GPU regressor (identical weights on both sides)."""
    from fwumious_wabbit_b200 import _lib

    n, nbytes = gpu_reg.block_len(_lib.BLOCK_LR)
    t = ora.lr_table
    gpu_reg.import_block(_lib.BLOCK_LR, t.reshape(-1) if nbytes == n * 8 else t[:, 0].copy(), nbytes == n * 8)
    n, nbytes = gpu_reg.block_len(_lib.BLOCK_FFM)
    if n:
        if nbytes == n * 8:
            gpu_reg.import_block(_lib.BLOCK_FFM, np.concatenate([ora.ffm_weights, ora.ffm_acc]), True)
        else:
            gpu_reg.import_block(_lib.BLOCK_FFM, ora.ffm_weights.copy(), False)
    for l in range(ora.nn_layer_count):
        n, nbytes = gpu_reg.block_len(_lib.BLOCK_NN0 + l)
        if nbytes == n * 8:
            gpu_reg.import_block(_lib.BLOCK_NN0 + l, np.concatenate([ora.nn_weights(l), ora.nn_acc(l)]), True)
        else:
            gpu_reg.import_block(_lib.BLOCK_NN0 + l, ora.nn_weights(l).copy(), False)


def sync_oracle_from_gpu(ora, gpu_reg):
    """Load the GPU regressor's current tables (weights and optimizer state) into the oracle."""
    from fwumious_wabbit_b200 import _lib

    ora.lr_table[:, :] = gpu_reg.get_lr_table()
    n, _ = gpu_reg.block_len(_lib.BLOCK_FFM)
    if n:
        w, acc = gpu_reg.get_ffm()
        ora.ffm_weights[:] = w
        ora.ffm_acc[:] = acc
    for l in range(ora.nn_layer_count):
        w, acc = gpu_reg.get_nn(l)
        ora.nn_weights(l)[:] = w
        ora.nn_acc(l)[:] = acc
