"""ctypes binding of libfwgpu.so (include/fwgpu.h).  There is no fallback: if the CUDA library
is missing or cannot be loaded this module raises, it never routes anywhere else."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfwgpu.so")

MAX_NN_LAYERS = 8
LUT_SIZE = 2048

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_TOO_LARGE, ERR_IMMUTABLE, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
OPT_SGD, OPT_ADAGRAD_FLEX, OPT_ADAGRAD_LUT = 0, 1, 2
BLOCK_LR, BLOCK_FFM, BLOCK_NN0 = 0, 1, 2

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)


class ModelDesc(C.Structure):
    _fields_ = [
        ("learning_rate", C.c_float), ("power_t", C.c_float), ("init_acc_gradient", C.c_float),
        ("ffm_learning_rate", C.c_float), ("ffm_power_t", C.c_float), ("ffm_init_acc_gradient", C.c_float),
        ("nn_learning_rate", C.c_float), ("nn_power_t", C.c_float), ("nn_init_acc_gradient", C.c_float),
        ("bit_precision", C.c_uint32), ("ffm_bit_precision", C.c_uint32), ("ffm_k", C.c_uint32),
        ("ffm_num_fields", C.c_uint32), ("num_combos", C.c_uint32), ("optimizer", C.c_uint32),
        ("immutable", C.c_uint32),
        ("ffm_init_width", C.c_float), ("ffm_init_zero_band", C.c_float), ("ffm_init_center", C.c_float),
        ("nn_num_layers", C.c_uint32),
        ("nn_width", C.c_uint32 * MAX_NN_LAYERS), ("nn_relu", C.c_uint32 * MAX_NN_LAYERS),
        ("nn_init", C.c_uint32 * MAX_NN_LAYERS),
        ("n_namespaces", C.c_uint32), ("ns_is_f32", u8p),
        ("n_combos", C.c_uint32), ("combo_off", u32p), ("combo_ns", u32p), ("combo_weight", f32p),
        ("add_constant", C.c_uint32),
        ("field_off", u32p), ("field_ns", u32p),
        ("max_ffm_per_example", C.c_uint32), ("max_lr_per_example", C.c_uint32),
        ("hogwild_ramp_div", C.c_uint32), ("hogwild_max_inflight", C.c_uint32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_examples", C.c_uint32), ("labels", f32p), ("importance", f32p),
        ("lr_off", u32p), ("lr_hash", u32p), ("lr_val", f32p), ("lr_combo", u32p),
        ("ffm_off", u32p), ("ffm_hash", u32p), ("ffm_val", f32p), ("ffm_field", u32p),
    ]


class FwgpuError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"fwgpu status {status}: {message}")
        self.status = status


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m fwumious_wabbit_b200.build` (needs nvcc). "
            "fwumious_wabbit_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.fwgpu_version.restype = C.c_char_p
    L.fwgpu_last_error.restype = C.c_char_p
    L.fwgpu_last_error.argtypes = [vp]
    L.fwgpu_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(vp)]
    L.fwgpu_create_sharded.argtypes = [C.POINTER(ModelDesc), C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(vp)]
    L.fwgpu_create_sharded.restype = C.c_int32
    L.fwgpu_debug_shard_plan.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, vp, vp]
    L.fwgpu_debug_shard_plan.restype = C.c_int32
    L.fwgpu_shard_barrier.argtypes = [vp]
    L.fwgpu_shard_barrier.restype = C.c_int32
    L.fwgpu_shard_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.fwgpu_shard_info.restype = C.c_int32
    L.fwgpu_destroy.argtypes = [vp]
    L.fwgpu_destroy.restype = None
    L.fwgpu_sync.argtypes = [vp]
    L.fwgpu_stream.argtypes = [vp]
    L.fwgpu_stream.restype = vp
    L.fwgpu_launch_count.argtypes = [vp]
    L.fwgpu_launch_count.restype = C.c_uint64
    L.fwgpu_learn_batch.argtypes = [vp, C.POINTER(Batch), vp, C.c_int]
    L.fwgpu_predict_batch.argtypes = [vp, C.POINTER(Batch), vp]
    L.fwgpu_learn_records.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, C.c_int]
    L.fwgpu_translate_records.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, vp, vp, vp, vp, vp, C.c_uint64,
                                          vp, vp, vp, vp, C.c_uint64]
    L.fwgpu_dataset_upload.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(vp)]
    L.fwgpu_dataset_learn.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_int]
    L.fwgpu_dataset_free.argtypes = [vp, vp]
    L.fwgpu_dataset_free.restype = None
    L.fwgpu_block_len.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.fwgpu_export_block.argtypes = [vp, C.c_int, vp, C.c_uint64]
    L.fwgpu_import_block.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_int]
    L.fwgpu_get_lut.argtypes = [vp, C.c_int, vp]
    L.fwgpu_debug_logistic.argtypes = [vp, vp, vp, C.c_uint64]
    L.fwgpu_debug_logistic.restype = C.c_int32
    L.fwgpu_debug_path_counts.argtypes = [vp, vp]
    L.fwgpu_debug_path_counts.restype = C.c_int32
    L.fwgpu_set_examples_seen.argtypes = [vp, C.c_uint64]
    L.fwgpu_get_examples_seen.argtypes = [vp]
    L.fwgpu_get_examples_seen.restype = C.c_uint64
    L.fwgpu_set_profiling.argtypes = [vp, C.c_int]
    L.fwgpu_kernel_time.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.fwgpu_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.fwgpu_host_free.argtypes = [vp]
    L.fwgpu_host_free.restype = None
    for name in ("fwgpu_create", "fwgpu_sync", "fwgpu_learn_batch", "fwgpu_predict_batch", "fwgpu_learn_records",
                 "fwgpu_translate_records", "fwgpu_dataset_upload", "fwgpu_dataset_learn", "fwgpu_block_len",
                 "fwgpu_export_block", "fwgpu_import_block", "fwgpu_get_lut", "fwgpu_set_profiling",
                 "fwgpu_kernel_time", "fwgpu_host_alloc", "fwgpu_set_examples_seen"):
        getattr(L, name).restype = C.c_int32
    _lib = L
    return L


# every symbol include/fwgpu.h declares (checked by tests/test_abi.py without a GPU)
EXPORTED_SYMBOLS = [
    "fwgpu_create", "fwgpu_create_sharded", "fwgpu_debug_shard_plan", "fwgpu_shard_barrier", "fwgpu_shard_info", "fwgpu_destroy", "fwgpu_last_error", "fwgpu_sync", "fwgpu_stream", "fwgpu_launch_count",
    "fwgpu_learn_batch", "fwgpu_predict_batch", "fwgpu_learn_records", "fwgpu_translate_records",
    "fwgpu_dataset_upload", "fwgpu_dataset_learn", "fwgpu_dataset_free",
    "fwgpu_block_len", "fwgpu_export_block", "fwgpu_import_block", "fwgpu_get_lut",
    "fwgpu_set_examples_seen", "fwgpu_get_examples_seen", "fwgpu_debug_logistic", "fwgpu_debug_path_counts",
    "fwgpu_set_profiling", "fwgpu_kernel_time", "fwgpu_host_alloc", "fwgpu_host_free", "fwgpu_version",
]
