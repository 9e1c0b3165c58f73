"""The device computes the sigmoid with a restatement of glibc's expf (the reference calls libm
through Rust's f32::exp, block_loss_functions.rs:15-17).  CPU: the host copy of that routine equals
libm's expf bit for bit.  GPU: the device routine equals the host copy."""
import ctypes as C

import numpy as np
import pytest

from fwumious_wabbit_b200 import _lib


def _host_expf(x):
    L = _lib.lib()
    L.fwhost_expf_libm_array.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.fwhost_expf_libm_array.restype = None
    out = np.empty_like(x)
    L.fwhost_expf_libm_array(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), x.size)
    return out


def _libm_expf(x):
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    return np.array([libm.expf(float(v)) for v in x], dtype=np.float32)


def _inputs(n, seed):
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.uniform(-50, 50, n), rng.normal(0, 2, n), rng.normal(0, 1e-3, n // 4),
                        np.array([0.0, -0.0, 50.0, -50.0, 1e-30, -1e-30, 0.5, -0.5, 17.25, -33.3])]).astype(np.float32)
    return x


def test_host_expf_equals_libm_bitwise():
    x = _inputs(60_000, 1)
    a, b = _host_expf(x), _libm_expf(x)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.gpu
def test_device_logistic_equals_host_bitwise():
    import fwumious_wabbit_b200 as fw

    re = fw.Regressor(fw.ModelInstance.new_empty())
    x = _inputs(500_000, 2)
    e = _host_expf(-x)
    want = (np.float32(1.0) / (np.float32(1.0) + e)).astype(np.float32)
    got = re.debug_logistic(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
