"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/fwgpu.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import pytest

import fwumious_wabbit_b200 as fw
from fwumious_wabbit_b200 import _lib, build
from tests.conftest import ROOT, has_cuda


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    L = _lib.lib()
    assert b"sm_100a" in L.fwgpu_version()


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "fwgpu.h")).read()
    declared = set(re.findall(r"\b(fwgpu_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"fwgpu_status"}
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    hdr2 = open(os.path.join(ROOT, "include", "fwhost.h")).read()
    for s in set(re.findall(r"\b(fwhost_[a-z_0-9]+)\s*\(", hdr2)):
        assert hasattr(L, s), s


def test_sass_has_vector_atomics():
    """The scatter really is 128-bit atomics (ATOMG/REDG .F32x4), not scalar CAS loops."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "ATOMG.E.ADD.F32x4" in sass and "REDG.E.ADD.F32x4" in sass
    assert "sm_100a" in sass


def test_sass_has_tensor_core_instructions():
    """The head's GEMMs really are tcgen05: UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit) and LDTM (tcgen05.ld) in the cubin."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTCBAR" in sass and "LDTM" in sass


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (Python or C++/CUDA) imports, links or opens it."""
    pkg = os.path.dirname(os.path.abspath(fw.__file__))
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert not re.search(r"(from|import)\s+oracle|fw_oracle|libfworacle|#include[^\n]*oracle", text), os.path.join(root, f)


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(_lib.FwgpuError) as ei:
        fw.Regressor(fw.ModelInstance.new_empty())
    assert ei.value.status == _lib.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)
