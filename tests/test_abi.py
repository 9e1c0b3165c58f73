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


# ------------------------------------------------------------------ struct layout: header == ctypes == the Rust stub
def _c_struct_fields(header_text, name):
    """[(field, ctype, array_len)] of `typedef struct <name> { ... } <name>;` (comments stripped)."""
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header_text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const )?(\w+) ?(\*?)(.*)", decl)
        base, ptr, names = m.group(2), m.group(3), m.group(4)
        for nm in names.split(","):
            nm = nm.strip()
            is_ptr = bool(ptr) or nm.startswith("*")
            nm = nm.lstrip("* ")
            arr = re.match(r"(\w+)\[(\w+)\]", nm)
            if arr:
                out.append((arr.group(1), base, arr.group(2), is_ptr))
            else:
                out.append((nm, base, None, is_ptr))
    return out


def test_model_desc_layout_matches_everywhere(tmp_path):
    """include/fwgpu.h, the ctypes mirror and the Rust #[repr(C)] stub in INTEGRATION.md declare fwgpu_model_desc with
    the same fields in the same order, and the compiler's sizeof/offsetof agree with ctypes (a drifted mirror shifts every
    later field silently)."""
    import subprocess

    hdr = open(os.path.join(ROOT, "include", "fwgpu.h")).read()
    c_fields = _c_struct_fields(hdr, "fwgpu_model_desc")
    c_names = [f[0] for f in c_fields]
    py_names = [f[0] for f in _lib.ModelDesc._fields_]
    assert c_names == py_names
    # the Rust stub
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    rust = re.search(r"pub struct FwgpuModelDesc \{(.*?)\n\}", md, re.S).group(1)
    rust = re.sub(r"//[^\n]*", "", rust)
    rust_fields = re.findall(r"pub (\w+): ([^,]+),", rust)
    assert [f[0] for f in rust_fields] == c_names
    rust_type = {"float": "f32", "uint32_t": "u32", "uint8_t": "u8"}
    for (name, base, arr, is_ptr), (_, rt) in zip(c_fields, rust_fields):
        rt = rt.strip()
        want = f"*const {rust_type[base]}" if is_ptr else (f"[{rust_type[base]}; 8]" if arr else rust_type[base])
        assert rt == want, (name, rt, want)
    # the initialiser in desc_from() names every field as well
    init = re.search(r"FwgpuModelDesc \{\n(.*?)\n    \}\n\}", md, re.S).group(1)
    assert [m for m in re.findall(r"(\w+):", re.sub(r"//[^\n]*", "", init)) if m in c_names] == c_names
    # sizeof / offsetof from the C compiler vs ctypes
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fwgpu.h"', "int main(void){",
             'printf("sizeof %zu\\n", sizeof(fwgpu_model_desc));']
    lines += [f'printf("{n} %zu\\n", offsetof(fwgpu_model_desc, {n}));' for n in c_names]
    lines += ['printf("batch %zu\\n", sizeof(fwgpu_batch));', "return 0;}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(got["sizeof"]) == ctypes.sizeof(_lib.ModelDesc)
    for n in c_names:
        assert int(got[n]) == getattr(_lib.ModelDesc, n).offset, n
    assert int(got["batch"]) == ctypes.sizeof(_lib.Batch)


def test_rust_extern_block_matches_the_header():
    """Every `pub fn` in INTEGRATION.md's extern "C" block is a prototype of include/fwgpu.h with the same parameters in the
    same order and equivalent types (a drifted binding compiles and then corrupts the call)."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fwgpu.h")).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\n((?:const )?\w+ \*?)\s*(fwgpu_\w+)\(([^)]*)\);", hdr):
        protos[name] = (ret.strip(), [a.strip() for a in args.replace("\n", " ").split(",")] if args.strip() not in ("", "void") else [])
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', md, re.S).group(1)
    block = re.sub(r"//[^\n]*", "", block)
    fns = re.findall(r"pub fn (\w+)\((.*?)\)\s*(?:->\s*([^;]+))?;", block, re.S)
    assert len(fns) >= 9

    def c_to_rust(t):
        t = re.sub(r"\s+", " ", t).strip()
        t = re.sub(r" (\w+)$", "", t) if not t.endswith("*") else t      # drop the parameter name
        t = t.replace("fwgpu_status", "i32")
        table = {"int": "c_int", "uint32_t": "u32", "uint64_t": "u64", "float": "f32", "void": "c_void", "char": "c_char", "fwgpu_ctx": "FwgpuCtx",
                 "fwgpu_model_desc": "FwgpuModelDesc", "i32": "i32"}
        m = re.match(r"(const )?(\w+) ?(\**)$", t)
        assert m, t
        const, base, stars = m.group(1), table[m.group(2)], m.group(3)
        out = base
        for i in range(len(stars)):
            out = ("*const " if (const and i == 0) else "*mut ") + out
        return out

    for name, args, ret in fns:
        assert name in protos, f"{name} is not declared in include/fwgpu.h"
        c_ret, c_args = protos[name]
        r_args = [a.split(":", 1)[1].strip() for a in re.sub(r"\s+", " ", args).split(",") if a.strip()]
        # a C parameter "type *name" / "type name": split the name off before comparing
        want = [c_to_rust(re.sub(r"(\w+)$", "", a).strip() if not a.rstrip().endswith("*") else a) for a in c_args]
        assert r_args == want, (name, r_args, want)
        assert (ret or "").strip() == ("" if c_ret == "void" else c_to_rust(c_ret)), (name, ret, c_ret)


def test_ctypes_bindings_have_the_headers_arity():
    """Every function the Python mirror binds with argtypes takes as many parameters as its prototype in include/*.h."""
    from fwumious_wabbit_b200 import host

    L = host._L()
    checked = 0
    for h in ("fwgpu.h", "fwhost.h"):
        hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", h)).read(), flags=re.S)
        for _, name, args in re.findall(r"\n((?:const )?\w+ \*?)\s*(fw(?:gpu|host)_\w+)\(([^)]*)\);", hdr):
            n = 0 if args.strip() in ("", "void") else len(args.split(","))
            at = getattr(L, name).argtypes
            if at is not None:
                assert len(at) == n, (name, len(at), n)
                checked += 1
    assert checked >= 50
