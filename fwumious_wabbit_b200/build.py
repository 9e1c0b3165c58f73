"""Builds libfwgpu.so (CUDA kernels + C ABI + host-side C++) in-tree for sm_100a."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfwgpu.so")
CLI_SRC = os.path.join(HERE, "cli", "fwgpu_main.cpp")
CLI_OUT = os.path.join(HERE, "fwgpu")
SYNTH_SRC = os.path.join(HERE, "synth_src", "fwsynth.cpp")  # bench / test data generator: its own library, not product code
SYNTH_OUT = os.path.join(HERE, "libfwsynth.so")
UMMA_TEST_SRC = os.path.join(os.path.dirname(HERE), "tools", "umma_gemm_test.cu")
UMMA_TEST_OUT = os.path.join(HERE, "umma_gemm_test")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    # (no -split-compile: it shortens the build from 140 s to 70 s but changes register allocation from build to build --
    #  the narrow fused kernel went from 8 to 176 bytes of spills and lost 12 % in one such build)
    "-shared",
]


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(root, f))
    return out


def deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(os.path.dirname(HERE), "include", "fwgpu.h"))
    out.append(os.path.join(os.path.dirname(HERE), "include", "fwhost.h"))
    out.append(CLI_SRC)
    out.append(UMMA_TEST_SRC)
    out.append(SYNTH_SRC)
    return out


def up_to_date():
    if not all(os.path.exists(p) for p in (OUT, CLI_OUT, UMMA_TEST_OUT, SYNTH_OUT)):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in deps())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libfwgpu.so must be built where the CUDA toolkit is (it ships prebuilt to the GPU box)")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    cmd1 = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", SYNTH_OUT, SYNTH_SRC]
    res1 = subprocess.run(cmd1, capture_output=True, text=True)
    with open(os.path.join(HERE, "build.log"), "a") as f:
        f.write(" ".join(cmd1) + "\n" + res1.stdout + res1.stderr)
    if res1.returncode != 0:
        sys.stderr.write(res1.stdout + res1.stderr)
        raise RuntimeError("building libfwsynth.so failed")
    # the command-line front end (fwgpu): plain C++ against the two C ABIs, finds the library next to itself
    cmd2 = ["g++", "-O2", "-std=c++17", "-Wall", "-o", CLI_OUT, CLI_SRC, "-L" + HERE, "-lfwgpu", "-lz", "-Wl,-rpath,$ORIGIN"]
    res2 = subprocess.run(cmd2, capture_output=True, text=True)
    with open(os.path.join(HERE, "build.log"), "a") as f:
        f.write(" ".join(cmd2) + "\n" + res2.stdout + res2.stderr)
    if res2.returncode != 0:
        sys.stderr.write(res2.stdout + res2.stderr)
        raise RuntimeError("building the fwgpu command-line front end failed")
    # stand-alone check of the tcgen05 GEMM tiles against a double-precision product (tests/test_gpu_head.py runs it on the GPU)
    cmd3 = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-o", UMMA_TEST_OUT, UMMA_TEST_SRC]
    res3 = subprocess.run(cmd3, capture_output=True, text=True)
    with open(os.path.join(HERE, "build.log"), "a") as f:
        f.write(" ".join(cmd3) + "\n" + res3.stdout + res3.stderr)
    if res3.returncode != 0:
        sys.stderr.write(res3.stdout + res3.stderr)
        raise RuntimeError("building tools/umma_gemm_test.cu failed")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
