"""Inputs for tools/host_fuzz/run.sh: mutated .fwcache files (plain and LZ4-framed), mutated regressor files and generated text
lines, written to the directory given as argv[1]."""
import ctypes as C
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from fwumious_wabbit_b200 import host, synth  # noqa: E402


def mutate(rnd, good):
    b = bytearray(good)
    for _ in range(rnd.randint(1, 4)):
        mode, pos = rnd.random(), rnd.randrange(len(b))
        if mode < 0.6:
            b[pos] = rnd.randrange(256)
        elif mode < 0.8:
            del b[pos:pos + rnd.randint(1, 50)]
        else:
            b[pos:pos] = bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 20)))
    if rnd.random() < 0.1:
        b = b[:rnd.randrange(len(b))]
    return bytes(b)


def main(d):
    os.makedirs(d, exist_ok=True)
    rnd = random.Random(5)
    w = synth.workload("c2")
    vw = host.VwNamespaceMap.new("".join(f"{c},feature{c}\n" for c in w.ns_names))
    recs = w.records(3000).reshape(-1)
    for suffix in ("vw.fwcache", "vw.gz.fwcache"):
        base = os.path.join(d, "base." + suffix)
        host.cache_write(base, vw, recs)
        good = open(base, "rb").read()
        for i in range(1500):
            open(os.path.join(d, f"m{i}.{suffix}"), "wb").write(mutate(rnd, good))
    vw = host.VwNamespaceMap.new("A,a\nBB,b\nC,c:f32\nD,d\n")
    open(os.path.join(d, "vwmap.json"), "w").write(vw.source_json)
    alphabet = "1-|: AaBCD.e05x\t'#\xe9||||    ::"
    with open(os.path.join(d, "lines.txt"), "w", encoding="utf-8") as f:
        for _ in range(200_000):
            f.write(rnd.choice(["1 ", "-1 ", "|", "1 0.5 ", "-1 |A ", "1 |C ", "|BB:2 "]) + "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 80))) + "\n")
    mi_json = host.model_instance_json_from_cmdline(["--keep", "A", "--ffm_k", "2", "--ffm_field", "A", "--ffm_field", "D", "--adaptive"], vw)
    lr = np.zeros(16, np.float32)
    ptrs = (C.c_void_p * 1)(lr.ctypes.data_as(C.c_void_p))
    sizes = (C.c_uint64 * 1)(lr.nbytes)
    err = C.create_string_buffer(512)
    base = os.path.join(d, "base.fw")
    assert host._L().fwhost_regressor_write(base.encode(), vw.source_json.encode(), mi_json.encode(), 16, ptrs, sizes, 1, err, 512) == 0
    good = open(base, "rb").read()
    for i in range(2000):
        open(os.path.join(d, f"r{i}.fw"), "wb").write(mutate(rnd, good))


if __name__ == "__main__":
    main(sys.argv[1])
