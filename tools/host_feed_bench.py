"""Throughput of the host-side feeders of the hot path (no GPU needed): the VW text parser (parser.rs:214-461 restated in
csrc/host/fwhost.cpp) and the .fwcache reader (cache.rs:187-232), on config 3's line shape (39 namespaces, ~290 B of text,
168 B of record per example).  The reference parses on one thread (main.rs:213-239); the batch parser here uses n threads.

    python tools/host_feed_bench.py [n_lines] > profiles/rNN_host_feed.txt
"""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fwumious_wabbit_b200 import host, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    L = host._L()
    print(f"host cores: {os.cpu_count()}")
    for name in ("c3", "c2"):
        w = synth.workload(name)
        vw = host.VwNamespaceMap.new("".join(f"{c},f{c}\n" for c in w.ns_names))
        text = "".join(w.line(i) + "\n" for i in range(n)).encode()
        p = host.VowpalParser(vw)
        out = np.empty(len(text), np.uint32)
        off = np.empty(n + 1, np.uint32)
        err = C.create_string_buffer(512)
        nw = C.c_uint64()
        for th in sorted({1, 2, 4, 8, os.cpu_count()}):
            best = 1e9
            for _ in range(3):
                t = time.perf_counter()
                k = L.fwhost_parser_parse_text(p.h, text, len(text), out.ctypes.data_as(C.c_void_p), out.size, off.ctypes.data_as(C.c_void_p), n,
                                               th, C.byref(nw), err, 512)
                best = min(best, time.perf_counter() - t)
                assert k == n, err.value
            print(f"{name}: parse_text {th:2d} threads  {n / best / 1e6:6.2f} M lines/s  {len(text) / best / 1e6:6.0f} MB/s of text ({len(text) / n:.0f} B/line)")
        recs = out[: nw.value]
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "feed.vw.fwcache")
            t = time.perf_counter()
            host.cache_write(path, vw, recs)
            tw = time.perf_counter() - t
            best = 1e9
            for _ in range(3):
                r, nwords, o, b = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_void_p()
                t = time.perf_counter()
                k = L.fwhost_cache_read(path.encode(), None, C.byref(r), C.byref(nwords), C.byref(o), C.byref(b), err, 512)
                best = min(best, time.perf_counter() - t)
                assert k == n and nwords.value == recs.size, err.value
                for q in (r, o, b):
                    L.fwhost_free(q)
            print(f"{name}: cache_write {recs.nbytes / tw / 1e6:6.0f} MB/s;  cache_read {n / best / 1e6:6.2f} M records/s  {recs.nbytes / best / 1e6:6.0f} MB/s "
                  f"({recs.nbytes / n:.0f} B/record)")


if __name__ == "__main__":
    main()
