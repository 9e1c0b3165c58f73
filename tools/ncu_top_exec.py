#!/usr/bin/env python
"""Instruction-count profile per captured kernel: groups the SASS of an ncu report into buckets of consecutive
instructions and prints each bucket's share of executed instructions and stall samples.
usage: ncu_top_exec.py report.ncu-rep [bucket] [kernel-substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
want = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
seen = set()
for a, b_ in zip(starts[:-1], starts[1:]):
    name = next(csv.reader([lines[a]]))[1]
    if want not in name or name in seen:
        continue
    seen.add(name)
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[a + 1:b_]))))
    ex = [int(r["Instructions Executed"] or 0) for r in rows]; sm = [int(r["# Samples"] or 0) for r in rows]
    T = sum(ex); S = sum(sm)
    print("==", name, "executed", T)
    for b in range(0, len(rows), B):
        e = sum(ex[b:b+B]); s = sum(sm[b:b+B])
        ops = {}
        for r in rows[b:b+B]:
            src = r["Source"].strip().split()
            op = src[1 if src[0].startswith("@") else 0].split(".")[0]
            ops[op] = ops.get(op, 0) + int(r["Instructions Executed"] or 0)
        top = sorted(ops.items(), key=lambda x: -x[1])[:6]
        print("#%4d-%4d exec %5.1f%% samples %5.1f%%  %s" % (b, b + B - 1, 100.0 * e / max(T, 1), 100.0 * s / max(S, 1), [(k, round(100.0 * v / max(T, 1), 1)) for k, v in top]))
