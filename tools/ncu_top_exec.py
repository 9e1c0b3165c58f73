#!/usr/bin/env python
"""Instruction-count profile by source line region: groups SASS of an ncu report into buckets of consecutive
instructions and prints executed-instruction share.  usage: ncu_top_exec.py report.ncu-rep [bucket]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ex = [int(r["Instructions Executed"] or 0) for r in rows]; sm = [int(r["# Samples"] or 0) for r in rows]
T = sum(ex); S = sum(sm)
for b in range(0, len(rows), B):
    e = sum(ex[b:b+B]); s = sum(sm[b:b+B])
    ops = {}
    for r in rows[b:b+B]:
        op = r["Source"].strip().split()[0 if not r["Source"].strip().startswith("@") else 1].split(".")[0]
        ops[op] = ops.get(op, 0) + int(r["Instructions Executed"] or 0)
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print("#%4d-%4d exec %5.1f%% samples %5.1f%%  %s" % (b, b + B - 1, 100.0 * e / T, 100.0 * s / max(S, 1), [(k, round(100.0 * v / T, 1)) for k, v in top]))
