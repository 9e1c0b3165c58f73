#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu report's source page (CSV).  usage: ncu_top_stalls.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows), "executed", sum(int(r["Instructions Executed"] or 0) for r in rows))
idx = {id(r): i for i, r in enumerate(rows)}
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
    n = int(r["# Samples"] or 0)
    st = {k[6:]: int(r[k]) for k in r if k.startswith("stall_") and "(" not in k and r[k] and int(r[k]) > 0}
    main = sorted(st.items(), key=lambda x: -x[1])[:2]
    print("%5.1f%% #%4d  %-70s %s" % (100.0 * n / max(tot, 1), idx[id(r)], r["Source"].strip()[:70], main))
