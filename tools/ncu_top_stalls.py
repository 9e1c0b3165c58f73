#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu report's source page (CSV), per captured kernel.
usage: ncu_top_stalls.py report.ncu-rep [N] [kernel-substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
for a, b in zip(starts[:-1], starts[1:]):
    name = next(csv.reader([lines[a]]))[1]
    if want not in name:
        continue
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[a + 1:b]))))
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    print("==", name)
    print("total samples", tot, "instructions", len(rows), "executed", sum(int(r["Instructions Executed"] or 0) for r in rows))
    agg = {}
    for r in rows:
        for k in r:
            if k and k.startswith("stall_") and "(" not in k and r[k]:
                agg[k[6:]] = agg.get(k[6:], 0) + int(r[k])
    print("by reason:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    idx = {id(r): i for i, r in enumerate(rows)}
    for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
        n = int(r["# Samples"] or 0)
        st = {k[6:]: int(r[k]) for k in r if k and k.startswith("stall_") and "(" not in k and r[k] and int(r[k]) > 0}
        main = sorted(st.items(), key=lambda x: -x[1])[:2]
        print("%5.1f%% #%4d  %-70s %s" % (100.0 * n / max(tot, 1), idx[id(r)], r["Source"].strip()[:70], main))
