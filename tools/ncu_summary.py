#!/usr/bin/env python
"""Summarise one `ncu --set full` capture of a learn kernel: key metrics and stall reasons as text (-> profiles/r02_*.txt) and,
with --traffic NAME --examples N, the DRAM bytes per example as profiles/traffic_NAME.json tagged with the hash of the kernel
source the capture was taken on (bench.py prints `roofline.traffic` only while that hash still matches).
usage: ncu_summary.py report.ncu-rep [kernel-substring] [--traffic NAME --examples N_PER_LAUNCH] [--top K]"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", nargs="?", default="")
    ap.add_argument("--traffic")
    ap.add_argument("--examples", type=float, default=0)
    ap.add_argument("--top", type=int, default=14)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if a.kernel not in d.get("Kernel Name", ""):
            continue
        u = dict(zip(hdr, units))
        print("# ncu --set full --clock-control none, key metrics + top stalled SASS (tools/ncu_summary.py)")
        print("%-70s %s" % ("Kernel Name", d["Kernel Name"]))
        for k in KEYS:
            if k in d:
                print("%-70s %s %s" % (k, d[k], u.get(k, "")))
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[k]) for k in hdr if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]}
        tot = sum(stalls.values())
        print("warp stall reasons (share of stalled+selected warps per issue):", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]))
        if a.traffic and a.examples:
            rd = float(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
            import bench

            out = {"dram_bytes_per_example": (rd + wr) / a.examples, "dram_read_bytes": rd, "dram_write_bytes": wr, "examples_in_launch": a.examples,
                   "kernel": d["Kernel Name"], "launch_us": float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u["gpu__time_duration.sum"], 1),
                   "kernel_source_sha": bench.kernel_source_sha(), "source": os.path.basename(a.report)}
            p = os.path.join(ROOT, "profiles", f"traffic_{a.traffic}.json")
            json.dump(out, open(p, "w"), indent=1)
            print("wrote", p, json.dumps(out))
        break
    top = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_top_stalls.py"), a.report, str(a.top), a.kernel], capture_output=True, text=True).stdout
    print(top)


if __name__ == "__main__":
    main()
