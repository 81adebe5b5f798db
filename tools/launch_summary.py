#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum,sm__pipe_alu_cycles_active...,dram__bytes_* --csv` launch
list: launches, total time, share of the GPU time, time-weighted ALU-pipe utilisation, DRAM bytes.
usage: launch_summary.py LAUNCHES.csv "the command that was profiled" > profiles/NAME_launches_summary.txt"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if r]
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    head, body = rows[hi], rows[hi + 1:]
    iK, iM, iV, iID, iU = (head.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    per, unit = collections.OrderedDict(), {}
    for r in body:
        if len(r) <= iV:
            continue
        unit[r[iM]] = r[iU]
        d = per.setdefault(r[iID], {"name": r[iK]})
        try:
            d[r[iM]] = float(r[iV].replace(",", ""))
        except ValueError:
            pass
    agg = collections.OrderedDict()
    for d in per.values():
        n = re.sub(r"\(int\)|\(bool\)|\(SwbScoreParams\)|void ", "", d["name"])
        a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
        t = d.get("gpu__time_duration.sum", 0.0)
        a[0] += 1
        a[1] += t
        a[2] += d.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
        a[3] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tu, du = unit.get("gpu__time_duration.sum", "ns"), unit.get("dram__bytes_read.sum", "byte")
    tf = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1e-6)
    df = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(du, 1e-9)
    tot = sum(a[1] for a in agg.values())
    print("# ncu launch list of `" + (sys.argv[2] if len(sys.argv) > 2 else "?") + "`")
    print("# (cold-cache, serialised launches: read the SHARES). kernel | launches | total ms | share | mean ALU pipe % | "
          "DRAM GB read+written   [csv units: " + tu + ", " + du + "]")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s %4d %10.2f ms %6.2f%%  alu %5.1f%%  dram %8.2f GB" % (
            n[:62], a[0], a[1] * tf, 100 * a[1] / tot, a[2] / a[1] if a[1] else 0, a[3] * df))


if __name__ == "__main__":
    main()
