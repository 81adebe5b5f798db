#!/usr/bin/env python3
"""Text summary of an `ncu --set full --import-source on` report for profiles/: per captured kernel the metrics the
roofline discussion uses (duration, ALU pipe, issue slots, warps, DRAM bytes, shared-memory wavefronts and bank
conflicts, registers), the warp-stall breakdown and the instructions with most stall samples.
usage: ncu_summarize.py REPORT.ncu-rep > profiles/NAME_details.txt"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.split("\n")) if r]
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    print("# %s -- summary by tools/ncu_summarize.py (ncu --set full --clock-control none --import-source on)" % rep)
    for r in rows[2:]:
        if len(r) < 10:
            continue
        print("\n== %s" % r[idx["Kernel Name"]])
        for m in METRICS:
            if m in idx and r[idx[m]] not in ("", "nan", "-nan"):
                print("  %-68s %s %s" % (m, r[idx[m]], units[idx[m]]))
        st = {}
        for h in head:
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[idx[h]])
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        print("  warp stall samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in
                                                  sorted(st.items(), key=lambda x: -x[1])[:10]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(src.split("\n")):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if b["name"] in seen or len(b["rows"]) < 3:
            continue
        seen.add(b["name"])
        hdr = b["rows"][0]
        data = [r for r in b["rows"][1:] if len(r) > 10]
        i_s, i_e = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        tot = sum(int(r[i_s]) for r in data) or 1
        base = int(data[0][0], 16)
        print("\n== %s: instructions with most warp-stall samples (of %d)" % (b["name"], tot))
        for r in sorted(sorted(data, key=lambda r: -int(r[i_s]))[:14], key=lambda r: int(r[0], 16)):
            why = sorted(((int(r[i]), h) for i, h in enumerate(hdr) if h.startswith("stall_") and
                          not h.endswith("(Not Issued)") and r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
            print("  +0x%04x %-58s %5.1f%%  executed %s  %s" % (int(r[0], 16) - base, r[1].strip()[:58],
                                                              100 * int(r[i_s]) / tot, r[i_e],
                                                              ", ".join("%s %d" % (h, v) for v, h in why)))


if __name__ == "__main__":
    main()
