#!/usr/bin/env python
"""One line per kernel launch of an ncu --set full report: python tools/ncu_summary.py report.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
cols = [("Kernel Name", "kernel", 46), ("launch__grid_size", "grid", 7), ("launch__block_size", "blk", 5), ("launch__registers_per_thread", "regs", 5),
        ("gpu__time_duration.sum", "time", 12), ("dram__bytes_read.sum", "dram_rd", 14), ("dram__bytes_write.sum", "dram_wr", 14),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 7), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 7),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%", 7), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 7),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 7)]
print("per-launch values from ncu --set full --clock-control none; time under ncu is cold-cache/serialised (compare shares); units as printed by ncu")
print(" ".join(f"{t:>{w}s}" if i else f"{t:{w}s}" for i, (_, t, w) in enumerate(cols)))
for r in rows[2:]:
    parts = []
    for i, (name, _, w) in enumerate(cols):
        c = col(name)
        v = r[c] if c is not None else "-"
        if name == "Kernel Name":
            v = v.split("(")[0].replace("void ", "")[:w]
            parts.append(f"{v:{w}s}")
        else:
            u = units[c] if c is not None and name.endswith(("sum",)) and "bytes" in name or name.startswith("gpu__time") else ""
            try:
                v = f"{float(v):.1f}" if "pct" in name else (f"{float(v):.6g}" if "." in v else v)
            except ValueError:
                pass
            parts.append(f"{(v + (' ' + u if u else '')):>{w}s}")
    print(" ".join(parts))
