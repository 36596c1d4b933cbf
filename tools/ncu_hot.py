#!/usr/bin/env python
"""Hot CUDA source lines of one kernel in an ncu report (needs -lineinfo + --import-source on):
   python tools/ncu_hot.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, os
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
cur, agg, seen_kernel = "", [], 0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1])
    elif r[0] == "Kernel Name":
        seen_kernel += 1
    elif len(r) > 8 and r[2] == "-" and r[0].isdigit() and seen_kernel <= 1:
        agg.append((int(r[6] or 0), int(r[7] or 0), cur, r[0], r[1].strip()[:100]))
tot_s = sum(a[0] for a in agg) or 1
tot_i = sum(a[1] for a in agg) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
for s, i, f, ln, text in sorted(agg, key=lambda a: -a[0])[:top]:
    print(f"{s / tot_s * 100:5.1f}% smp {i / tot_i * 100:5.1f}% inst  {f}:{ln:>4s}  {text}")
