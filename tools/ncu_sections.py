#!/usr/bin/env python
"""Per-source-line instruction / stall-sample table of an .ncu-rep (first profiled launch),
optionally grouped by line ranges.  Usage: tools/ncu_sections.py <rep> <file.cuh> [top N]"""
import csv, io, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
cur, agg = None, {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if len(r) > 8 and r[2] == "-":
        try:
            agg[(cur, int(r[0]))] = (int(r[4] or 0), int(r[7] or 0), int(r[8] or 0), r[1].strip()[:100])
        except ValueError:
            pass
ti = sum(v[1] for v in agg.values()) or 1
ts = sum(v[0] for v in agg.values()) or 1
print(f"total warp-inst {ti}, stall samples {ts}")
print("by file:")
byf = {}
for (f, l), v in agg.items():
    b = byf.setdefault(f, [0, 0]); b[0] += v[0]; b[1] += v[1]
for f, b in byf.items():
    print(f"  {f:28s} inst {b[1]:9d} ({100*b[1]/ti:4.1f}%) samples {b[0]:6d} ({100*b[0]/ts:4.1f}%)")
print("top lines by samples:")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<4d} inst={v[1]:8d} ({100*v[1]/ti:4.1f}%) samp={v[0]:5d} ({100*v[0]/ts:4.1f}%) lanes={v[2]/max(v[1],1):4.1f} | {v[3]}")
