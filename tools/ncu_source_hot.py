#!/usr/bin/env python
"""Per-source-line hot spots of a kernel from an .ncu-rep (needs -lineinfo + --import-source on).
Usage: python tools/ncu_source_hot.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
li, si = hdr.index("Line No"), hdr.index("Source")
src_lines = {}
src_out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
for r in csv.reader(io.StringIO(src_out)):   # the source embedded in the report (not the file on disk, which may have moved on)
    if len(r) >= 2 and r[0].isdigit():
        src_lines[r[0]] = r[1].strip()[:110]
ii = hdr.index("Instructions Executed"); sa = hdr.index("# Samples")
tot_i = tot_s = 0
agg = collections.defaultdict(lambda: [0, 0])
for r in rows[hi + 1:]:
    if len(r) <= ii or not r[ii]:
        continue
    try:
        n = int(float(r[ii])); s = int(float(r[sa] or 0))
    except ValueError:
        continue
    tot_i += n; tot_s += s
    agg[r[li]][0] += n; agg[r[li]][1] += s
lines = [(v[0], v[1], k, src_lines.get(k, "")) for k, v in agg.items()]
print(f"total instructions executed {tot_i}, samples {tot_s}")
for n, s, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100.0*n/tot_i:6.2f}% inst {100.0*s/max(1,tot_s):6.2f}% smp  L{ln:>4s}  {src}")
