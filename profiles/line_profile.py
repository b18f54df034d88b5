#!/usr/bin/env python
"""Aggregates an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) by CUDA source line
using nvdisasm -g line markers of the cubin extracted from the library.

usage: line_profile.py <source.csv> <cubin> <kernel-mangled-substring> [top]
"""
import collections
import csv
import re
import subprocess
import sys

src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the kernel's section: map instruction ordinal -> (file:line, inline chain)
lines = []
in_k = False
cur = ("?", 0)
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        in_k = kname in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
n = min(len(body), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot_inst = tot_samp = 0
for k in range(n):
    r = body[k]
    inst = int(r[ci["Instructions Executed"]] or 0)
    samp = int(r[ci["# Samples"]] or 0)
    noi = int(r[ci["stall_no_inst"]] or 0)
    key = lines[k][:2]
    a = agg[key]
    a[0] += inst; a[1] += samp; a[2] += noi; a[3] += 1
    tot_inst += inst; tot_samp += samp
print(f"kernel instrs (static) csv={len(body)} disasm={len(lines)}; executed warp-instr={tot_inst}; samples={tot_samp}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), a in agg.items():
    byfile[f][0] += a[0]; byfile[f][1] += a[1]; byfile[f][2] += a[3]
print("by file:")
for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:40s} exec={a[0] / tot_inst:6.3f} samples={a[1] / max(tot_samp, 1):6.3f} static={a[2]}")
print("top lines by executed instructions:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {f}:{l:<5d} exec={a[0] / tot_inst:6.3f} samples={a[1] / max(tot_samp, 1):6.3f} no_inst={a[2] / max(tot_samp, 1):6.3f} static={a[3]}")

# coarse buckets: (file, first line, last line, label) given as extra args file:lo-hi=label
buckets = [a for a in sys.argv[5:] if "=" in a]
if buckets:
    print("buckets:")
    for bdef in buckets:
        rng, label = bdef.split("=")
        f, lr = rng.split(":")
        lo, hi = [int(x) for x in lr.split("-")]
        ex = sum(a[0] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        sm = sum(a[1] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        print(f"  {label:32s} exec={ex / tot_inst:6.3f} ({ex / 312500:7.0f}/warp) samples={sm / max(tot_samp, 1):6.3f}")
