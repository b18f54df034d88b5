#!/usr/bin/env python
"""Per-source-line executed-instruction counts of one kernel from an ncu source page.
usage: line_detail.py <source.csv> <cubin> <kernel-substring> <file.cuh> [min_per_warp]"""
import collections, csv, os, re, subprocess, sys
src_csv, cubin, kname, fname = sys.argv[1:5]
minpw = float(sys.argv[5]) if len(sys.argv) > 5 else 3
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inst = []; in_k = False; cur = ("?", 0)
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        in_k = kname in ln; continue
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): inst.append(cur)
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:2 + len(inst)]
agg = collections.Counter(); smp = collections.Counter(); nw = None
for k, r in enumerate(body):
    ex = int(r[ci["Instructions Executed"]] or 0)
    if nw is None: nw = ex
    agg[inst[k]] += ex; smp[inst[k]] += int(r[ci["# Samples"]] or 0)
lines = open(fname).read().splitlines()
base = os.path.basename(fname)
for (f, l), ex in sorted(agg.items(), key=lambda kv: kv[0][1]):
    if f != base or ex / nw < minpw: continue
    print(f"{l:5d} {ex / nw:7.1f} {smp[(f, l)]:6d}  {lines[l - 1].strip()[:110]}")
