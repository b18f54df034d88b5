#!/usr/bin/env python
"""Aggregates an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) by CUDA source
function, using the nvdisasm -g line markers of the cubin the profile was taken with.

usage: line_profile.py <source.csv> <cubin> <kernel-mangled-substring> <csrc dir>
"""
import collections
import csv
import os
import re
import subprocess
import sys

src_csv, cubin, kname, csrc = sys.argv[1:5]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inst = []           # per instruction: (file, line, subroutine label or None)
in_k = False
cur = ("?", 0)
label = None
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        in_k = kname in ln
        label = None
        continue
    if not in_k:
        continue
    m = re.match(r"^(\$[^\s:]+):\s*$", ln.strip())
    if m:
        t = m.group(1)
        t = re.sub(r"^\$+_ZN3kid\w+?Ex\$", "", t)
        label = t
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        inst.append((cur[0], cur[1], label))

# function ranges of our own sources
ranges = {}
for f in os.listdir(csrc):
    if not f.endswith((".cuh", ".cu")):
        continue
    lines = open(os.path.join(csrc, f)).read().splitlines()
    starts = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"^(?:static\s+)?(?:__device__|__global__|inline|template).*?\b([A-Za-z_]\w*)\s*\(", l)
        if m and not l.startswith("template"):
            starts.append((i, m.group(1)))
        elif l.startswith("k_") and "(" in l:
            starts.append((i, l.split("(")[0]))
    ranges[f] = starts


def func_of(f, l):
    st = ranges.get(f)
    if not st:
        return f
    name = f
    for a, n in st:
        if a <= l:
            name = n
        else:
            break
    return name


rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:2 + len(inst)]
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0])
tot_i = tot_s = 0
nwarps = None
for k, r in enumerate(body):
    ex = int(r[ci["Instructions Executed"]] or 0)
    sm = int(r[ci["# Samples"]] or 0)
    if nwarps is None:
        nwarps = ex
    f, l, lab = inst[k]
    key = ("[sub] " + lab[:60]) if lab else func_of(f, l)
    a = agg[key]
    a[0] += ex; a[1] += sm; a[2] += 1
    a[3] += int(r[ci['stall_long_sb']] or 0); a[4] += int(r[ci['stall_wait']] or 0); a[5] += int(r[ci['stall_no_inst']] or 0); a[6] += int(r[ci['stall_short_sb']] or 0)
    tot_i += ex; tot_s += sm
print(f"static instr={len(inst)} executed warp-instr={tot_i} ({tot_i / nwarps:.0f} per warp, {nwarps} warps) samples={tot_s}")
print(f"{'function':62s} {'exec%':>6s} {'/warp':>7s} {'smp%':>6s} {'static':>6s} {'longsb%':>7s} {'wait%':>6s} {'noinst%':>7s} {'shortsb%':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] == 0 and a[2] < 50:
        continue
    try:
        print(f"{k:62s} {100 * a[0] / tot_i:6.1f} {a[0] / nwarps:7.0f} {100 * a[1] / max(tot_s, 1):6.1f} {a[2]:6d} {100 * a[3] / max(tot_s, 1):7.1f} {100 * a[4] / max(tot_s, 1):6.1f} {100 * a[5] / max(tot_s, 1):7.1f} {100 * a[6] / max(tot_s, 1):8.1f}")
    except BrokenPipeError:
        break
