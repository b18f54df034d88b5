#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --workload bonded --elements 432 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2z_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2z_launches.csv")))
hdr = None
agg = collections.OrderedDict(); seq = []
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum": continue
    name = d["Kernel Name"].split("(")[0][:60]
    t = float(d["Metric Value"]) / (1000.0 if d["Metric Unit"] in ("ns", "nsecond") else 1.0)
    seq.append((name, t))
print("launches", len(seq))
# the last step = the launches after the last k_mts_finish-but-one ... simply aggregate the last 178
last = seq[-178:]
a = collections.Counter(); c = collections.Counter()
for n, t in last: a[n] += t; c[n] += 1
for n, t in a.most_common(25): print("%-60s n=%3d  %.1f us" % (n, c[n], t))
print("sum us", sum(a.values()))
PY
