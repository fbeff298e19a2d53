#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: where the warp-stall samples of a kernel fall.
usage: ncu_hot.py <source.csv> [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
tot = sum(int(d["# Samples"]) for d in data)
print("instructions", len(data), "total samples", tot)
# cumulative profile along the instruction stream, in 5% chunks of the address range
n = len(data)
acc = 0
marks = []
for i, d in enumerate(data):
    acc += int(d["# Samples"])
    if (i + 1) % max(1, n // 20) == 0:
        marks.append((i + 1, round(100 * acc / tot, 1)))
print("cumulative % of samples by instruction index:", marks)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.Counter()
for d in data:
    for k in stalls:
        agg[k] += int(d[k] or 0)
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in agg.most_common(10)))
by_op = collections.Counter()
for d in data:
    op = d["Source"].split()[0] if not d["Source"].strip().startswith("@") else d["Source"].split()[1]
    by_op[op.split(".")[0]] += int(d["# Samples"])
print("samples by opcode:", ", ".join(f"{k}={v}" for k, v in by_op.most_common(15)))
print("top instructions:")
for i, d in sorted(enumerate(data), key=lambda x: -int(x[1]["# Samples"]))[:top]:
    s = {k[6:]: int(d[k] or 0) for k in stalls if int(d[k] or 0) > 0}
    s = dict(sorted(s.items(), key=lambda x: -x[1])[:3])
    print(f"  #{i:5d} {int(d['# Samples']):6d}  {d['Source'].strip()[:70]:70s} {s}")
