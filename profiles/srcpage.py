"""Aggregate an exported ncu source page (ncu -i X.ncu-rep --page source --csv --print-source sass,cuda) by CUDA source line:
executed warp instructions and stall samples per line of svb_kernels.cuh.  Usage: python profiles/srcpage.py src.csv [top_n] [kernel substring]"""
import csv
import sys
import collections

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ""   # substring of the kernel name
rows = list(csv.reader(open(path)))
# locate header rows; the file holds one table per function/file
agg = collections.OrderedDict()
hdr = None
fn = None
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        fn = r[1][:60]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit() or (only and only not in (fn or "")):
        continue
    d = dict(zip(hdr, r))
    try:
        inst = int(d.get("Instructions Executed", "0") or 0)
        samp = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    key = (fn, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1][:110]])
    a[0] += inst
    a[1] += samp
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
for (f, line), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100 * s / tot_s:5.1f}% samp {100 * i / tot_i:5.1f}% inst  L{line:<5d} {src}")
