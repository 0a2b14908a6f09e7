"""Prints the last N launches of an ncu --csv launch list with their durations: python tools/launch_table.py file.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr, L = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        L.setdefault((int(d["ID"]), d["Kernel Name"][:70]), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
tot = 0.0
for k in sorted(L)[-n:]:
    t = L[k]["gpu__time_duration.sum"] / 1000
    tot += t
    print(f"{k[0]:5d} {t:9.1f} us  {k[1]}")
print(f"total {tot:.1f} us")
