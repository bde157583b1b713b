"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 2:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0][:64]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches (ns, cold-cache serialized: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{t/tot*100:6.2f}%  n={n:4d}  avg={t/n/1e3:9.1f} us  {k}")
