"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0]; v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'n':>4s} {'total us':>12s} {'share':>6s}")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:72]:72s} {a[0]:4d} {a[1]:12.1f} {100*a[1]/tot:5.1f}%")
print(f"{'TOTAL':72s} {sum(a[0] for a in agg.values()):4d} {tot:12.1f}")
