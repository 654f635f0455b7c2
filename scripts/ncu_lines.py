"""Per-source-line summary of an ncu capture.
   ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass > dump.csv ; python scripts/ncu_lines.py dump.csv [top]
Prints, per CUDA source line: share of warp-stall samples, share of warp instructions, average active lanes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, fname, out = None, "", []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    g = lambda k: int(r[hdr.index(k)] or 0)
    out.append((fname, int(r[0]), r[1].strip()[:100], g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed")))
ts, ti, tt = (sum(o[k] for o in out) for k in (3, 4, 5))
print(f"total samples {ts}  warp-inst {ti}  thread-inst {tt}  lanes/inst {tt / max(ti, 1):.2f}")
byfile = {}
for o in out:
    a = byfile.setdefault(o[0], [0, 0, 0]); a[0] += o[3]; a[1] += o[4]; a[2] += o[5]
for f, a in byfile.items():
    print(f"  {f:18s} samples {100 * a[0] / ts:5.1f}%  warp-inst {100 * a[1] / ti:5.1f}%  lanes {a[2] / max(a[1], 1):5.1f}")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{100 * o[3] / ts:6.2f}% smp {100 * o[4] / ti:6.2f}% inst lanes {o[5] / max(o[4], 1):5.1f} | {o[0]}:{o[1]}: {o[2]}")
