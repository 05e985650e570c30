"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv):  python tools/launch_summary.py file.csv [last_n_launches]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
recs = []
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    recs.append((r[ki].split("(")[0], v))
if len(sys.argv) > 2:
    recs = recs[-int(sys.argv[2]):]
tot = sum(v for _, v in recs)
agg = collections.OrderedDict()
for k, v in recs:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
print(f"{len(recs)} launches, {tot:.0f} us")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {v:9.0f} us {100 * v / tot:5.1f} %  x{n:<4d} {k}")
