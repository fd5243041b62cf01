"""Per-kernel durations from an `ncu --metrics gpu__time_duration.sum --csv` log: python scripts/launch_times.py log.csv [filter]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(r for r in rows if "Kernel Name" in r)
i0 = rows.index(hdr)
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
flt = sys.argv[2] if len(sys.argv) > 2 else "k_"
agg = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) > mv and flt in r[kn]:
        name = r[kn].split("(")[0]
        agg.setdefault(name, []).append(float(r[mv].replace(",", "")))
for k, v in agg.items():
    print(f"{k[:70]:70s} n={len(v):3d} avg={sum(v) / len(v) / 1e3:9.2f} us  total={sum(v) / 1e3:10.2f} us")
