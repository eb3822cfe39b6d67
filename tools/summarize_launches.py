"""Per-kernel share table from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r[kn].split("(")[0].replace("void ", "").replace("spihtb::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name[:60]:60s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / tot:6.1f}%")
print(f"{'total':60s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}")
