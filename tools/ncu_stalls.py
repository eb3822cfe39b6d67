"""Summarise the per-instruction stall samples of an `ncu --page source --csv` export."""
import csv, sys
path = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
kern = {}; k = None; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        k = r[1]; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if k and hdr and len(r) == len(hdr):
        kern.setdefault(k, {})
        d = dict(zip(hdr, r))
        kern[k].setdefault(d["Address"], d)
for k, dd in kern.items():
    if pat not in k:
        continue
    first = list(dd.values())
    tot = sum(int(d["# Samples"] or 0) for d in first)
    cols = [c for c in first[0] if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(d[c] or 0) for d in first) for c in cols}
    print(k[:90], "samples", tot, "instrs", len(first))
    print("  ", [(c[6:], v, round(100 * v / max(tot, 1))) for c, v in sorted(agg.items(), key=lambda x: -x[1])[:9]])
    for d in sorted(first, key=lambda d: -int(d["# Samples"] or 0))[:ntop]:
        st = sorted(((c[6:], int(d[c] or 0)) for c in cols), key=lambda x: -x[1])[:2]
        print("  ", d["Address"][-5:], d["# Samples"], d["Source"][:64], st)
