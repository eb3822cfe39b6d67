"""Opcode mix (warp-level instructions executed) of one kernel from an `ncu --page source --csv` export.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv
    python tools/sass_mix.py src.csv [rows]   # rows: divide by this many warp-rows to print per-row counts
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
per = float(sys.argv[2]) if len(sys.argv) > 2 else None
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # n-th kernel section of the export
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
starts.append(len(rows))
print(rows[starts[which]][1][:100] if rows[starts[which]][0] == "Kernel Name" else "")
rows = rows[starts[which]:starts[which + 1]]
hdr = next(r for r in rows if r and r[0] == "Address")
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
mix = collections.Counter()
tot = 0
for r in rows:
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    try:
        n = int(r[iex])
    except ValueError:
        continue
    src = r[isrc].split()
    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("F2", "I2", "LDG", "STG", "LDS", "STS", "SHFL")) and "." in op else "")
    mix[op] += n
    tot += n
print("total warp instructions", tot, ("= %.1f per row" % (tot / per)) if per else "")
for op, n in mix.most_common(40):
    print(f"{op:20s} {n:12d} {100 * n / tot:5.1f}%" + (f" {n / per:8.1f}/row" if per else ""))
