"""Stall samples per CUDA source line: joins an `ncu --page source --csv` export (SASS level) with the line table
of the same build (`nvdisasm -g` on the cubin extracted from the object file).

    cuobjdump -xelf all spiht_b200/csrc/_obj/spiht_enc.o && nvdisasm -g spiht_enc.sm_100a.cubin > enc.dis
    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv enc.dis spiht_encode_kernel spiht_enc.cu [top]
"""
import collections
import csv
import re
import sys

src_csv, dis, kern, fname = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
# line table: instruction offset -> (file, line) of the function whose name contains `kern`
line_of = {}
cur = None
infn = False
for ln in open(dis, errors="replace"):
    if ".text." in ln and ln.strip().startswith(".section"):
        infn = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and infn and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
sec = next((rows[a:b] for a, b in zip(starts, starts[1:]) if kern in rows[a][1]), rows)
hdr = next(r for r in sec if r and r[0] == "Address")
ia, isamp = hdr.index("Address"), hdr.index("# Samples")
stall_cols = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
base = None
agg = collections.defaultdict(lambda: [0, collections.Counter()])
for r in sec:
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    addr = int(r[ia], 16)
    base = addr if base is None else base
    key = line_of.get(addr - base, ("?", 0))
    a = agg[key]
    a[0] += int(r[isamp] or 0)
    for i, n in stall_cols:
        a[1][n] += int(r[i] or 0)
tot = sum(a[0] for a in agg.values())
srcl = {}
try:
    path = next(k[0] for k in agg if k[0].endswith(fname))
    srcl = {i + 1: l.rstrip() for i, l in enumerate(open(path))}
except Exception:
    pass
print("total samples", tot)
for (f, l), (n, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    code = srcl.get(l, "").strip()[:70] if f.endswith(fname) else f.split("/")[-1]
    print(f"{n:7d} {100 * n / tot:5.1f}%  {l:5d}  {code:70s} {st.most_common(2)}")
