import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import spiht_oracle as oracle
from spiht_b200 import spiht as rs
rng = np.random.default_rng(11)
for t in range(120):
    c = int(rng.integers(1, 5)); h = int(rng.integers(4, 70)); w = int(rng.integers(4, 70))
    llh = int(rng.integers(2, max(3, h // 2 + 1))); llw = int(rng.integers(2, max(3, w // 2 + 1)))
    if not oracle.geom_ok(h, w, llh, llw):
        continue
    x = rng.normal(0, 16 * rng.random() ** 2 * 40 + 0.5, (c, h, w)).astype(np.int32)
    if rng.random() < 0.08:
        x[:] = 0
    mb = int(rng.integers(0, 6000)) if rng.random() < 0.7 else 10 ** 9
    want, want_n = oracle.encode(x, llh, llw, mb)
    got, got_n = rs.encode(x, llh, llw, mb)
    ok_enc = (got == want and got_n == want_n)
    rec = rs.decode(got, got_n, c, h, w, llh, llw)
    ref = oracle.decode(got, got_n, c, h, w, llh, llw)
    bad = np.argwhere(rec != ref)
    if not ok_enc or len(bad):
        print(f"case {t}: c={c} h={h} w={w} ll={llh}x{llw} mb={mb} nbytes={len(got)} n={got_n} enc_ok={ok_enc} nbad={len(bad)}")
        for b in bad[:12]:
            k, i, j = b
            print("   ", (int(k), int(i), int(j)), "gpu", int(rec[k, i, j]), "ref", int(ref[k, i, j]), "orig", int(x[k, i, j]))
