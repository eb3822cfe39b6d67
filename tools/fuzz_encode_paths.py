"""Randomised cross-check of the image encode path (run on a GPU box; not part of the test suite): spihtb_encode_images
(pyramid base pass fused into the transform, straddling cells fixed up, first ring hoisted onto the side stream) against
the two-step path (spihtb_forward, then spihtb_encode_coeffs with the stand-alone base pass) on random geometries,
wavelets, modes, levels and budgets; cell planes poisoned first.

    python tools/fuzz_encode_paths.py [--cases 100] [--seed 1] [--max-size 900]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["SPIHTB_DEBUG_POISON"] = "1"
import numpy as np
import torch
import spiht_b200 as spiht
from spiht_b200 import _lib, batch
from conftest import synth_image

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=100)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-size", type=int, default=900)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
bad = done = 0
for case in range(a.cases):
    c = int(rng.integers(1, 4))
    h, w = int(rng.integers(24, a.max_size)), int(rng.integers(24, a.max_size))
    wavelet = ["bior2.2", "bior2.2", "bior4.4", "bior6.8"][int(rng.integers(0, 4))]
    mode = ["reflect", "symmetric", "periodization"][int(rng.integers(0, 3))]
    level = None if rng.random() < 0.5 else int(rng.integers(1, 6))
    try:
        g = _lib.plan(h, w, wavelet, mode, level)
    except Exception:
        continue
    if min(g.ll_h, g.ll_w) < 2:
        continue
    B = 3
    st = spiht.SpihtSettings(wavelet=wavelet, mode=mode)
    px = torch.from_numpy(np.stack([synth_image(c, h, w, 5000 + case * 5 + s) for s in range(B)])).cuda()
    mb = 0 if rng.random() < 0.3 else max(64, int(h * w * rng.choice([0.05, 0.3, 1.0, 2.5])))
    stride = batch.stream_stride(mb, c, g)
    try:
        s1, nb1, n1, _, co1 = batch.encode_images(px, g, st, mb, out_stride=stride,
                                                  out=torch.zeros((B, stride), dtype=torch.uint8, device="cuda"))
    except Exception:
        continue
    co2 = batch.forward(px, g, st)
    s2, nb2, n2, _ = batch.encode_coeffs(co2, g.ll_h, g.ll_w, mb, out_stride=stride,
                                         out=torch.zeros((B, stride), dtype=torch.uint8, device="cuda"))
    done += 1
    ok = torch.equal(co1, co2) and torch.equal(n1, n2) and torch.equal(nb1, nb2)
    if ok:
        for b in range(B):
            nbytes = (int(nb1[b]) + 7) // 8
            ok = ok and torch.equal(s1[b, :nbytes], s2[b, :nbytes])
    if not ok:
        bad += 1
        print("MISMATCH", dict(case=case, c=c, h=h, w=w, wavelet=wavelet, mode=mode, level=level, mb=mb,
                               levels=g.levels, enc=(g.enc_h, g.enc_w)), flush=True)
print("cases run", done, "mismatches", bad)
sys.exit(1 if bad else 0)
