"""Randomised cross-check of the decoder's alternative paths (run on a GPU box; not part of the test suite):
pipelined / serial LIS rounds, branch-free / round-1 walk, lazy / full zero fill, fused / level-by-level finest inverse
level must give identical pixels (and identical coefficient arrays where those are defined) on random geometries,
wavelets, modes, levels and prefixes.

    python tools/fuzz_decode_paths.py [--cases 60] [--seed 1]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import spiht_b200 as spiht
from spiht_b200 import _lib, batch
from conftest import synth_image

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=60)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-size", type=int, default=700)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
ENV = ("SPIHTB_DEC_PIPE", "SPIHTB_WALK", "SPIHTB_NO_LAZY_ZERO", "SPIHTB_LAZY_ZERO", "SPIHTB_NO_FUSED_INV", "SPIHTB_FUSED_INV_F64")


def setenv(**kw):
    for k in ENV:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = v


bad = 0
for case in range(a.cases):
    c = int(rng.integers(1, 4))
    h, w = int(rng.integers(40, a.max_size)), int(rng.integers(40, a.max_size))
    wavelet = ["bior2.2", "bior2.2", "bior4.4", "bior6.8"][int(rng.integers(0, 4))]
    mode = ["reflect", "symmetric", "periodization"][int(rng.integers(0, 3))]
    level = None if rng.random() < 0.5 else int(rng.integers(1, 5))
    try:
        g = _lib.plan(h, w, wavelet, mode, level)
    except Exception:
        continue
    if min(g.ll_h, g.ll_w) < 2:
        continue
    B = 5
    st = spiht.SpihtSettings(wavelet=wavelet, mode=mode)
    px = torch.from_numpy(np.stack([synth_image(c, h, w, 1000 + case * 7 + s) for s in range(B)])).cuda()
    setenv()
    try:
        s, nbits, max_n, _, _ = batch.encode_images(px, g, st, 0)
    except Exception as e:          # geometry the coder refuses (LL-root offspring outside the array)
        continue
    full = (nbits + 7) // 8
    frac = torch.from_numpy(rng.choice([0.0003, 0.002, 0.01, 0.05, 0.2, 0.6, 1.0], B)).cuda()
    nbytes = torch.clamp((full.double() * frac).long(), min=1)
    for dtype in (torch.float32, torch.float64):
        setenv(SPIHTB_DEC_PIPE="0", SPIHTB_WALK="0", SPIHTB_NO_LAZY_ZERO="1", SPIHTB_NO_FUSED_INV="1")
        ref, co = batch.decode_images(s, nbytes, max_n, c, g, st, dtype=dtype)
        setenv(SPIHTB_LAZY_ZERO="1", SPIHTB_FUSED_INV_F64="1")
        poison = torch.full((B, c, g.enc_h, g.enc_w), 0x7f7f7f7f, dtype=torch.int32, device="cuda")
        got, _ = batch.decode_images(s, nbytes, max_n, c, g, st, dtype=dtype, coeffs=poison, scratch_coeffs=True)
        setenv()
        got2, co2 = batch.decode_images(s, nbytes, max_n, c, g, st, dtype=dtype)
        ok = torch.equal(ref, got) and torch.equal(ref, got2) and torch.equal(co, co2)
        if not ok:
            bad += 1
            print("MISMATCH", dict(case=case, c=c, h=h, w=w, wavelet=wavelet, mode=mode, level=level, dtype=str(dtype),
                                   ll=(g.ll_h, g.ll_w), nbytes=nbytes.tolist(),
                                   px_all=int((ref != got).sum()), px_default=int((ref != got2).sum()),
                                   coeffs=int((co != co2).sum())), flush=True)
print("cases", a.cases, "mismatches", bad)
sys.exit(1 if bad else 0)
