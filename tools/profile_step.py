"""One encode step + one decode step of the bench workload at a reduced batch, for ncu captures."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spiht_b200 as spiht
from spiht_b200 import _lib, batch
from spiht_b200.utils import synthetic_images

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--bpp", type=float, default=0.5)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--wavelet", default="bior2.2")
ap.add_argument("--mode", default="reflect")
ap.add_argument("--color", default=None)
a = ap.parse_args()
st = spiht.SpihtSettings(wavelet=a.wavelet, mode=a.mode, color_model=a.color)
g = _lib.plan(a.size, a.size, a.wavelet, a.mode)
px = synthetic_images(a.batch, 3, a.size, a.size, seed=7)
mb = int(a.size * a.size * a.bpp)
for _ in range(a.steps):
    s, nbits, max_n, status, coeffs = batch.encode_images(px, g, st, mb)
    out, _ = batch.decode_images(s, (nbits + 7) // 8, max_n, 3, g, st, dtype=torch.float32)
torch.cuda.synchronize()
print("ok", int(nbits.min()), float(((out[:, :, :a.size, :a.size] - px) ** 2).mean()))
