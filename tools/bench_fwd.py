"""Times the forward transform + pyramid + coder stages of one encode step (stage timers), for kernel work.

    python tools/bench_fwd.py [--batch 256] [--size 1024] [--steps 5] [--dtype f32]
Environment switches read by the library (SPIHTB_NO_FUSED12, SPIHTB_F12_CHUNKS, ...) apply.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spiht_b200 as spiht
from spiht_b200 import _lib, batch
from spiht_b200.utils import synthetic_images

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--bpp", type=float, default=0.5)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--wavelet", default="bior2.2")
ap.add_argument("--mode", default="reflect")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--tag", default="")
a = ap.parse_args()
st = spiht.SpihtSettings(wavelet=a.wavelet, mode=a.mode)
g = _lib.plan(a.size, a.size, a.wavelet, a.mode)
px = synthetic_images(a.batch, 3, a.size, a.size, seed=7)
if a.dtype == "u8":
    px = (px * 255).round().to(torch.uint8)
elif a.dtype == "f64":
    px = px.double()
mb = int(a.size * a.size * a.bpp)
stride = batch.stream_stride(mb, 3, g)
coeffs = torch.empty((a.batch, 3, g.enc_h, g.enc_w), dtype=torch.int32, device="cuda")
out = torch.zeros((a.batch, stride), dtype=torch.uint8, device="cuda")
ctx = _lib.get_context(0)
for _ in range(3):
    batch.encode_images(px, g, st, mb, out_stride=stride, coeffs=coeffs, out=out)
torch.cuda.synchronize()
ctx.profile(True)
ctx.profile_read(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    batch.encode_images(px, g, st, mb, out_stride=stride, coeffs=coeffs, out=out)
e1.record()
torch.cuda.synchronize()
stages = ctx.profile_read(reset=True)
ctx.profile(False)
print(json.dumps({"tag": a.tag, "path": ctx.forward_path(), "step_ms": round(e0.elapsed_time(e1) / a.steps, 4),
                  "stages_ms": {k: round(v[0] / a.steps, 4) for k, v in stages.items() if v[1]},
                  "env": {k: v for k, v in os.environ.items() if k.startswith("SPIHTB_")}}), flush=True)
