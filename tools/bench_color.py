"""Times the stand-alone RGB <-> IPT passes (spihtb_convert_color) on a batch of planar images."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spiht_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
B, S = a.batch, a.size
px = torch.rand((B, 3, S, S), dtype=torch.float32, device="cuda")
ipt = torch.empty((B, 3, S, S), dtype=torch.float64, device="cuda")
back = torch.empty((B, 3, S, S), dtype=torch.float32, device="cuda")
ctx = _lib.get_context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
L = _lib.lib()
def fwd():
    _lib.check(L.spihtb_convert_color(ctx.handle, ctypes.c_void_p(px.data_ptr()), _lib.F32, B, S * S, _lib.COLOR_NONE,
                                      _lib.COLOR_IPT, ctypes.c_void_p(ipt.data_ptr()), _lib.F64))
def inv():
    _lib.check(L.spihtb_convert_color(ctx.handle, ctypes.c_void_p(ipt.data_ptr()), _lib.F64, B, S * S, _lib.COLOR_IPT,
                                      _lib.COLOR_NONE, ctypes.c_void_p(back.data_ptr()), _lib.F32))
for f, name, bytes_px in ((fwd, "rgb_to_ipt f32->f64", 36), (inv, "ipt_to_rgb f64->f32", 36)):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"{name}: {ms:.3f} ms for {B} x 3 x {S}^2  = {B*S*S/ms/1e6:.1f} Gpixel/s, {B*S*S*bytes_px/ms/1e6:.0f} GB/s")
print("roundtrip max err", float((back - px).abs().max()))
