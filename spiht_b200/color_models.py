"""spiht/color_models.py of the reference.  Only RGB <-> IPT is supported: it is the one colour model on the
accelerated path (north star).  `convert(im, src, dest)` keeps the reference's signature (color_models.py:6-13,
CHW in, CHW out) and raises the reference's ValueError for every other model; the arithmetic runs on the GPU
(spihtb_convert_color).  Inside encode_image / decode_image the library runs the same kernels as a pass of its own
before the level-1 analysis / behind the level-1 synthesis (csrc/color.cu; not fused into the transform kernels,
DESIGN.md 4.8), without leaving the device.
"""
import ctypes

import numpy as np

SUPPORTED_MODELS = {"RGB", "IPT"}


def normalise(name):
    """None, or the canonical name of a supported model (case-insensitive, as colour.convert is)."""
    if name is None:
        return None
    up = str(name).upper()
    if up not in SUPPORTED_MODELS:
        raise ValueError(f'{name} is not a supported color model. Supported models are {SUPPORTED_MODELS}')
    return up


def convert(im, src: str, dest: str):
    """color_models.py:6-13: image (C,H,W) in colour model `src` -> `dest` (float64, CHW).
    numpy in, numpy out; a CUDA tensor in gives a CUDA tensor out."""
    import torch
    from . import _lib
    s, d = normalise(src), normalise(dest)
    is_tensor = isinstance(im, torch.Tensor)
    if not torch.cuda.is_available():
        raise RuntimeError("spiht_b200 needs a CUDA device: the codec has no CPU fallback")
    t = im if is_tensor else torch.from_numpy(np.ascontiguousarray(im))
    if t.ndim != 3 or t.shape[0] != 3:
        raise ValueError("convert expects a (3, H, W) image")
    if s == d:
        out = t.to(torch.float64)
        return out if is_tensor else out.numpy()
    if s == "RGB":
        if t.dtype not in (torch.float32, torch.float64, torch.uint8):
            t = t.to(torch.float64)
    else:
        t = t.to(torch.float64)
    t = t.cuda().contiguous()
    dev = t.device.index
    ctx = _lib.get_context(dev)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    out = torch.empty(t.shape, dtype=torch.float64, device=t.device)
    dt = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.uint8: _lib.U8}[t.dtype]
    ids = {"RGB": _lib.COLOR_NONE, "IPT": _lib.COLOR_IPT}
    _lib.check(_lib.lib().spihtb_convert_color(ctx.handle, ctypes.c_void_p(t.data_ptr()), dt, 1,
                                               t.shape[1] * t.shape[2], ids[s], ids[d],
                                               ctypes.c_void_p(out.data_ptr()), _lib.F64))
    return out if (is_tensor and im.is_cuda) else out.cpu().numpy() if not is_tensor else out.cpu()
