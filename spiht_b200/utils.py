"""spiht/utils.py of the reference: image loading and bit helpers (plotting left out)."""
import numpy as np


def bytes_to_bits(spiht_bytes: bytes):
    """spiht/utils.py:6-9"""
    np_bytes = np.frombuffer(spiht_bytes, np.uint8)
    return np.unpackbits(np_bytes, bitorder='little')


def imload(path) -> np.ndarray:
    """spiht/utils.py:12-20: PIL -> (C,H,W) float64 in [0,1]"""
    from PIL import Image
    im = np.asarray(Image.open(path))
    if im.ndim > 2:
        im = np.moveaxis(im, -1, 0)
    else:
        im = im[None, :, :]
    return im / 255


def synthetic_images(batch, channels, height, width, seed=0, device="cuda", dtype=None, chunk=16):
    """Synthetic natural-image-statistics batch: per image a 1/f-amplitude
    base field shared by the channels (0.8) plus an independent 1/f field per channel (0.2), each
    channel min-max normalised to [0,1].  Returns a float32 (B,C,H,W) tensor on `device`."""
    import torch
    dtype = dtype or torch.float32
    dev = torch.device(device)
    fy = torch.fft.fftfreq(height, device=dev)[:, None]
    fx = torch.fft.fftfreq(width, device=dev)[None, :]
    f = torch.sqrt(fy * fy + fx * fx)
    f[0, 0] = 1.0
    out = torch.empty((batch, channels, height, width), dtype=dtype, device=dev)
    gen = torch.Generator(device=dev)
    for s in range(0, batch, chunk):
        n = min(chunk, batch - s)
        gen.manual_seed(int(seed) * 1000003 + s)
        noise = torch.randn((n, channels + 1, height, width), generator=gen, device=dev, dtype=torch.float32)
        field = torch.fft.ifft2(torch.fft.fft2(noise) / f).real
        ch = 0.8 * field[:, :1] + 0.2 * field[:, 1:]
        lo = ch.amin(dim=(2, 3), keepdim=True)
        hi = ch.amax(dim=(2, 3), keepdim=True)
        out[s:s + n] = ((ch - lo) / (hi - lo)).to(dtype)
        del noise, field, ch
    return out
