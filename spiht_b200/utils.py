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
