"""The fused two-level forward kernel (csrc/dwt_fwd2.cu: TMA-staged tiles, level-1 approximation kept in shared
memory) against the level-by-level kernel (csrc/dwt_fwd.cu) on the same inputs: bit-identical coefficient arrays,
pyramid cells (through the stream of an untruncated encode with the cell planes poisoned) and per-image maxima.
The level-by-level path is itself checked against the float64 oracle in tests/test_gpu_transform.py, and
both paths are checked against the oracle at the BASELINE sizes in tests/test_gpu_configs.py.

Geometries put strip, row-chunk, band and plane boundaries at every parity: several strips (level-2 band wider
than 60 columns), several row chunks (SPIHTB_F12_CHUNKS), odd heights, all wavelets / modes / pixel dtypes.
"""
import os

import numpy as np
import pytest

from conftest import synth_image

pytestmark = pytest.mark.gpu

CASES = [
    # (c, h, w), dtype, wavelet, mode, level, chunks
    ((3, 64, 96), "f64", "bior2.2", "reflect", None, None),
    ((3, 256, 384), "f32", "bior2.2", "reflect", None, None),        # config-1 shape
    ((1, 301, 264), "f32", "bior2.2", "reflect", 3, 3),               # odd height, two strips, three row chunks
    ((2, 517, 1000), "f32", "bior2.2", "reflect", None, 5),           # five strips
    ((3, 1024, 1024), "f32", "bior2.2", "reflect", None, None),      # config-2 shape
    ((1, 1024, 1024), "u8", "bior2.2", "reflect", None, None),
    ((3, 257, 132), "f32", "bior2.2", "symmetric", None, 2),
    ((3, 128, 256), "f32", "bior2.2", "periodization", None, None),
    ((2, 512, 512), "f32", "bior2.2", "periodization", None, 4),      # config-5 shape
    ((1, 192, 72), "f64", "bior2.2", "periodization", 3, 2),
    ((1, 336, 304), "u8", "bior2.2", "periodization", 3, None),
    ((2, 256, 192), "f32", "bior4.4", "symmetric", None, None),
    ((3, 403, 368), "f32", "bior4.4", "reflect", 3, 3),
    ((1, 520, 528), "f64", "bior4.4", "periodization", 4, 2),
    ((1, 640, 648), "f32", "bior6.8", "reflect", None, 2),
    ((2, 768, 512), "u8", "bior6.8", "periodization", 3, None),
    ((1, 577, 808), "f32", "bior6.8", "symmetric", 3, 3),
    ((1, 2048, 2048), "f32", "bior2.2", "reflect", None, None),      # config-3 shape
]


def _pixels(shape, dtype, seed):
    import torch
    c, h, w = shape
    imgs = np.stack([synth_image(c, h, w, seed + s) for s in range(2)])
    if dtype == "u8":
        return torch.from_numpy(np.round(imgs * 255).astype(np.uint8)).cuda()
    return torch.from_numpy(imgs.astype(np.float32 if dtype == "f32" else np.float64)).cuda()


@pytest.mark.parametrize("shape,dtype,wavelet,mode,level,chunks", CASES)
def test_fused12_equals_level_by_level(monkeypatch, shape, dtype, wavelet, mode, level, chunks):
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    monkeypatch.setenv("SPIHTB_DEBUG_POISON", "1")
    c, h, w = shape
    px = _pixels(shape, dtype, 300)
    st = spiht.SpihtSettings(wavelet=wavelet, mode=mode, quantization_scale=50.0 if c != 2 else 20.0,
                             per_channel_quant_scales=[1.0, 0.5][:c] if c == 2 else None)
    g = _lib.plan(h, w, wavelet, mode, level)
    ctx = _lib.get_context(0)
    stride = batch.stream_stride(0, c, g)

    def run():
        co = batch.forward(px, g, st)
        path_fwd = ctx.forward_path()
        s, nbits, max_n, status, co2 = batch.encode_images(px, g, st, 0, out_stride=stride)
        path_enc = ctx.forward_path()
        torch.cuda.synchronize()
        return co, path_fwd, s, nbits, max_n, co2, path_enc

    monkeypatch.setenv("SPIHTB_FUSED12", "1")     # the fused kernel is opt-in (slower than the level-by-level pair)
    if chunks:
        monkeypatch.setenv("SPIHTB_F12_CHUNKS", str(chunks))
    a = run()
    assert a[1] == 12 and a[6] == 12, "the fused two-level kernel did not run on this geometry"
    monkeypatch.setenv("SPIHTB_FUSED12", "0")
    b = run()
    assert b[1] == 1 and b[6] == 1
    bad = (a[0] != b[0])
    assert not bool(bad.any()), f"{int(bad.sum())} coefficients differ, first at {bad.nonzero()[0].tolist()}"
    assert torch.equal(a[5], b[5])
    assert torch.equal(a[4], b[4]), "max_n differs: the per-image maximum of the fused kernel is wrong"
    assert torch.equal(a[3], b[3])
    for i in range(px.shape[0]):
        nb = (int(a[3][i]) + 7) // 8
        assert torch.equal(a[2][i, :nb], b[2][i, :nb]), f"image {i}: streams differ (a pyramid cell is wrong)"


def test_fused12_falls_back_on_unsupported_geometry(monkeypatch):
    """row strides the TMA unit cannot address (not a multiple of 16 bytes), fewer than three levels, planes only a
    few filter lengths wide: the level-by-level kernel runs and the results stay correct (oracle-checked elsewhere)"""
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    ctx = _lib.get_context(0)
    monkeypatch.setenv("SPIHTB_FUSED12", "1")
    for shape, mode, level in [((1, 61, 83), "reflect", None), ((1, 148, 140), "reflect", 2),
                               ((3, 20, 24), "reflect", None),
                               # periodization pads an odd level-1 band before level 2: not the periodic halo
                               ((1, 190, 72), "periodization", 3)]:
        c, h, w = shape
        px = torch.from_numpy(synth_image(c, h, w, 1)[None].astype(np.float32)).cuda()
        g = _lib.plan(h, w, "bior2.2", mode, level)
        batch.forward(px, g, spiht.SpihtSettings(mode=mode))
        assert ctx.forward_path() == 1
