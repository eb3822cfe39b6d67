"""The fused path (spihtb_encode_images: pyramid base pass inside the forward transform's epilogue plus
the fix-up of straddling cells) must produce exactly the stream of the two-step path (spihtb_forward,
then spihtb_encode_coeffs with the stand-alone base pass over the finished array) -- on geometries
chosen to put band, chunk and strip boundaries at every parity: odd sizes, every wavelet and mode,
shallow and deep levels, more than one row chunk (> 64 band rows) and more than one strip.
"""
import numpy as np
import pytest

from conftest import synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def poison_cell_planes(monkeypatch):
    """the library fills the cell planes with 0xff before the fused pass: a cell that neither the epilogue,
    the gap fill nor the fix-up pass writes cannot pass on a stale (or zero) value"""
    monkeypatch.setenv("SPIHTB_DEBUG_POISON", "1")

CASES = [
    # (c, h, w), wavelet, mode, level, bpp
    ((3, 64, 96), "bior2.2", "reflect", None, 0.5),
    ((1, 61, 83), "bior2.2", "reflect", None, 1.0),
    ((3, 67, 131), "bior2.2", "reflect", 3, 0.7),
    ((2, 148, 140), "bior2.2", "reflect", 1, 0.5),        # one level: even LL so that the root offspring fit
    ((1, 301, 263), "bior2.2", "reflect", 2, 0.3),      # several row chunks and strips at level 1
    ((3, 257, 129), "bior2.2", "symmetric", None, 0.5),
    ((3, 128, 256), "bior2.2", "periodization", None, 0.5),
    ((1, 190, 70), "bior2.2", "periodization", 2, 1.0),
    ((2, 128, 96), "bior4.4", "symmetric", None, 0.5),
    ((3, 203, 169), "bior4.4", "reflect", 2, 0.4),
    ((1, 130, 66), "bior4.4", "periodization", 2, 1.0),
    ((1, 320, 333), "bior6.8", "reflect", None, 0.5),
    ((3, 96, 96), "bior6.8", "periodization", 1, 0.5),
    ((1, 277, 405), "bior6.8", "symmetric", 2, 0.25),
    # a cell that straddles two GAPS of different levels (the gap below the level-2 'ad' band starts at the odd column
    # 75, next to the level-3 gap): found by tools/fuzz_encode_paths.py, written by nobody before round 2's fix
    ((3, 45, 113), "bior6.8", "reflect", 4, 1.0),
    ((2, 113, 45), "bior6.8", "reflect", 4, 1.0),
]


@pytest.mark.parametrize("shape,wavelet,mode,level,bpp", CASES)
def test_fused_equals_two_step(shape, wavelet, mode, level, bpp):
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    c, h, w = shape
    imgs = np.stack([synth_image(c, h, w, 100 + s) for s in range(3)])
    px = torch.from_numpy(imgs).cuda()
    st = spiht.SpihtSettings(wavelet=wavelet, mode=mode)
    g = _lib.plan(h, w, wavelet, mode, level)
    mb = max(64, int(h * w * bpp))
    stride = batch.stream_stride(mb, c, g)
    s1, nbits1, n1, _, coeffs1 = batch.encode_images(px, g, st, mb, out_stride=stride,
                                                     out=torch.zeros((3, stride), dtype=torch.uint8, device="cuda"))
    coeffs2 = batch.forward(px, g, st)
    assert torch.equal(coeffs1, coeffs2)
    s2, nbits2, n2, _ = batch.encode_coeffs(coeffs2, g.ll_h, g.ll_w, mb, out_stride=stride,
                                            out=torch.zeros((3, stride), dtype=torch.uint8, device="cuda"))
    assert torch.equal(n1, n2), "max_n differs: the fused per-image maximum is wrong"
    assert torch.equal(nbits1, nbits2)
    for b in range(3):
        nb = (int(nbits1[b]) + 7) // 8
        assert torch.equal(s1[b, :nb], s2[b, :nb]), f"image {b}: fused and two-step streams differ"


def test_fused_untruncated():
    """no budget: every plane is coded, so every cell byte (and the per-image maximum) matters"""
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    c, h, w = 3, 93, 141
    px = torch.from_numpy(np.stack([synth_image(c, h, w, 7)])).cuda()
    st = spiht.SpihtSettings()
    g = _lib.plan(h, w, "bior2.2", "reflect", None)
    stride = batch.stream_stride(0, c, g)
    s1, nbits1, n1, _, coeffs = batch.encode_images(px, g, st, 0, out_stride=stride)
    s2, nbits2, n2, _ = batch.encode_coeffs(batch.forward(px, g, st), g.ll_h, g.ll_w, 0, out_stride=stride)
    assert torch.equal(n1, n2) and torch.equal(nbits1, nbits2)
    nb = (int(nbits1[0]) + 7) // 8
    assert torch.equal(s1[0, :nb], s2[0, :nb])
    # and the stream decodes to the array itself wherever the reference's coder reaches (it never visits
    # the last row / column of an odd-sized array, tests/test_oracle_spiht.py::test_kat3_odd_dims...)
    rec = batch.decode_coeffs(s1, (nbits1 + 7) // 8, n1, c, g.enc_h, g.enc_w, g.ll_h, g.ll_w)
    rec2 = batch.decode_coeffs(s2, (nbits2 + 7) // 8, n2, c, g.enc_h, g.enc_w, g.ll_h, g.ll_w)
    assert torch.equal(rec, rec2)


@pytest.mark.parametrize("shape,wavelet,mode,level,bpp", CASES + [((3, 1024, 1024), "bior2.2", "reflect", None, 0.5),
                                                                 ((1, 600, 840), "bior4.4", "reflect", None, 0.5),
                                                                 ((2, 1100, 900), "bior6.8", "symmetric", None, 0.3)])
def test_tail_kernel_equals_one_launch_per_level(monkeypatch, shape, wavelet, mode, level, bpp):
    """dwt_fwd_tail_kernel (all coarse levels of a plane in one CTA) against one launch per level: identical
    coefficient arrays, pyramid cells and per-image maxima (through the stream of an untruncated encode)"""
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    c, h, w = shape
    imgs = np.stack([synth_image(c, h, w, 200 + s) for s in range(2)]).astype(np.float32)
    px = torch.from_numpy(imgs).cuda()
    st = spiht.SpihtSettings(wavelet=wavelet, mode=mode)
    g = _lib.plan(h, w, wavelet, mode, level)
    stride = batch.stream_stride(0, c, g)

    def run():
        co = batch.forward(px, g, st)
        s, nbits, max_n, _, co2 = batch.encode_images(px, g, st, 0, out_stride=stride)
        torch.cuda.synchronize()
        return co, s, nbits, max_n, co2
    monkeypatch.delenv("SPIHTB_NO_TAIL", raising=False)
    monkeypatch.setenv("SPIHTB_TAIL_TASKS", "16")     # the tail kernel is opt-in (slower than one launch per level)
    a = run()
    monkeypatch.setenv("SPIHTB_NO_TAIL", "1")
    b = run()
    assert torch.equal(a[0], b[0]) and torch.equal(a[4], b[4])
    assert torch.equal(a[3], b[3]) and torch.equal(a[2], b[2])
    for i in range(2):
        nb = (int(a[2][i]) + 7) // 8
        assert torch.equal(a[1][i, :nb], b[1][i, :nb])


def test_tail_kernel_many_planes_vs_oracle(monkeypatch):
    """many planes running through their coarse levels at their own pace (more CTAs than the device holds at once):
    every coefficient array still equals the float64 oracle's -- a plane that read another plane's scratch would not"""
    import torch
    import spiht_b200 as spiht
    from oracle import wrapper_ref
    from spiht_b200 import _lib, batch
    B, c, h, w = 160, 3, 200, 264
    rng = np.random.default_rng(9)
    imgs = rng.random((B, c, h, w))
    px = torch.from_numpy(imgs).cuda()
    st = spiht.SpihtSettings()
    g = _lib.plan(h, w)
    monkeypatch.setenv("SPIHTB_TAIL_TASKS", "32")      # put as many levels as possible into the tail
    got = batch.forward(px, g, st).cpu().numpy()
    for b in (0, 1, 57, 101, 158, 159):
        of, _, _ = wrapper_ref.forward_coeffs(imgs[b], return_float=True)
        assert np.array_equal(got[b], of.astype(np.int32)), b
    monkeypatch.setenv("SPIHTB_NO_TAIL", "1")
    assert np.array_equal(batch.forward(px, g, st).cpu().numpy(), got)
