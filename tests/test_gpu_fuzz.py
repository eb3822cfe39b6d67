"""Randomised geometries against the CPU oracle, end to end through the image path (fixed seeds, small sizes so that the
numpy oracle finishes in seconds): forward coefficients (counted off-by-one ties only), encode_images streams == the
oracle's coder on the GPU's own coefficients, decode_images == the oracle's decoder + inverse.  The fixed parametrised
cases elsewhere pick geometries by hand; these pick them blindly (odd sizes, non-square images, deep levels on small
images, every wavelet and mode) -- tools/fuzz_encode_paths.py found a cell-plane hole this way."""
import numpy as np
import pytest

from conftest import synth_image

pytestmark = pytest.mark.gpu


def _cases(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        c = int(rng.integers(1, 4))
        h, w = int(rng.integers(24, 260)), int(rng.integers(24, 260))
        wavelet = ["bior2.2", "bior4.4", "bior6.8"][int(rng.integers(0, 3))]
        mode = ["reflect", "symmetric", "periodization"][int(rng.integers(0, 3))]
        level = None if rng.random() < 0.4 else int(rng.integers(1, 6))
        bpp = float(rng.choice([0.0, 0.1, 0.6, 2.0]))
        out.append((c, h, w, wavelet, mode, level, bpp))
    return out


@pytest.mark.parametrize("case", _cases(2024, 48))
def test_random_geometry_end_to_end(oracle, case):
    import torch
    import spiht_b200 as spiht
    from oracle import wrapper_ref
    from spiht_b200 import _lib, batch
    c, h, w, wavelet, mode, level, bpp = case
    try:
        g = _lib.plan(h, w, wavelet, mode, level)
    except Exception:
        pytest.skip("level too deep for this size")
    if min(g.ll_h, g.ll_w) < 2:
        pytest.skip("LL band too small for the coder")
    st = spiht.SpihtSettings(wavelet=wavelet, mode=mode)
    img = synth_image(c, h, w, 11 * h + w)
    px = torch.from_numpy(img[None]).cuda()
    mb = 0 if bpp == 0.0 else max(64, int(h * w * bpp))
    stride = batch.stream_stride(mb, c, g)
    try:
        s, nbits, max_n, _, co = batch.encode_images(px, g, st, mb, out_stride=stride)
    except _lib.SpihtB200Error as e:
        assert e.code == _lib.EGEOM      # LL-root offspring outside the array: the reference panics there too
        pytest.skip("geometry the coder refuses")
    co_h = co[0].cpu().numpy()
    # forward: the float64 oracle, ties counted
    of, ll_h, ll_w = wrapper_ref.forward_coeffs(img, wavelet, mode, level, 50.0, return_float=True)
    assert (ll_h, ll_w) == (g.ll_h, g.ll_w) and of.shape == co_h.shape
    ref_q = of.astype(np.int32)
    bad = np.argwhere(co_h != ref_q)
    assert len(bad) <= max(2, co_h.size // 1000), len(bad)
    for t in bad:
        t = tuple(t)
        assert abs(int(co_h[t]) - int(ref_q[t])) == 1 and abs(of[t] - round(of[t])) < 1e-9
    # coder: the oracle on the GPU's own coefficients, bit for bit
    want, want_n = oracle.model_encode(co_h, g.ll_h, g.ll_w, mb if mb else 10 ** 9)
    nby = (int(nbits[0]) + 7) // 8
    assert int(max_n[0]) == want_n
    assert s[0, :nby].cpu().numpy().tobytes() == want
    # decoder + inverse: a prefix and the whole stream
    for L in sorted({max(1, nby // 3), nby}):
        nb = torch.tensor([L], dtype=torch.int64)
        pix, rec = batch.decode_images(s, nb, max_n, c, g, st, dtype=torch.float64)
        ref_rec = oracle.decode(want[:L], want_n, c, g.enc_h, g.enc_w, g.ll_h, g.ll_w)
        assert np.array_equal(rec[0].cpu().numpy(), ref_rec)
        ref_pix = wrapper_ref.inverse_coeffs(ref_rec, h, w, wavelet, mode, level, 50.0)
        got = pix[0].cpu().numpy()[:, :ref_pix.shape[1], :ref_pix.shape[2]]
        assert np.abs(got - ref_pix).max() <= 1e-9 * max(1.0, np.abs(ref_pix).max())
        pix2, _ = batch.decode_images(s, nb, max_n, c, g, st, dtype=torch.float32, scratch_coeffs=True)
        assert np.abs(pix2[0].cpu().numpy()[:, :ref_pix.shape[1], :ref_pix.shape[2]] - ref_pix).max() <= 1e-6 * max(1.0, np.abs(ref_pix).max())
