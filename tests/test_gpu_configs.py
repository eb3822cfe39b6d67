"""BASELINE.json configurations at their full image sizes (small batches), on the GPU.

Parity at full size:
  * SPIHT: the CUDA bitstream must equal the CPU oracle's bitstream on the identical int32 coefficient array
    (the array the GPU forward transform produced), byte for byte, and the CUDA decoder must return the array
    the oracle decoder returns for those bytes;
  * float stages (configs 2, 3, 5): quantised coefficients against the float64 oracle with mismatches counted
    (rule of tests/test_gpu_transform.py);
  * size-independent properties everywhere: nbits == budget, decode(encode(x)) reproduces every coefficient the
    budget reached to within its last coded bit-plane, PSNR rises with bpp.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _mismatches(got, oracle_float):
    want = oracle_float.astype(np.int32)
    bad = np.nonzero(got != want)
    n_bad = len(bad[0])
    if n_bad:
        d = np.abs(got[bad].astype(np.int64) - want[bad])
        near = np.abs(oracle_float[bad] - np.round(oracle_float[bad]))
        assert d.max() <= 1 and near.max() < 1e-6
    return n_bad


def _spiht_parity(T, oracle, streams, nbits, max_n, coeffs, geom, max_bits, fast=False):
    from spiht_b200 import batch
    B, C = coeffs.shape[:2]
    nbytes = (nbits + 7) // 8
    rec = batch.decode_coeffs(streams, nbytes, max_n, C, geom.enc_h, geom.enc_w, geom.ll_h, geom.ll_w)
    enc = oracle.model_encode if fast else oracle.encode
    for b in range(B):
        arr = coeffs[b].cpu().numpy()
        want, want_n = enc(arr, geom.ll_h, geom.ll_w, max_bits)
        nb = int(nbytes[b])
        got = streams[b, :nb].cpu().numpy().tobytes()
        assert int(max_n[b]) == want_n
        assert int(nbits[b]) == min(max_bits, int(nbits[b]))
        assert got == want, f"image {b}: CUDA stream differs from the oracle"
        assert np.array_equal(rec[b].cpu().numpy(),
                              oracle.decode(want, want_n, C, geom.enc_h, geom.enc_w, geom.ll_h, geom.ll_w))


def _psnr(a, b):
    return float(10 * np.log10(1.0 / np.mean((a - b) ** 2)))


def test_config2_1024_bior22_reflect_half_bpp(T, oracle):
    """batch of synthetic 1024x1024 RGB images, bior2.2 reflect, 0.5 bpp (BASELINE.json configs[1])"""
    import spiht_b200 as spiht
    from oracle import wrapper_ref
    from spiht_b200 import _lib, batch
    from spiht_b200.utils import synthetic_images
    px = synthetic_images(3, 3, 1024, 1024, seed=2000)
    st = spiht.SpihtSettings()
    g = _lib.plan(1024, 1024)
    assert (g.levels, g.enc_h, g.enc_w, g.ll_h, g.ll_w) == (7, 1053, 1053, 12, 12)
    mb = int(1024 * 1024 * 0.5)
    s, nbits, max_n, status, coeffs = batch.encode_images(px, g, st, mb)
    assert int(status.max()) == 0 and int(nbits.min()) == mb
    _spiht_parity(T, oracle, s, nbits, max_n, coeffs, g, mb)
    of, _, _ = wrapper_ref.forward_coeffs(px[0].cpu().numpy().astype(np.float64), return_float=True)
    assert _mismatches(coeffs[0].cpu().numpy(), of) == 0
    out, _ = batch.decode_images(s, (nbits + 7) // 8, max_n, 3, g, st, dtype=T.float32)
    ref = wrapper_ref.decode_image(dict(encoded_bytes=s[0, :mb // 8].cpu().numpy().tobytes(), h=1024, w=1024, c=3,
                                        max_n=int(max_n[0]), level=None))
    assert np.abs(out[0].cpu().numpy() - ref).max() < 1e-5
    assert _psnr(out[0, :, :1024, :1024].cpu().numpy(), px[0].cpu().numpy()) > 20


def test_config3_2048_ipt_tenth_bpp(T, oracle):
    """2048x2048, IPT colour space with [50,15,15] quantisation scales, 0.1 bpp (configs[2])"""
    import spiht_b200 as spiht
    from oracle import wrapper_ref
    from spiht_b200 import _lib, batch
    from spiht_b200.utils import synthetic_images
    px = synthetic_images(2, 3, 2048, 2048, seed=3000)
    kw = dict(quantization_scale=1.0, color_model="IPT", per_channel_quant_scales=[50.0, 15.0, 15.0])
    st = spiht.SpihtSettings(**kw)
    g = _lib.plan(2048, 2048)
    assert (g.levels, g.enc_h, g.ll_h) == (8, 2081, 12)
    mb = int(2048 * 2048 * 0.1)
    s, nbits, max_n, status, coeffs = batch.encode_images(px, g, st, mb)
    assert int(nbits.min()) == mb
    _spiht_parity(T, oracle, s, nbits, max_n, coeffs, g, mb)
    of, _, _ = wrapper_ref.forward_coeffs(px[0].cpu().numpy().astype(np.float64), return_float=True, **kw)
    n_bad = _mismatches(coeffs[0].cpu().numpy(), of)
    print("config 3: quantised coefficient mismatches vs float64 oracle:", n_bad, "of", of.size)
    from conftest import record_count
    record_count("config3_2048_ipt_quantised_mismatches", mismatches=int(n_bad), coefficients=int(of.size),
                 max_abs_err_before_truncation=float(np.abs(coeffs[0].cpu().numpy() - of).max()))
    assert n_bad <= 1e-5 * of.size
    out, _ = batch.decode_images(s, (nbits + 7) // 8, max_n, 3, g, st, dtype=T.float32)
    assert _psnr(out[0, :, :2048, :2048].cpu().numpy(), px[0].cpu().numpy()) > 18


def test_config4_8192_bior68_one_bpp(T, oracle):
    """bior6.8 reflect, max decomposition level, 8192x8192 at 1.0 bpp (configs[3]: deep tree, large halo)"""
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    from spiht_b200.utils import synthetic_images
    px = synthetic_images(1, 3, 8192, 8192, seed=4000, chunk=1)
    st = spiht.SpihtSettings(wavelet="bior6.8")
    g = _lib.plan(8192, 8192, "bior6.8", "reflect")
    assert (g.levels, g.enc_h, g.ll_h) == (8, 8321, 48)
    mb = 8192 * 8192
    s, nbits, max_n, status, coeffs = batch.encode_images(px, g, st, mb)
    assert int(nbits[0]) == mb and int(status[0]) == 0
    # the pyramid formulation of the oracle (bit-identical to the faithful coder, tests/test_oracle_spiht.py)
    _spiht_parity(T, oracle, s, nbits, max_n, coeffs, g, mb, fast=True)
    out, rec = batch.decode_images(s, (nbits + 7) // 8, max_n, 3, g, st, dtype=T.float32)
    # every coefficient the budget reached is reproduced to within its last coded bit-plane
    err = (rec[0].to(T.int64) - coeffs[0].to(T.int64)).abs()
    coded = rec[0] != 0
    assert int(err[coded].max()) < int(coeffs[0].abs().max())
    assert _psnr(out[0, :, :8192, :8192].cpu().numpy(), px[0].cpu().numpy()) > 28


@pytest.mark.parametrize("bpp", [0.075, 0.1, 0.5, 1.0])
def test_config5_periodization_mixed_sizes(T, oracle, bpp):
    """bpp sweep plus periodization mode on a mixed-size batch (configs[4])"""
    import spiht_b200 as spiht
    from conftest import synth_image
    st = spiht.SpihtSettings(mode="periodization")
    sizes = [512, 1024, 2048, 512] if bpp != 1.0 else [512, 4096]
    imgs = [synth_image(3, n, n, 50 + i).astype(np.float32) for i, n in enumerate(sizes)]
    budgets = [int(n * n * bpp) for n in sizes]
    encs = [spiht.encode_images([im], st, max_bits=mb)[0] for im, mb in zip(imgs, budgets)]
    from spiht_b200 import _lib, batch
    for im, mb, e in zip(imgs, budgets, encs):
        n = im.shape[1]
        g = _lib.plan(n, n, "bior2.2", "periodization")
        assert (g.enc_h, g.enc_w) == (n, n)
        coeffs = batch.forward(T.from_numpy(im[None]).cuda(), g, st)[0].cpu().numpy()
        want, want_n = oracle.model_encode(coeffs, g.ll_h, g.ll_w, mb)
        assert (e.encoded_bytes, e.max_n) == (want, want_n)
        assert len(e.encoded_bytes) == (mb + 7) // 8
    recs = spiht.decode_images(encs, st)
    ps = [_psnr(r[:, :im.shape[1], :im.shape[2]], im) for r, im in zip(recs, imgs)]
    assert min(ps) > 15


@pytest.mark.parametrize("mode", ["reflect", "periodization"])
def test_config5_one_mixed_size_call(T, oracle, mode):
    """configs[4] as the config names it: ONE encode_images([...]) call over a mixed-size batch (512 .. 2048 px,
    two images per size so that equal shapes really are batched), every stream and every decoded image against
    the oracle; then one decode_images call over the mixed results"""
    import spiht_b200 as spiht
    from conftest import synth_image
    from oracle import wrapper_ref
    st = spiht.SpihtSettings(mode=mode)
    sizes = [512, 1024, 512, 2048, 1024, 768]
    imgs = [synth_image(3, n, n, 150 + i).astype(np.float32) for i, n in enumerate(sizes)]
    bpps = [0.075, 0.1, 0.5, 1.0, 0.1, 0.5]
    budgets = [int(n * n * b) for n, b in zip(sizes, bpps)]       # one budget per image, as the config's bpp sweep
    encs = spiht.encode_images(imgs, st, None, budgets)
    assert [(e.h, e.w) for e in encs] == [(n, n) for n in sizes]
    assert [len(e.encoded_bytes) for e in encs] == [(m + 7) // 8 for m in budgets]
    refs = [wrapper_ref.encode_image(im.astype(np.float64), mode=mode, max_bits=m) for im, m in zip(imgs, budgets)]
    for e, r in zip(encs, refs):
        assert e.max_n == r["max_n"] and e.encoded_bytes == r["encoded_bytes"]
    recs = spiht.decode_images(encs, st)
    for rec, r, im in zip(recs, refs, imgs):
        want = wrapper_ref.decode_image(r, mode=mode)
        assert rec.shape == want.shape and np.abs(rec - want).max() < 1e-9
        assert _psnr(rec[:, :im.shape[1], :im.shape[2]], im) > 15


def test_repeated_runs_are_bit_identical(T):
    """the coder's passes are scans and compactions over many warps, the forward transform forks a side stream and
    the decoder marks blocks with plain stores: repeated runs over the same batch must give the same bytes, the same
    decoded coefficients and the same pixels every time (a race would show up as a difference)"""
    torch = T
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    from spiht_b200.utils import synthetic_images
    B, S = 24, 256
    px = synthetic_images(B, 3, S, S, seed=77)
    st = spiht.SpihtSettings()
    g = _lib.plan(S, S, "bior2.2", "reflect", None)
    mb = int(S * S * 0.8)
    ref = None
    for it in range(12):
        s, nbits, max_n, _, coeffs = batch.encode_images(px, g, st, mb)
        nbytes = (nbits + 7) // 8
        out, rec = batch.decode_images(s, nbytes, max_n, 3, g, st, dtype=torch.float32)
        nb = int(nbytes.min())   # every image hits the budget: rows are written up to ceil(max_bits / 8) bytes
        assert int(nbytes.max()) == nb
        cur = (s[:, :nb].clone(), nbits.clone(), max_n.clone(), coeffs.clone(), rec.clone(), out.clone())
        if ref is None:
            ref = cur
        else:
            for a, b in zip(ref, cur):
                assert torch.equal(a, b), f"run {it} differs from run 0"


@pytest.mark.parametrize("group,streams", [(5, 3), (1, 2), (7, 4)])
def test_group_pipeline_equals_single_pass(T, monkeypatch, group, streams):
    """spihtb_encode_images cut into groups of images on alternating sub-contexts (own streams and workspaces, the
    coder of one group beside the transform of the next) gives exactly the single-pass result; repeated to let a
    workspace race show up as a difference"""
    torch = T
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    from spiht_b200.utils import synthetic_images
    B, S = 23, 256
    px = synthetic_images(B, 3, S, S, seed=91)
    st = spiht.SpihtSettings()
    g = _lib.plan(S, S, "bior2.2", "reflect", None)
    budgets = torch.tensor([int(S * S * (0.1 + 0.05 * (i % 7))) for i in range(B)], dtype=torch.int64, device="cuda")
    stride = batch.stream_stride(int(budgets.max()), 3, g)
    monkeypatch.delenv("SPIHTB_GROUP", raising=False)
    ref = batch.encode_images(px, g, st, budgets, out_stride=stride)
    torch.cuda.synchronize()
    monkeypatch.setenv("SPIHTB_GROUP", str(group))
    monkeypatch.setenv("SPIHTB_GROUP_STREAMS", str(streams))
    for it in range(4):
        got = batch.encode_images(px, g, st, budgets, out_stride=stride)
        torch.cuda.synchronize()
        assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]) and torch.equal(got[4], ref[4]), it
        for b in range(B):
            nb = (int(ref[1][b]) + 7) // 8
            assert torch.equal(got[0][b, :nb], ref[0][b, :nb]), (it, b)
