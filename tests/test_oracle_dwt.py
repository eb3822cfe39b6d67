"""CPU checks of the float64 DWT / IPT oracle (PyWavelets and colour-science are
absent from this image: filter identities, perfect reconstruction and band
geometry are what can be pinned)."""
import numpy as np
import pytest

from oracle import dwt_ref as d, ipt_ref, wrapper_ref


@pytest.mark.parametrize("name", d.WAVELETS)
def test_filter_identities(name):
    wv = d.Wavelet(name)
    assert abs(wv.dec_lo.sum() - 2 ** 0.5) < 1e-12
    assert abs(wv.rec_lo.sum() - 2 ** 0.5) < 1e-12
    assert abs(wv.dec_hi.sum()) < 1e-11
    assert abs(wv.rec_hi.sum()) < 1e-11


@pytest.mark.parametrize("name", d.WAVELETS)
@pytest.mark.parametrize("mode", d.MODES)
def test_perfect_reconstruction_1d(name, mode):
    rng = np.random.default_rng(0)
    wv = d.Wavelet(name)
    for n in (16, 17, 40, 41, 64, 100, 7):
        x = rng.normal(size=(2, n))
        lo, hi = d.dwt_axis(x, wv, mode, -1)
        assert lo.shape[-1] == d.dwt_coeff_len(n, wv.dec_len, mode)
        r = d.idwt_axis(lo, hi, wv, mode, -1)
        assert np.abs(r[:, :n] - x).max() < 1e-10


def test_geometry_of_the_configs():
    """coefficient-array sizes of the BASELINE configs"""
    assert d.get_slices_and_h_w(256, 384, "bior2.2", "reflect", None)[1:] == (277, 405)
    assert d.wavedecn_shapes_2d(256, 384, "bior2.2")[:2] == (12, 16)
    assert d.get_slices_and_h_w(1024, 1024, "bior2.2", "reflect", None)[1:] == (1053, 1053)
    assert d.get_slices_and_h_w(2048, 2048, "bior2.2", "reflect", None)[1:] == (2081, 2081)
    assert d.get_slices_and_h_w(8192, 8192, "bior6.8", "reflect", None)[1:] == (8321, 8321)
    assert d.wavedecn_shapes_2d(8192, 8192, "bior6.8")[:2] == (48, 48)
    assert d.get_slices_and_h_w(1024, 1024, "bior2.2", "periodization", None)[1:] == (1024, 1024)


@pytest.mark.parametrize("mode", d.MODES)
@pytest.mark.parametrize("name", ["bior2.2", "bior4.4"])
def test_wavedec2_roundtrip_and_layout(name, mode):
    rng = np.random.default_rng(1)
    for (h, w) in [(64, 96), (67, 131)]:
        img = rng.random((3, h, w))
        co = d.wavedec2(img, name, mode)
        arr = d.coeffs_to_array(co)
        sl, eh, ew = d.get_slices_and_h_w(h, w, name, mode, None)
        assert arr.shape[1:] == (eh, ew)
        back = d.array_to_coeffs(arr, sl)
        for a, b in zip(co[1:], back[1:]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        rec = d.waverec2(back, name, mode)
        assert np.abs(rec[:, :h, :w] - img).max() < 1e-10


def test_ipt_roundtrip_close():
    x = np.random.default_rng(2).random((3, 16, 16))
    y = ipt_ref.convert(x, "RGB", "IPT")
    z = ipt_ref.convert(y, "IPT", "RGB")
    assert np.abs(z - x).max() < 1e-12     # every matrix of the way back is the float64 inverse of the forward one
    with pytest.raises(ValueError):
        ipt_ref.convert(x, "RGB", "CIE Lab")


def test_wrapper_ref_roundtrip():
    from conftest import synth_image
    img = synth_image(3, 64, 96, 0)
    enc = wrapper_ref.encode_image(img, max_bits=64 * 96)
    assert len(enc["encoded_bytes"]) == 64 * 96 // 8
    rec = wrapper_ref.decode_image(enc)
    assert rec.shape == img.shape
    mse1 = np.mean((rec - img) ** 2)
    rec_full = wrapper_ref.decode_image(wrapper_ref.encode_image(img))
    mse_full = np.mean((rec_full - img) ** 2)
    assert mse_full < 1e-3 and mse_full < mse1 < 0.05


@pytest.mark.parametrize("shape,wavelet,mode,level", [
    ((3, 64, 96), "bior2.2", "reflect", None),
    ((1, 67, 131), "bior2.2", "reflect", 3),
    ((2, 70, 70), "bior2.2", "symmetric", None),
    ((3, 128, 64), "bior2.2", "periodization", None),
    ((1, 131, 66), "bior4.4", "periodization", 2),
    ((2, 128, 96), "bior4.4", "symmetric", None),
    ((1, 200, 168), "bior6.8", "reflect", 2),
])
def test_compiled_transform_of_the_cpu_baseline_matches_the_checker(shape, wavelet, mode, level):
    """oracle/dwt_fast.c (timed by bench.py's CPU legs) == oracle/dwt_ref.py (the checker) to rounding"""
    from oracle import dwt_fast, dwt_ref, wrapper_ref
    rng = np.random.default_rng(11)
    x = rng.normal(size=shape)
    a = dwt_ref.wavedec2(x, wavelet, mode, level)
    b = dwt_fast.wavedec2(x, wavelet, mode, level)
    assert len(a) == len(b)
    assert np.abs(a[0] - b[0]).max() <= 1e-12
    for ta, tb in zip(a[1:], b[1:]):
        for u, v in zip(ta, tb):
            assert u.shape == v.shape and np.abs(u - v).max() <= 1e-12
    assert np.abs(dwt_ref.waverec2(a, wavelet, mode) - dwt_fast.waverec2(a, wavelet, mode)).max() <= 1e-12
    c1, lh, lw = wrapper_ref.forward_coeffs(x, wavelet, mode, level)
    c2, lh2, lw2 = wrapper_ref.forward_coeffs(x, wavelet, mode, level, fast=True)
    assert (lh, lw) == (lh2, lw2) and c1.shape == c2.shape and np.abs(c1 - c2).max() <= 1


# ---- independent derivations of the recalled filter tables -------------------------------------------------
# oracle/dwt_ref.py restates PyWavelets' bior tables from memory.  bior2.2 and bior4.4 are the CDF 5/3 and 9/7
# pairs of JPEG 2000, whose *lifting* factorisations (ITU-T T.800 Annex F constants; Daubechies & Sweldens
# 1998) share nothing with those tables: running the lifting steps must reproduce dwt_axis on interior samples
# up to the sqrt(2) normalisation.  Alignment of the PyWavelets convention out[k] = sum_j f[j] x[2k+1-j]:
# 5/3: lo[k] = sqrt(2) s[k-1], hi[k] = -d[k-1] / sqrt(2);  9/7: lo[k] = zeta s[k-2], hi[k] = -d[k-2] / zeta.
def _lift53(x):
    e, o = x[0::2].copy(), x[1::2].copy()
    o -= 0.5 * (e + np.roll(e, -1))          # d[n] = x[2n+1] - (x[2n] + x[2n+2]) / 2
    e += 0.25 * (np.roll(o, 1) + o)          # s[n] = x[2n] + (d[n-1] + d[n]) / 4
    return 2 ** 0.5 * e, o / 2 ** 0.5


def _lift97(x):
    alpha, beta, gamma, delta = -1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971
    zeta = 1.149604398860241                  # = sqrt(2) / K, K = 1.230174104914001 (T.800 Table F.4)
    e, o = x[0::2].copy(), x[1::2].copy()
    o += alpha * (e + np.roll(e, -1))
    e += beta * (np.roll(o, 1) + o)
    o += gamma * (e + np.roll(e, -1))
    e += delta * (np.roll(o, 1) + o)
    return zeta * e, o / zeta


@pytest.mark.parametrize("name,lift,shift,tol", [("bior2.2", _lift53, 1, 1e-14), ("bior4.4", _lift97, 2, 1e-11)])
@pytest.mark.parametrize("mode", ["reflect", "symmetric"])
def test_lifting_factorisation_reproduces_the_filter_bank(name, lift, shift, tol, mode):
    rng = np.random.default_rng(5)
    x = rng.normal(size=128)
    lo, hi = d.dwt_axis(x[None], d.Wavelet(name), mode, -1)
    s, dd = lift(x)
    idx = np.arange(8, len(s) - 8)            # interior: away from the boundary extension and the lifting's wrap
    assert np.abs(lo[0][idx + shift] - s[idx]).max() < tol
    assert np.abs(hi[0][idx + shift] + dd[idx]).max() < tol


@pytest.mark.parametrize("name,nr,nd", [(n, int(n[4]), int(n[6])) if n != "bior5.5" else (n, 6, 4) for n in d.WAVELETS])
def test_filter_tables_have_the_vanishing_moments_their_name_states(name, nr, nd):
    """biorNr.Nd: dec_lo has Nd zeros at z = -1 and rec_lo has Nr (so that the decomposition / reconstruction
    wavelets have Nd / Nr vanishing moments) -- and not one more.  Together with perfect reconstruction,
    symmetry and the tap counts this determines the pair; it uses no PyWavelets-specific knowledge."""
    wv = d.Wavelet(name)
    for f, nz in ((wv.dec_lo, nd), (wv.rec_lo, nr)):
        k = np.arange(len(f), dtype=np.float64)
        k -= k[np.abs(f) > 0].mean()          # centre: keeps the high moments well conditioned
        alt = (-1.0) ** np.arange(len(f))
        moms = [abs(np.sum(alt * k ** p * f)) for p in range(nz + 1)]
        assert max(moms[:nz]) < 1e-9, moms
        assert moms[nz] > 1e-2, moms
        assert np.allclose(f[np.abs(f) > 0], f[np.abs(f) > 0][::-1], atol=0, rtol=0)   # exactly symmetric


def test_derived_spline_pair_equals_the_stored_bior22_table_and_remembered_pywavelets_values():
    """oracle/dwt_ref._spline_bior derives the spline members of the family from the CDF construction; the derivation
    reproduces the stored bior2.2 table exactly and the PyWavelets values of other members as far as they are
    remembered (4 digits) -- layout (zero padding, centring) included."""
    dl, rl = d._spline_bior(2, 2)
    assert dl == list(d._FILTERS["bior2.2"][0]) and rl == list(d._FILTERS["bior2.2"][1])
    known = {
        "bior1.3": ([-0.0884, 0.0884, 0.7071, 0.7071, 0.0884, -0.0884], [0, 0, 0.7071, 0.7071, 0, 0]),
        "bior3.1": ([-0.3536, 1.0607, 1.0607, -0.3536], [0.1768, 0.5303, 0.5303, 0.1768]),
        "bior2.4": ([0, 0.0331, -0.0663, -0.1768, 0.4198, 0.9944, 0.4198, -0.1768, -0.0663, 0.0331],
                    [0, 0, 0, 0.3536, 0.7071, 0.3536, 0, 0, 0, 0]),
        "bior3.3": ([0.0663, -0.1989, -0.1547, 0.9944, 0.9944, -0.1547, -0.1989, 0.0663],
                    [0, 0, 0.1768, 0.5303, 0.5303, 0.1768, 0, 0]),
    }
    for name, (kd, kr) in known.items():
        wv = d.Wavelet(name)
        assert np.abs(wv.dec_lo - np.array(kd)).max() < 6e-5 and np.abs(wv.rec_lo - np.array(kr)).max() < 6e-5, name
    assert {d.Wavelet(n).dec_len for n in ("bior1.1", "bior1.5", "bior2.6", "bior2.8", "bior3.5", "bior3.7", "bior3.9")} \
        == {2, 10, 14, 18, 12, 16, 20}


@pytest.mark.parametrize("name", d.WAVELETS)
def test_two_dimensional_round_trip_every_wavelet(name):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 45, 70))
    for mode in d.MODES:
        co = d.wavedec2(x, name, mode, 2)
        r = d.waverec2(co, name, mode)
        assert np.abs(r[:, :45, :70] - x).max() < 1e-10
