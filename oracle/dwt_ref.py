"""float64 numpy restatement of the PyWavelets calls on the reference's hot path
(TEST INFRASTRUCTURE, NOT PRODUCT CODE).

The reference calls an un-vendored third party: PyWavelets==1.5.0
(pyproject.toml:17, requirements.txt:8) at spiht/spiht_wrapper.py:102
(wavedecn_shapes), :163 (wavedec2), :165 (coeffs_to_array), :275
(array_to_coeffs), :276 (waverec2).  PyWavelets is absent from this image and
from the wheelhouse, so its published algorithm is restated here:

  * non-periodization modes: len_out = (N + F - 1) // 2,
        out[k] = sum_j filt[j] * x_ext[2k + 1 - j]               (analysis)
        rec[n] = sum_k cA[k] rec_lo[n + F - 2 - 2k] + cD[k] rec_hi[...]  (synthesis)
  * periodization: len_out = ceil(N / 2), odd N padded with its last sample,
        out[k] = sum_j filt[j] * x_per[(2k + F/2 - j) mod Np]
        rec[n] = sum_k c[k] * g[n - 2k + F/2 - 1] (k periodic)
  * wavedec2: per level axis -2 first, then axis -1; cH = 'da', cV = 'ad', cD = 'dd'
  * coeffs_to_array: 'ad' top-right, 'da' bottom-left, 'dd' diagonal, zero gaps.

PARITY UNPINNED: the reference's tests assert no transform value
(spiht/tests/test_spiht.py has no numeric assertion; test_rust.py:56 only
asserts SPIHT losslessness).  What is pinned instead: filter identities,
perfect reconstruction, and the band geometry used by spiht_wrapper.py:92-139.
"""
import math

import numpy as np

_S2 = math.sqrt(2.0)

# dec_lo / rec_lo exactly as PyWavelets stores them (zero padding included)
_FILTERS = {
    "bior2.2": (
        [0.0, -0.1767766952966369, 0.3535533905932738, 1.0606601717798214,
         0.3535533905932738, -0.1767766952966369],
        [0.0, 0.3535533905932738, 0.7071067811865476, 0.3535533905932738, 0.0, 0.0],
    ),
    "bior4.4": (
        [0.0, 0.03782845550726404, -0.023849465019556843, -0.11062440441843718,
         0.37740285561283066, 0.8526986790088938, 0.37740285561283066,
         -0.11062440441843718, -0.023849465019556843, 0.03782845550726404],
        [0.0, -0.06453888262869706, -0.04068941760916406, 0.41809227322161724,
         0.7884856164055829, 0.41809227322161724, -0.04068941760916406,
         -0.06453888262869706, 0.0, 0.0],
    ),
    "bior6.8": (
        [0.0, 0.0019088317364812906, -0.0019142861290887667, -0.016990639867602342,
         0.01193456527972926, 0.04973290349094079, -0.07726317316720414,
         -0.09405920349573646, 0.4207962846098268, 0.8259229974584023,
         0.4207962846098268, -0.09405920349573646, -0.07726317316720414,
         0.04973290349094079, 0.01193456527972926, -0.016990639867602342,
         -0.0019142861290887667, 0.0019088317364812906],
        [0.0, 0.0, 0.0, 0.014426282505624435, 0.014467504896790148,
         -0.07872200106262882, -0.04036797903033992, 0.41784910915027457,
         0.7589077294536541, 0.41784910915027457, -0.04036797903033992,
         -0.07872200106262882, 0.014467504896790148, 0.014426282505624435,
         0.0, 0.0, 0.0, 0.0],
    ),
}


def _spline_bior(nr, nd):
    """Cohen-Daubechies-Feauveau biorthogonal spline pair biorNr.Nd, derived (not recalled), laid out as PyWavelets
    stores it.  rec_lo is the B-spline of order nr, sqrt2 ((1 + z) / 2)^nr; dec_lo is
        sqrt2 ((1 + z) / 2)^nd  sum_{m < K} C(K-1+m, m) ((2 - z - 1/z) / 4)^m,   K = (nr + nd) / 2
    (Daubechies, Ten Lectures, 8.3.4), nr + 2 nd - 1 taps.  Layout: even common length F; odd-length filters (even nr)
    get a leading zero, dec_lo centred on F/2 and rec_lo on F/2 - 1; even-length ones (odd nr) are both centred on
    (F-1)/2.  The derived bior2.2 equals the stored table; bior1.3, 2.4, 3.1, 3.3 equal the PyWavelets tables as far
    as they are remembered; tests/test_oracle_dwt.py checks perfect reconstruction and the vanishing moments."""
    from fractions import Fraction
    from math import comb
    assert (nr + nd) % 2 == 0
    K = (nr + nd) // 2

    def mul(a, b):
        out = [Fraction(0)] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                out[i + j] += x * y
        return out
    half = [Fraction(1, 2), Fraction(1, 2)]
    s2 = [Fraction(-1, 4), Fraction(1, 2), Fraction(-1, 4)]          # sin^2(w/2) as a polynomial in z, centred
    rec = [Fraction(1)]
    for _ in range(nr):
        rec = mul(rec, half)
    q = [Fraction(0)] * (2 * K - 1)                                  # centred at index K - 1
    pw = [Fraction(1)]
    for m in range(K):
        off = K - 1 - m
        for i, v in enumerate(pw):
            q[off + i] += comb(K - 1 + m, m) * v
        pw = mul(pw, s2)
    dec = q
    for _ in range(nd):
        dec = mul(dec, half)
    assert len(dec) == nr + 2 * nd - 1 and len(rec) == nr + 1
    if nr % 2 == 0:
        F = len(dec) + 1
        d0, r0 = 1, F // 2 - 1 - nr // 2
    else:
        F = len(dec)
        d0, r0 = 0, (F - len(rec)) // 2
    dec_lo, rec_lo = [0.0] * F, [0.0] * F
    for i, v in enumerate(dec):
        dec_lo[d0 + i] = _S2 * v.numerator / v.denominator          # denominators are powers of two: one rounding
    for i, v in enumerate(rec):
        rec_lo[r0 + i] = _S2 * v.numerator / v.denominator
    return dec_lo, rec_lo


# bior5.5 is not a spline pair: PyWavelets' table as remembered (sums sqrt2 to the last bit, perfect reconstruction
# to 4e-12 -- the accuracy of the published table itself; 4 and 6 zeros at z = -1, as odd-length symmetric filters must
# have an even number)
_FILTERS["bior5.5"] = (
    [0.0, 0.0, 0.03968708834740544, 0.007948108637240322, -0.05446378846823691, 0.34560528195603346,
     0.7366601814282105, 0.34560528195603346, -0.05446378846823691, 0.007948108637240322, 0.03968708834740544, 0.0],
    [0.013456709459118716, -0.002694966880111507, -0.13670658466432914, -0.09350469740093886, 0.47680326579848425,
     0.8995061097486484, 0.47680326579848425, -0.09350469740093886, -0.13670658466432914, -0.002694966880111507,
     0.013456709459118716, 0.0],
)
# the rest of PyWavelets' bior family: the spline pairs, derived
for _nr, _nd in ((1, 1), (1, 3), (1, 5), (2, 4), (2, 6), (2, 8), (3, 1), (3, 3), (3, 5), (3, 7), (3, 9)):
    _FILTERS["bior%d.%d" % (_nr, _nd)] = _spline_bior(_nr, _nd)
WAVELETS = tuple(sorted(_FILTERS))

MODES = ("reflect", "symmetric", "periodization")


class Wavelet:
    def __init__(self, name):
        if name not in _FILTERS:
            raise ValueError(f"Unknown wavelet name '{name}'")
        dec_lo, rec_lo = _FILTERS[name]
        F = len(dec_lo)
        self.name = name
        self.dec_len = F
        self.dec_lo = np.array(dec_lo, np.float64)
        self.rec_lo = np.array(rec_lo, np.float64)
        self.dec_hi = np.array([(-1.0) ** (F - 1 - i) * rec_lo[i] for i in range(F)])
        self.rec_hi = np.array([(-1.0) ** i * dec_lo[i] for i in range(F)])


def dwt_coeff_len(n, f, mode):
    if mode == "periodization":
        return (n + 1) // 2
    return (n + f - 1) // 2


def dwt_max_level(n, f):
    if f < 2 or n < f - 1:
        return 0
    return int(math.floor(math.log2(n // (f - 1))))


def _ext_index(idx, n, mode):
    """map arbitrary integer sample indices to [0, n) for the boundary mode"""
    idx = np.asarray(idx)
    if mode == "reflect":       # whole-sample: ... x2 x1 | x0 x1 ... xN-1 | xN-2 ...
        if n == 1:
            return np.zeros_like(idx)
        p = 2 * n - 2
        m = np.mod(idx, p)
        return np.where(m >= n, p - m, m)
    if mode == "symmetric":     # half-sample: ... x1 x0 | x0 x1 ... xN-1 | xN-1 ...
        p = 2 * n
        m = np.mod(idx, p)
        return np.where(m >= n, p - 1 - m, m)
    raise ValueError(f"unsupported mode {mode}")


def _check_mode(mode):
    if mode not in MODES:
        raise ValueError(f"Unknown mode name '{mode}' (supported: {MODES})")


def dwt_axis(x, wav, mode, axis):
    """single-level analysis along `axis` -> (cA, cD)"""
    _check_mode(mode)
    x = np.moveaxis(np.asarray(x, np.float64), axis, -1)
    n = x.shape[-1]
    f = wav.dec_len
    lo = np.zeros(x.shape[:-1] + (dwt_coeff_len(n, f, mode),))
    hi = np.zeros_like(lo)
    k = np.arange(lo.shape[-1])
    if mode == "periodization":
        if n % 2:
            x = np.concatenate([x, x[..., -1:]], axis=-1)
        npad = x.shape[-1]
        for j in range(f):
            src = x[..., np.mod(2 * k + f // 2 - j, npad)]
            lo += wav.dec_lo[j] * src
            hi += wav.dec_hi[j] * src
    else:
        for j in range(f):
            src = x[..., _ext_index(2 * k + 1 - j, n, mode)]
            lo += wav.dec_lo[j] * src
            hi += wav.dec_hi[j] * src
    return np.moveaxis(lo, -1, axis), np.moveaxis(hi, -1, axis)


def idwt_axis(ca, cd, wav, mode, axis):
    """single-level synthesis along `axis`"""
    _check_mode(mode)
    ca = np.moveaxis(np.asarray(ca, np.float64), axis, -1)
    cd = np.moveaxis(np.asarray(cd, np.float64), axis, -1)
    if ca.shape != cd.shape:
        raise ValueError("coefficient shape mismatch")
    m = ca.shape[-1]
    f = wav.dec_len
    if mode == "periodization":
        nout = 2 * m
        out = np.zeros(ca.shape[:-1] + (nout,))
        n = np.arange(nout)
        # rec[n] = sum_t g[t] c[k], t = n - 2k + f/2 - 1  =>  k = (n + f/2 - 1 - t)/2 (periodic)
        for t in range(f):
            num = n + f // 2 - 1 - t
            ok = (num % 2) == 0
            kk = np.mod(num // 2, m)
            out += np.where(ok, wav.rec_lo[t] * ca[..., kk] + wav.rec_hi[t] * cd[..., kk], 0.0)
    else:
        nout = 2 * m - f + 2
        if nout <= 0:
            raise ValueError("coefficient arrays too short for this wavelet")
        out = np.zeros(ca.shape[:-1] + (nout,))
        n = np.arange(nout)
        for t in range(f):
            num = n + f - 2 - t          # = 2k
            kk = num // 2
            ok = ((num % 2) == 0) & (kk >= 0) & (kk < m)
            kk = np.clip(kk, 0, m - 1)
            out += np.where(ok, wav.rec_lo[t] * ca[..., kk] + wav.rec_hi[t] * cd[..., kk], 0.0)
    return np.moveaxis(out, -1, axis)


def wavedec2(data, wavelet, mode="reflect", level=None):
    """pywt.wavedec2 over the last two axes -> [cA_L, (cH_L, cV_L, cD_L), ..., (cH_1, cV_1, cD_1)]"""
    wav = wavelet if isinstance(wavelet, Wavelet) else Wavelet(wavelet)
    data = np.asarray(data, np.float64)
    h, w = data.shape[-2], data.shape[-1]
    maxlev = min(dwt_max_level(h, wav.dec_len), dwt_max_level(w, wav.dec_len))
    if level is None:
        level = maxlev
    elif level < 0:
        raise ValueError(f"Level value of {level} is too low . Minimum level is 0.")
    out = []
    a = data
    for _ in range(level):
        lo, hi = dwt_axis(a, wav, mode, -2)
        aa, ad = dwt_axis(lo, wav, mode, -1)
        da, dd = dwt_axis(hi, wav, mode, -1)
        out.append((da, ad, dd))     # (cH, cV, cD)
        a = aa
    out.append(a)
    out.reverse()
    return out


def wavedecn_shapes_2d(h, w, wavelet, mode="reflect", level=None):
    """band geometry: returns (ll_h, ll_w, [(dh, dw) coarse -> fine])"""
    _check_mode(mode)
    wav = wavelet if isinstance(wavelet, Wavelet) else Wavelet(wavelet)
    f = wav.dec_len
    maxlev = min(dwt_max_level(h, f), dwt_max_level(w, f))
    if level is None:
        level = maxlev
    elif level < 0:
        raise ValueError(f"Level value of {level} is too low . Minimum level is 0.")
    det = []
    for _ in range(level):
        h, w = dwt_coeff_len(h, f, mode), dwt_coeff_len(w, f, mode)
        det.append((h, w))
    det.reverse()
    return h, w, det


def get_slices_and_h_w(h, w, wavelet, mode, level):
    """spiht_wrapper.py:92-139: (slices, enc_h, enc_w); slices[0] is the LL block"""
    ll_h, ll_w, det = wavedecn_shapes_2d(h, w, wavelet, mode, level)
    sh, sw = ll_h, ll_w
    slices = [(slice(None), slice(sh), slice(sw))]
    for dh, dw in det:
        slices.append({
            "ad": (slice(None), slice(0, dh), slice(sw, sw + dw)),
            "da": (slice(None), slice(sh, sh + dh), slice(0, dw)),
            "dd": (slice(None), slice(sh, sh + dh), slice(sw, sw + dw)),
        })
        sh += dh
        sw += dw
    return slices, sh, sw


def coeffs_to_array(coeffs):
    """pywt.coeffs_to_array(coeffs, axes=(-2,-1)) for wavedec2-format input"""
    a = coeffs[0]
    lead = a.shape[:-2]
    sh, sw = a.shape[-2:]
    hh = sh + sum(c[2].shape[-2] for c in coeffs[1:])
    ww = sw + sum(c[2].shape[-1] for c in coeffs[1:])
    arr = np.zeros(lead + (hh, ww), a.dtype)
    arr[..., :sh, :sw] = a
    for (da, ad, dd) in coeffs[1:]:
        arr[..., :ad.shape[-2], sw:sw + ad.shape[-1]] = ad
        arr[..., sh:sh + da.shape[-2], :da.shape[-1]] = da
        arr[..., sh:sh + dd.shape[-2], sw:sw + dd.shape[-1]] = dd
        sh += dd.shape[-2]
        sw += dd.shape[-1]
    return arr


def array_to_coeffs(arr, slices):
    """pywt.array_to_coeffs(arr, slices, output_format='wavedec2')"""
    out = [arr[slices[0]]]
    for s in slices[1:]:
        out.append((arr[s["da"]], arr[s["ad"]], arr[s["dd"]]))
    return out


def waverec2(coeffs, wavelet, mode="reflect"):
    """pywt.waverec2 over the last two axes (synthesis: axis -1 first, then -2)"""
    wav = wavelet if isinstance(wavelet, Wavelet) else Wavelet(wavelet)
    a = np.asarray(coeffs[0], np.float64)
    for (da, ad, dd) in coeffs[1:]:
        # approximation one sample longer than the details (odd sizes): drop it
        if a.shape[-2] == dd.shape[-2] + 1:
            a = a[..., :-1, :]
        if a.shape[-1] == dd.shape[-1] + 1:
            a = a[..., :-1]
        lo = idwt_axis(a, ad, wav, mode, -1)
        hi = idwt_axis(da, dd, wav, mode, -1)
        a = idwt_axis(lo, hi, wav, mode, -2)
    return a
