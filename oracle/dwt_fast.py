"""ctypes front end of oracle/dwt_fast.c (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

wavedec2 / waverec2 with the structure and values of oracle/dwt_ref.py (the checker; the two agree to 1e-12,
tests/test_oracle_dwt.py), but with the per-axis filtering in compiled C -- what PyWavelets itself does.  Used by
the CPU baseline legs of bench.py so that they time a compiled transform like the reference's.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import dwt_ref

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdwt_fast.so")
_MODE = {"reflect": 0, "symmetric": 1, "periodization": 2}
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "dwt_fast.c")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(_LIB_PATH)
        dp, i = ctypes.POINTER(ctypes.c_double), ctypes.c_int
        L.dwt_out_len.argtypes = [i, i, i]
        L.idwt_out_len.argtypes = [i, i, i]
        L.dwt_last.argtypes = [dp, i, i, i, dp, dp, i, dp, dp]
        L.dwt_first.argtypes = [dp, i, i, i, dp, dp, i, dp, dp]
        L.idwt_last.argtypes = [dp, dp, i, i, i, dp, dp, i, dp]
        L.idwt_first.argtypes = [dp, dp, i, i, i, dp, dp, i, dp]
        for f in (L.dwt_last, L.dwt_first, L.idwt_last, L.idwt_first):
            f.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _dwt2_plane(x, wav, mode):
    """one level over a 2-D plane: axis -2 first, then axis -1 (pywt.wavedec2's order) -> aa, ad, da, dd"""
    L, F, m = lib(), wav.dec_len, _MODE[mode]
    flo, fhi = np.ascontiguousarray(wav.dec_lo), np.ascontiguousarray(wav.dec_hi)
    h, w = x.shape
    mh, mw = L.dwt_out_len(h, F, m), L.dwt_out_len(w, F, m)
    lo, hi = np.empty((mh, w)), np.empty((mh, w))
    L.dwt_first(_p(x), h, w, m, _p(flo), _p(fhi), F, _p(lo), _p(hi))
    aa, ad, da, dd = (np.empty((mh, mw)) for _ in range(4))
    L.dwt_last(_p(lo), mh, w, m, _p(flo), _p(fhi), F, _p(aa), _p(ad))
    L.dwt_last(_p(hi), mh, w, m, _p(flo), _p(fhi), F, _p(da), _p(dd))
    return aa, ad, da, dd


def wavedec2(data, wavelet, mode="reflect", level=None):
    """same result structure as dwt_ref.wavedec2 for data [c, h, w]"""
    wav = wavelet if isinstance(wavelet, dwt_ref.Wavelet) else dwt_ref.Wavelet(wavelet)
    data = np.ascontiguousarray(data, np.float64)
    c, h, w = data.shape
    maxlev = min(dwt_ref.dwt_max_level(h, wav.dec_len), dwt_ref.dwt_max_level(w, wav.dec_len))
    level = maxlev if level is None else level
    planes = [data[k] for k in range(c)]
    out = []
    for _ in range(level):
        res = [_dwt2_plane(np.ascontiguousarray(p), wav, mode) for p in planes]
        out.append(tuple(np.stack([r[q] for r in res]) for q in (2, 1, 3)))   # (cH = da, cV = ad, cD = dd)
        planes = [r[0] for r in res]
    out.append(np.stack(planes))
    out.reverse()
    return out


def _idwt2_plane(aa, ad, da, dd, wav, mode):
    """one level of synthesis: axis -1 first, then axis -2 (pywt.waverec2)"""
    L, F, m = lib(), wav.dec_len, _MODE[mode]
    glo, ghi = np.ascontiguousarray(wav.rec_lo), np.ascontiguousarray(wav.rec_hi)
    mh, mw = aa.shape
    ow, oh = L.idwt_out_len(mw, F, m), L.idwt_out_len(mh, F, m)
    lo, hi = np.empty((mh, ow)), np.empty((mh, ow))
    L.idwt_last(_p(aa), _p(ad), mh, mw, m, _p(glo), _p(ghi), F, _p(lo))
    L.idwt_last(_p(da), _p(dd), mh, mw, m, _p(glo), _p(ghi), F, _p(hi))
    out = np.empty((oh, ow))
    L.idwt_first(_p(lo), _p(hi), mh, ow, m, _p(glo), _p(ghi), F, _p(out))
    return out


def waverec2(coeffs, wavelet, mode="reflect"):
    """same result as dwt_ref.waverec2 for coefficient lists over [c, h, w] arrays"""
    wav = wavelet if isinstance(wavelet, dwt_ref.Wavelet) else dwt_ref.Wavelet(wavelet)
    a = np.asarray(coeffs[0], np.float64)
    for (da, ad, dd) in coeffs[1:]:
        if a.shape[-2] == dd.shape[-2] + 1:   # approximation one sample longer than the details: drop it
            a = a[..., :-1, :]
        if a.shape[-1] == dd.shape[-1] + 1:
            a = a[..., :-1]
        a = np.stack([_idwt2_plane(*(np.ascontiguousarray(np.asarray(v[k], np.float64)) for v in (a, ad, da, dd)),
                                   wav, mode) for k in range(a.shape[0])])
    return a
