"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/spiht_b200.h declares, plans geometry like the oracle, and
fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from spiht_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "spiht_b200.h")).read()
    declared = set(re.findall(r"\b(spihtb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.spihtb_version() == 100


def test_geom_struct_matches_header():
    from spiht_b200 import _lib
    assert ctypes.sizeof(_lib.Geom) == 4 * (11 + 6 * _lib.MAX_LEVELS)


@pytest.mark.parametrize("wavelet", sorted(__import__("spiht_b200")._lib.WAVELET_IDS))
@pytest.mark.parametrize("mode", ["reflect", "symmetric", "periodization"])
def test_plan_matches_oracle_geometry(wavelet, mode):
    from oracle import dwt_ref
    from spiht_b200 import _lib
    for (h, w) in [(256, 384), (1024, 1024), (70, 70), (67, 131), (2048, 2048), (511, 300)]:
        for level in (None, 1, 2):
            try:
                ll_h, ll_w, det = dwt_ref.wavedecn_shapes_2d(h, w, wavelet, mode, level)
            except ValueError:
                continue
            if len(det) == 0:
                with pytest.raises(ValueError):
                    _lib.plan(h, w, wavelet, mode, level)
                continue
            g = _lib.plan(h, w, wavelet, mode, level)
            assert (g.ll_h, g.ll_w, g.levels) == (ll_h, ll_w, len(det))
            _, eh, ew = dwt_ref.get_slices_and_h_w(h, w, wavelet, mode, level)
            assert (g.enc_h, g.enc_w) == (eh, ew)
            assert [(g.band_h[l], g.band_w[l]) for l in range(g.levels - 1, -1, -1)] == det


def test_get_slices_and_h_w_matches_oracle():
    from oracle import dwt_ref
    from spiht_b200.spiht_wrapper import SpihtSettings, get_slices_and_h_w
    for (h, w) in [(256, 384), (70, 70), (1024, 1024)]:
        a = get_slices_and_h_w(h, w, SpihtSettings(), None)
        b = dwt_ref.get_slices_and_h_w(h, w, "bior2.2", "reflect", None)
        assert a == b


def test_plan_errors():
    from spiht_b200 import _lib
    with pytest.raises(ValueError):
        _lib.plan(64, 64, "db4")
    with pytest.raises(ValueError):
        _lib.plan(64, 64, "bior2.2", "zero")
    with pytest.raises(ValueError):
        _lib.plan(64, 64, level=-1)
    with pytest.raises(ValueError):
        _lib.plan(4, 4)          # max level 0: nothing to code


def test_api_types_mirror_reference():
    import spiht_b200 as spiht
    st = spiht.SpihtSettings()
    assert (st.wavelet, st.quantization_scale, st.mode, st.color_model, st.per_channel_quant_scales) == \
        ("bior2.2", 50.0, "reflect", None, None)
    # demonstrate.py:23-29 passes the settings positionally
    st2 = spiht.SpihtSettings("bior4.4", 1.0, "symmetric", "IPT", [100, 20, 20])
    assert st2.color_model == "IPT"
    er = spiht.EncodingResult(b"\x01\x02", 4, 5, 3, 7, None)
    assert er._encoding_version == spiht.ENCODER_DECODER_VERSION == "0.0.2"
    assert spiht.EncodingResult.from_dict(er.to_dict()) == er
    from spiht_b200.spiht_wrapper import quantize, dequantize
    assert quantize(np.array([-1.99, 1.99, 0.5]), 1.0).tolist() == [-1, 1, 0]   # truncation toward zero
    assert dequantize(np.array([5.0]), 10.0)[0] == 0.5


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import spiht_b200 as spiht
    from spiht_b200 import _lib
    with pytest.raises((RuntimeError, _lib.SpihtB200Error)):
        spiht.encode(np.zeros((1, 8, 8), np.int32), 2, 2, 100)
    with pytest.raises((RuntimeError, _lib.SpihtB200Error)):
        spiht.encode_image(np.zeros((3, 64, 64)), spiht.SpihtSettings())


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spiht_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", ""), f


def test_release_library_exports_no_debug_symbols():
    """the phase counters (-DSPIHTB_PROF) are not part of the product library"""
    import subprocess
    from spiht_b200 import build
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], capture_output=True, text=True).stdout
    names = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("spihtb_")}
    from spiht_b200 import _lib
    assert names == set(_lib.EXPORTS), names ^ set(_lib.EXPORTS)


def test_spiht_alias_package_matches_reference_imports():
    """the import lines of the reference's scripts (encode_decode.py:10-14, make_gif.py, demonstrate.py,
    spiht/tests/*.py) resolve against the `spiht` alias package"""
    import spiht
    from spiht import encode_image, decode_image, EncodingResult, SpihtSettings, ENCODER_DECODER_VERSION  # noqa: F401
    from spiht import encode, decode  # noqa: F401
    from spiht.spiht_wrapper import SpihtSettings as S2, get_slices_and_h_w, decode_rec_array, decode_from_rec_arr  # noqa: F401
    from spiht.utils import imload, bytes_to_bits  # noqa: F401
    from spiht.color_models import convert  # noqa: F401
    import spiht.spiht as spiht_rs
    import spiht_b200
    assert S2 is spiht_b200.SpihtSettings and spiht.spiht_wrapper is spiht_b200.spiht_wrapper
    assert spiht_rs.encode is spiht_b200.encode and hasattr(spiht_rs, "decode_with_metadata")
    with pytest.raises(ValueError):
        convert(np.zeros((3, 4, 4)), "RGB", "CIE Lab")     # color_models.py:7-10


def test_library_filter_banks_equal_the_oracle_tables():
    """spihtb_wavelet_filters: the filter bank the kernels use, per wavelet id, against oracle/dwt_ref.py -- the three
    stored tables digit for digit, the derived spline pairs to rounding (two independent derivations: exact integer
    polynomials in C++, Fractions in Python)."""
    import ctypes
    import numpy as np
    from oracle import dwt_ref
    from spiht_b200 import _lib
    assert set(_lib.WAVELET_IDS) == set(dwt_ref.WAVELETS)
    L = _lib.lib()
    for name, wid in _lib.WAVELET_IDS.items():
        F = ctypes.c_int32()
        dl = (ctypes.c_double * 20)()
        rl = (ctypes.c_double * 20)()
        assert L.spihtb_wavelet_filters(wid, ctypes.byref(F), dl, rl) == 0, name
        wv = dwt_ref.Wavelet(name)
        assert F.value == wv.dec_len
        got_d, got_r = np.array(dl[:F.value]), np.array(rl[:F.value])
        tol = 0.0 if name in ("bior2.2", "bior4.4", "bior6.8", "bior5.5") else 4e-16
        assert np.abs(got_d - wv.dec_lo).max() <= tol and np.abs(got_r - wv.rec_lo).max() <= tol, name
        assert np.array_equal(got_d == 0.0, wv.dec_lo == 0.0) and np.array_equal(got_r == 0.0, wv.rec_lo == 0.0)
    F = ctypes.c_int32()
    assert L.spihtb_wavelet_filters(99, ctypes.byref(F), dl, rl) != 0
