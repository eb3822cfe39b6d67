"""float64 numpy restatement of colour.convert(x, 'RGB', 'IPT') and back
(TEST INFRASTRUCTURE, NOT PRODUCT CODE).

The reference calls an un-vendored third party: colour-science==0.4.4
(pyproject.toml:10, requirements.txt:1) from spiht/color_models.py:12.  It is
absent from this image, so the published transform is restated: linear sRGB
-> CIE XYZ with the 4-digit IEC 61966-2-1 matrix (no CCTF decoding, no
chromatic adaptation), then Ebner & Fairchild (1998) IPT:
    LMS = M_xyz2lms . XYZ ; LMS' = sign(LMS) |LMS|^0.43 ; IPT = M_lms2ipt . LMS'
The way back uses numpy-inverted matrices throughout: colour-science 0.4.4
defines MATRIX_XYZ_TO_sRGB = np.linalg.inv(MATRIX_sRGB_TO_XYZ)
(colour/models/rgb/datasets/srgb.py; releases before 0.4 hard-coded the rounded
4-digit inverse, which is not an exact inverse), so RGB -> IPT -> RGB is the
identity to rounding.

PARITY UNPINNED: the reference has no test on colour values.
"""
import numpy as np

M_RGB_TO_XYZ = np.array([[0.4124, 0.3576, 0.1805],
                         [0.2126, 0.7152, 0.0722],
                         [0.0193, 0.1192, 0.9505]])
M_XYZ_TO_RGB = np.linalg.inv(M_RGB_TO_XYZ)
M_XYZ_TO_LMS = np.array([[0.4002, 0.7075, -0.0807],
                         [-0.2280, 1.1500, 0.0612],
                         [0.0000, 0.0000, 0.9184]])
M_LMSP_TO_IPT = np.array([[0.4000, 0.4000, 0.2000],
                          [4.4550, -4.8510, 0.3960],
                          [0.8056, 0.3572, -1.1628]])
M_LMS_TO_XYZ = np.linalg.inv(M_XYZ_TO_LMS)
M_IPT_TO_LMSP = np.linalg.inv(M_LMSP_TO_IPT)

SUPPORTED_MODELS = ("RGB", "IPT")


def _spow(a, p):
    return np.sign(a) * np.abs(a) ** p


def _mat(m, x):  # x: (..., 3)
    return np.einsum("ij,...j->...i", m, x)


def rgb_to_ipt_hwc(rgb):
    xyz = _mat(M_RGB_TO_XYZ, np.asarray(rgb, np.float64))
    lms = _mat(M_XYZ_TO_LMS, xyz)
    return _mat(M_LMSP_TO_IPT, _spow(lms, 0.43))


def ipt_to_rgb_hwc(ipt):
    lmsp = _mat(M_IPT_TO_LMSP, np.asarray(ipt, np.float64))
    lms = _spow(lmsp, 1.0 / 0.43)
    return _mat(M_XYZ_TO_RGB, _mat(M_LMS_TO_XYZ, lms))


def convert(im, src, dest):
    """spiht/color_models.py:6-13 (CHW in, CHW out), restricted to RGB <-> IPT"""
    s, d = str(src).upper(), str(dest).upper()
    for name in (s, d):
        if name not in SUPPORTED_MODELS:
            raise ValueError(f"{name} is not a supported color model. Supported models are {SUPPORTED_MODELS}")
    im = np.moveaxis(np.asarray(im, np.float64), 0, -1)
    if s == "RGB" and d == "IPT":
        im = rgb_to_ipt_hwc(im)
    elif s == "IPT" and d == "RGB":
        im = ipt_to_rgb_hwc(im)
    return np.moveaxis(im, -1, 0)
