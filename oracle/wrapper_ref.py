"""The reference's encode_image / decode_image restated over the CPU oracle
(TEST INFRASTRUCTURE, NOT PRODUCT CODE).  Follows spiht/spiht_wrapper.py:142-189
(encode order of operations) and :218-281 (decode), with PyWavelets replaced by
oracle/dwt_ref.py, colour-science by oracle/ipt_ref.py and the Rust coder by
oracle/spiht_ref.c.
"""
import numpy as np

from . import dwt_ref, ipt_ref, spiht_oracle


def quantize(arr, q_scale=10.0):          # spiht_wrapper.py:9-11
    arr = arr * q_scale
    return arr.astype(np.int32)


def dequantize(arr, q_scale=10.0):        # spiht_wrapper.py:13-14
    return arr / q_scale


def forward_coeffs(image, wavelet="bior2.2", mode="reflect", level=None, quantization_scale=50.0,
                   color_model=None, per_channel_quant_scales=None, return_float=False, fast=False):
    """spiht_wrapper.py:158-172 -> (int32 coeffs [c,Hc,Wc], ll_h, ll_w).  fast: the transform in compiled C
    (oracle/dwt_fast.c, for the timed CPU baseline) instead of the numpy checker."""
    image = np.asarray(image, np.float64)
    if color_model is not None:
        image = ipt_ref.convert(image, "RGB", color_model)
    if fast:
        from . import dwt_fast
        coeffs = dwt_fast.wavedec2(image, wavelet, mode, level)
    else:
        coeffs = dwt_ref.wavedec2(image, wavelet, mode, level)
    ll_h, ll_w = coeffs[0].shape[1], coeffs[0].shape[2]
    arr = dwt_ref.coeffs_to_array(coeffs)
    if per_channel_quant_scales is not None:
        arr = np.array(per_channel_quant_scales)[:, None, None] * arr
    if return_float:
        return arr * quantization_scale, ll_h, ll_w
    return quantize(arr, quantization_scale), ll_h, ll_w


def encode_image(image, wavelet="bior2.2", mode="reflect", level=None, quantization_scale=50.0,
                 color_model=None, per_channel_quant_scales=None, max_bits=None):
    """spiht_wrapper.py:142-189 -> dict(encoded_bytes, h, w, c, max_n, level)"""
    image = np.asarray(image)
    if image.ndim != 3:
        raise ValueError("image ndim must be 3: c,h,w")
    c, h, w = image.shape
    arr, ll_h, ll_w = forward_coeffs(image, wavelet, mode, level, quantization_scale, color_model,
                                     per_channel_quant_scales)
    if max_bits is None:
        max_bits = 99999999999999999
    data, max_n = spiht_oracle.encode(arr, ll_h, ll_w, max_bits)
    return dict(encoded_bytes=data, h=h, w=w, c=c, max_n=max_n, level=level)


def inverse_coeffs(rec_arr, h, w, wavelet="bior2.2", mode="reflect", level=None, quantization_scale=50.0,
                   color_model=None, per_channel_quant_scales=None, fast=False):
    """spiht_wrapper.py:259-281.  fast: see forward_coeffs."""
    slices, _, _ = dwt_ref.get_slices_and_h_w(h, w, wavelet, mode, level)
    rec_arr = np.asarray(rec_arr, np.float64)
    if per_channel_quant_scales is not None:
        rec_arr = rec_arr / np.array(per_channel_quant_scales)[:, None, None]
    rec_arr = dequantize(rec_arr, quantization_scale)
    if fast:
        from . import dwt_fast
        img = dwt_fast.waverec2(dwt_ref.array_to_coeffs(rec_arr, slices), wavelet, mode)
    else:
        img = dwt_ref.waverec2(dwt_ref.array_to_coeffs(rec_arr, slices), wavelet, mode)
    if color_model is not None:
        img = ipt_ref.convert(img, color_model, "RGB")
    return img


def decode_image(enc, wavelet="bior2.2", mode="reflect", quantization_scale=50.0,
                 color_model=None, per_channel_quant_scales=None):
    """spiht_wrapper.py:192-257"""
    h, w, c, level = enc["h"], enc["w"], enc["c"], enc["level"]
    slices, enc_h, enc_w = dwt_ref.get_slices_and_h_w(h, w, wavelet, mode, level)
    ll_h, ll_w = slices[0][1].stop, slices[0][2].stop
    rec = spiht_oracle.decode(enc["encoded_bytes"], enc["max_n"], c, enc_h, enc_w, ll_h, ll_w)
    return inverse_coeffs(rec, h, w, wavelet, mode, level, quantization_scale, color_model,
                          per_channel_quant_scales)
