"""Device-resident batched entry points over libspiht_b200.so.

torch is used only for device memory and streams (plumbing); every stage of the
codec runs in the library's own CUDA kernels.  All functions take and return
CUDA tensors and do not synchronise, except where a host value is needed.
"""
import ctypes

import torch

from . import _lib
from .color_models import normalise as _norm_color


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _ctx_for(t):
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    ctx = _lib.get_context(dev)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    return ctx


def _settings_args(settings, C):
    color = _norm_color(settings.color_model)
    color_id = _lib.COLOR_IPT if color == "IPT" else _lib.COLOR_NONE
    scales = settings.per_channel_quant_scales
    if scales is not None:
        if len(scales) != C:
            raise ValueError(f"per_channel_quant_scales has {len(scales)} entries for {C} channels")
        sc = (ctypes.c_double * C)(*[float(s) for s in scales])
    else:
        sc = None
    return color_id, sc, float(settings.quantization_scale)


def _pixel_dtype(t, allow_u8=False):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    if allow_u8 and t.dtype == torch.uint8:
        return _lib.U8   # image bytes as stored on disk: scaled by 1 / 255 in float64 like utils.imload
    raise TypeError(f"pixels must be float32 or float64{' or uint8' if allow_u8 else ''}, got {t.dtype}")


def _check_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (libspiht_b200 has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def stream_stride(max_bits, C, geom):
    """row size in bytes (multiple of 8) that holds a stream of max_bits bits"""
    if max_bits is None or max_bits <= 0:
        return int(_lib.lib().spihtb_stream_bound(C, geom.enc_h, geom.enc_w, geom.ll_h, geom.ll_w))
    return ((int(max_bits) + 7) // 8 + 15) // 8 * 8


def forward(pixels, geom, settings):
    """spihtb_forward: pixels [B,C,H,W] -> int32 coefficient arrays [B,C,enc_h,enc_w]"""
    _check_cuda(pixels, "pixels")
    B, C, H, W = pixels.shape
    if (H, W) != (geom.h, geom.w):
        raise ValueError("pixel tensor does not match the planned geometry")
    ctx = _ctx_for(pixels)
    color_id, sc, q = _settings_args(settings, C)
    coeffs = torch.empty((B, C, geom.enc_h, geom.enc_w), dtype=torch.int32, device=pixels.device)
    _lib.check(_lib.lib().spihtb_forward(ctx.handle, _ptr(pixels), _pixel_dtype(pixels, True), B, C, ctypes.byref(geom),
                                         color_id, sc, q, _ptr(coeffs)))
    return coeffs


def inverse(coeffs, geom, settings, dtype=torch.float64):
    """spihtb_inverse: int32 coefficient arrays -> pixels [B,C,rec_h,rec_w]"""
    _check_cuda(coeffs, "coeffs")
    if coeffs.dtype != torch.int32:
        raise TypeError("coeffs must be int32")
    B, C = coeffs.shape[:2]
    ctx = _ctx_for(coeffs)
    color_id, sc, q = _settings_args(settings, C)
    out = torch.empty((B, C, geom.rec_h, geom.rec_w), dtype=dtype, device=coeffs.device)
    _lib.check(_lib.lib().spihtb_inverse(ctx.handle, _ptr(coeffs), B, C, ctypes.byref(geom), color_id, sc, q,
                                         _ptr(out), _pixel_dtype(out)))
    return out


def max_abs(coeffs):
    """spihtb_max_abs: per-image maximum magnitude of an int32 [B,...] coefficient batch -> int64 [B] (CUDA)"""
    _check_cuda(coeffs, "coeffs")
    if coeffs.dtype != torch.int32:
        raise TypeError("coeffs must be int32")
    B = coeffs.shape[0]
    ctx = _ctx_for(coeffs)
    out = torch.empty((B,), dtype=torch.int32, device=coeffs.device)   # uint32 bit patterns, all < 2^31
    _lib.check(_lib.lib().spihtb_max_abs(ctx.handle, _ptr(coeffs), B, coeffs[0].numel(), _ptr(out)))
    return out.to(torch.int64) & 0xffffffff


def _max_bits_args(max_bits, B, device):
    if isinstance(max_bits, torch.Tensor):
        mb = max_bits.to(device=device, dtype=torch.int64).contiguous()
        if mb.numel() != B:
            raise ValueError("per-image max_bits must have one entry per image")
        return 0, mb
    return (0 if max_bits is None else min(int(max_bits), 2 ** 64 - 1)), None


def encode_coeffs(coeffs, ll_h, ll_w, max_bits, out_stride=None, out=None):
    """spihtb_encode_coeffs -> (streams uint8 [B,stride], nbits int64 [B], max_n int32 [B], status int32 [B])"""
    _check_cuda(coeffs, "coeffs")
    if coeffs.dtype != torch.int32 or coeffs.dim() != 4:
        raise TypeError("coeffs must be an int32 [B,C,H,W] tensor")
    B, C, H, W = coeffs.shape
    ctx = _ctx_for(coeffs)
    scalar, per = _max_bits_args(max_bits, B, coeffs.device)
    if out_stride is None:
        if per is not None:
            raise ValueError("out_stride is required with per-image max_bits")
        if scalar == 0 or scalar > 8 * int(_lib.lib().spihtb_stream_bound(C, H, W, ll_h, ll_w)):
            out_stride = int(_lib.lib().spihtb_stream_bound(C, H, W, ll_h, ll_w))
        else:
            out_stride = ((scalar + 7) // 8 + 15) // 8 * 8
    if out is None:
        out = torch.empty((B, out_stride), dtype=torch.uint8, device=coeffs.device)
    nbits = torch.empty((B,), dtype=torch.int64, device=coeffs.device)
    max_n = torch.empty((B,), dtype=torch.int32, device=coeffs.device)
    status = torch.empty((B,), dtype=torch.int32, device=coeffs.device)
    _lib.check(_lib.lib().spihtb_encode_coeffs(ctx.handle, _ptr(coeffs), B, C, H, W, int(ll_h), int(ll_w), scalar,
                                               _ptr(per), _ptr(out), out_stride, _ptr(nbits), _ptr(max_n),
                                               _ptr(status)))
    return out, nbits, max_n, status


def decode_coeffs(streams, nbytes, max_n, C, H, W, ll_h, ll_w, out=None):
    """spihtb_decode_coeffs: streams uint8 [B,stride] (zero padded), nbytes int64 [B], max_n int32 [B]"""
    _check_cuda(streams, "streams")
    B, stride = streams.shape
    if stride % 8:
        raise ValueError("stream row stride must be a multiple of 8 bytes")
    ctx = _ctx_for(streams)
    nbytes = nbytes.to(device=streams.device, dtype=torch.int64).contiguous()
    max_n = max_n.to(device=streams.device, dtype=torch.int32).contiguous()
    if out is None:
        out = torch.empty((B, C, H, W), dtype=torch.int32, device=streams.device)
    _lib.check(_lib.lib().spihtb_decode_coeffs(ctx.handle, _ptr(streams), stride, _ptr(nbytes), _ptr(max_n), B, C, H,
                                               W, int(ll_h), int(ll_w), _ptr(out)))
    return out


def encode_images(pixels, geom, settings, max_bits, out_stride=None, coeffs=None, out=None):
    """spihtb_encode_images -> (streams, nbits, max_n, status, coeffs)"""
    _check_cuda(pixels, "pixels")
    B, C, H, W = pixels.shape
    if (H, W) != (geom.h, geom.w):
        raise ValueError("pixel tensor does not match the planned geometry")
    ctx = _ctx_for(pixels)
    color_id, sc, q = _settings_args(settings, C)
    scalar, per = _max_bits_args(max_bits, B, pixels.device)
    if out_stride is None:
        if per is not None:
            raise ValueError("out_stride is required with per-image max_bits")
        out_stride = stream_stride(scalar, C, geom)
    if coeffs is None:
        coeffs = torch.empty((B, C, geom.enc_h, geom.enc_w), dtype=torch.int32, device=pixels.device)
    if out is None:
        out = torch.empty((B, out_stride), dtype=torch.uint8, device=pixels.device)
    nbits = torch.empty((B,), dtype=torch.int64, device=pixels.device)
    max_n = torch.empty((B,), dtype=torch.int32, device=pixels.device)
    status = torch.empty((B,), dtype=torch.int32, device=pixels.device)
    _lib.check(_lib.lib().spihtb_encode_images(ctx.handle, _ptr(pixels), _pixel_dtype(pixels, True), B, C,
                                               ctypes.byref(geom), color_id, sc, q, scalar, _ptr(per), _ptr(coeffs),
                                               _ptr(out), out_stride, _ptr(nbits), _ptr(max_n), _ptr(status)))
    return out, nbits, max_n, status, coeffs


def decode_images(streams, nbytes, max_n, C, geom, settings, dtype=torch.float64, coeffs=None, out=None,
                  scratch_coeffs=False):
    """spihtb_decode_images -> (pixels [B,C,rec_h,rec_w], coeffs).

    scratch_coeffs=True (SPIHTB_OPT_SCRATCH_COEFFS): the caller does not read `coeffs`; the finest detail bands of an
    image are then zeroed only if its stream reaches them (undefined otherwise) and None is returned in their place.
    The pixels are the same either way."""
    _check_cuda(streams, "streams")
    B, stride = streams.shape
    if stride % 8:
        raise ValueError("stream row stride must be a multiple of 8 bytes")
    ctx = _ctx_for(streams)
    color_id, sc, q = _settings_args(settings, C)
    nbytes = nbytes.to(device=streams.device, dtype=torch.int64).contiguous()
    max_n = max_n.to(device=streams.device, dtype=torch.int32).contiguous()
    if coeffs is None:
        coeffs = torch.empty((B, C, geom.enc_h, geom.enc_w), dtype=torch.int32, device=streams.device)
    if out is None:
        out = torch.empty((B, C, geom.rec_h, geom.rec_w), dtype=dtype, device=streams.device)
    ctx.set_option(_lib.OPT_SCRATCH_COEFFS, 1 if scratch_coeffs else 0)
    try:
        _lib.check(_lib.lib().spihtb_decode_images(ctx.handle, _ptr(streams), stride, _ptr(nbytes), _ptr(max_n), B, C,
                                                   ctypes.byref(geom), color_id, sc, q, _ptr(coeffs), _ptr(out),
                                                   _pixel_dtype(out)))
    finally:
        if scratch_coeffs:
            ctx.set_option(_lib.OPT_SCRATCH_COEFFS, 0)
    return out, (None if scratch_coeffs else coeffs)
