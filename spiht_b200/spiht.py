"""Drop-in for the reference's native module `spiht.spiht` (pyo3, src/lib.rs:58-65).

Same names and positional signatures; the work runs on the GPU through
libspiht_b200.so (spihtb_encode / spihtb_decode).  Host buffers in, host
objects out, exactly like the pyo3 functions.
"""
import ctypes

import numpy as np

from . import _lib


def encode(x, ll_h, ll_w, max_bits):
    """src/lib.rs:24-32: encode(x: int32[c,h,w], ll_h, ll_w, max_bits) -> (bytes, max_n)"""
    if not isinstance(x, np.ndarray) or x.dtype != np.int32 or x.ndim != 3:
        # pyo3: PyReadonlyArray3<i32> extraction failure
        raise TypeError("argument 'x': expected a 3-D numpy int32 array")
    if ll_h < 0 or ll_w < 0 or max_bits < 0:
        raise OverflowError("can't convert negative int to unsigned")
    x = np.ascontiguousarray(x)
    c, h, w = x.shape
    if c == 0 or h == 0 or w == 0:
        raise _lib.SpihtB200Error(_lib.EINVAL, "empty coefficient array (the reference panics on .max().unwrap())")
    ctx = _lib.get_context(_current_device())
    _bind_stream(ctx)
    out = ctypes.c_void_p()
    nbits = ctypes.c_uint64()
    max_n = ctypes.c_int32()
    _lib.check(_lib.lib().spihtb_encode(ctx.handle, x.ctypes.data, c, h, w, int(ll_h), int(ll_w),
                                        min(int(max_bits), 2 ** 64 - 1), ctypes.byref(out),
                                        ctypes.byref(nbits), ctypes.byref(max_n)))
    nbytes = (nbits.value + 7) // 8
    return ctypes.string_at(out, nbytes), int(max_n.value)


def decode(data_u8, n, c, h, w, ll_h, ll_w):
    """src/lib.rs:35-42: decode(data: bytes, n, c, h, w, ll_h, ll_w) -> int32[c,h,w]"""
    data = bytes(data_u8)
    if not 0 <= int(n) <= 255:
        raise OverflowError("out of range integral type conversion attempted")  # n: u8
    out = np.empty((int(c), int(h), int(w)), dtype=np.int32)
    ctx = _lib.get_context(_current_device())
    _bind_stream(ctx)
    _lib.check(_lib.lib().spihtb_decode(ctx.handle, data, len(data), int(n), int(c), int(h), int(w),
                                        int(ll_h), int(ll_w), out.ctypes.data))
    return out


def decode_with_metadata(data_u8, n, c, h, w, ll_h, ll_w, top_slice, other_slices):
    """src/lib.rs:47-56: decode_with_metadata(data, n, c, h, w, ll_h, ll_w, top_slice, other_slices)
    -> (int32[c,h,w], int32[8 * len(data) + 1, 8]).  top_slice = [(start_i, end_i), (start_j, end_j)];
    other_slices = per detail level, coarsest first, three [(start_i, end_i), (start_j, end_j)] in the order
    da, ad, dd (spiht_wrapper.py:232-250)."""
    data = bytes(data_u8)
    if not 0 <= int(n) <= 255:
        raise OverflowError("out of range integral type conversion attempted")  # n: u8
    out = np.empty((int(c), int(h), int(w)), dtype=np.int32)
    meta = np.empty((8 * len(data) + 1, 8), dtype=np.int32)
    top = np.array([top_slice[0][0], top_slice[0][1], top_slice[1][0], top_slice[1][1]], dtype=np.int32)
    other = np.array([[[f[0][0], f[0][1], f[1][0], f[1][1]] for f in lvl] for lvl in other_slices], dtype=np.int32)
    other = np.ascontiguousarray(other.reshape(len(other_slices), 3, 4))
    ctx = _lib.get_context(_current_device())
    _bind_stream(ctx)
    _lib.check(_lib.lib().spihtb_decode_with_metadata(ctx.handle, data, len(data), int(n), int(c), int(h), int(w),
                                                      int(ll_h), int(ll_w), top.ctypes.data,
                                                      other.ctypes.data if len(other_slices) else None,
                                                      len(other_slices), out.ctypes.data, meta.ctypes.data))
    return out, meta


def _current_device():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def _bind_stream(ctx):
    try:
        import torch
        if torch.cuda.is_available():
            ctx.set_stream(torch.cuda.current_stream(ctx.device).cuda_stream)
    except ImportError:
        pass
