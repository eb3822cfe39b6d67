"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Same call shapes as the reference's pyo3 module `spiht.spiht` (src/lib.rs:24-42):
    encode(x: int32[c,h,w], ll_h, ll_w, max_bits) -> (bytes, max_n)
    decode(data: bytes, n, c, h, w, ll_h, ll_w)   -> int32[c,h,w]
backed by oracle/spiht_ref.c (faithful restatement of src/encoder_decoder.rs).
`model_encode` is the pyramid/generation formulation (oracle/spiht_model.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libspiht_oracle.so")
_lib = None


class OraclePanic(RuntimeError):
    """The reference would have panicked (assert! / out-of-bounds index)."""


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("spiht_ref.c", "spiht_model.c", "spiht_meta.c", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_LIB_PATH):
                raise
        L = ctypes.CDLL(_LIB_PATH)
        u64, i32p = ctypes.c_uint64, ctypes.POINTER(ctypes.c_int32)
        u8pp = ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8))
        for name in ("spiht_ref_encode", "spiht_model_encode"):
            f = getattr(L, name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p, u64, u64, u64, u64, u64, u64, u8pp,
                          ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_int)]
        L.spiht_ref_decode.restype = ctypes.c_int
        L.spiht_ref_decode.argtypes = [ctypes.c_char_p, u64, ctypes.c_uint, u64, u64, u64, u64, u64,
                                       ctypes.c_void_p]
        L.spiht_ref_decode_with_metadata.restype = ctypes.c_int
        L.spiht_ref_decode_with_metadata.argtypes = [ctypes.c_char_p, u64, ctypes.c_uint, u64, u64, u64, u64, u64,
                                                     ctypes.c_void_p, ctypes.c_void_p, u64, ctypes.c_void_p,
                                                     ctypes.c_void_p]
        L.spiht_ref_free.argtypes = [ctypes.c_void_p]
        L.spiht_ref_set_bit.restype = ctypes.c_int32
        L.spiht_ref_set_bit.argtypes = [ctypes.c_int32, ctypes.c_uint, ctypes.c_int]
        L.spiht_ref_is_bit_set.argtypes = [ctypes.c_int32, ctypes.c_uint]
        L.spiht_ref_is_element_sig.argtypes = [ctypes.c_int32, ctypes.c_uint]
        L.spiht_ref_has_descendents_past_offspring.argtypes = [u64, u64, u64, u64]
        L.spiht_ref_get_offspring.argtypes = [u64, u64, u64, u64, u64, u64, ctypes.c_void_p]
        L.spiht_ref_max_n.argtypes = [ctypes.c_int32]
        L.spiht_model_geom_ok.argtypes = [u64, u64, u64, u64]
        L.spiht_model_pyramid.argtypes = [ctypes.c_void_p, u64, u64, u64, u64, u64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc == 1:
        raise OraclePanic("assertion failed: ll_h > 1 && ll_w > 1")
    if rc == 2:
        raise OraclePanic("index out of bounds (the reference would panic)")
    if rc:
        raise MemoryError(f"oracle error {rc}")


def _encode(fn, x, ll_h, ll_w, max_bits):
    x = np.asarray(x)
    if x.dtype != np.int32:
        raise TypeError("x must be int32")  # pyo3: PyReadonlyArray3<i32>
    if x.ndim != 3:
        raise TypeError("x must be 3-D (c,h,w)")
    x = np.ascontiguousarray(x)
    c, h, w = x.shape
    out = ctypes.POINTER(ctypes.c_uint8)()
    nbits = ctypes.c_uint64()
    max_n = ctypes.c_int()
    _check(fn(x.ctypes.data, c, h, w, ll_h, ll_w, int(max_bits), ctypes.byref(out),
              ctypes.byref(nbits), ctypes.byref(max_n)))
    nbytes = (nbits.value + 7) // 8
    data = ctypes.string_at(out, nbytes)
    lib().spiht_ref_free(out)
    return data, max_n.value, nbits.value


def encode(x, ll_h, ll_w, max_bits):
    """lib.rs:24-32 -> (bytes, max_n)"""
    data, max_n, _ = _encode(lib().spiht_ref_encode, x, ll_h, ll_w, max_bits)
    return data, max_n


def encode_nbits(x, ll_h, ll_w, max_bits):
    return _encode(lib().spiht_ref_encode, x, ll_h, ll_w, max_bits)


def model_encode(x, ll_h, ll_w, max_bits):
    data, max_n, _ = _encode(lib().spiht_model_encode, x, ll_h, ll_w, max_bits)
    return data, max_n


def decode(data, n, c, h, w, ll_h, ll_w):
    """lib.rs:35-42 -> int32[c,h,w]"""
    data = bytes(data)
    out = np.empty((c, h, w), dtype=np.int32)
    _check(lib().spiht_ref_decode(data, len(data), int(n), c, h, w, ll_h, ll_w, out.ctypes.data))
    return out


def decode_with_metadata(data, n, c, h, w, ll_h, ll_w, top_slice, other_slices):
    """lib.rs:47-56 -> (int32[c,h,w], int32[8 * len(data) + 1, 8]); top_slice = [(si, ei), (sj, ej)],
    other_slices = per level (coarsest first) three [(si, ei), (sj, ej)] in the caller's order da, ad, dd
    (spiht_wrapper.py:232-250)"""
    data = bytes(data)
    out = np.empty((c, h, w), dtype=np.int32)
    meta = np.empty((8 * len(data) + 1, 8), dtype=np.int32)
    top = np.array([top_slice[0][0], top_slice[0][1], top_slice[1][0], top_slice[1][1]], dtype=np.int64)
    other = np.array([[[f[0][0], f[0][1], f[1][0], f[1][1]] for f in lvl] for lvl in other_slices], dtype=np.int64)
    other = np.ascontiguousarray(other.reshape(len(other_slices), 3, 4))
    _check(lib().spiht_ref_decode_with_metadata(data, len(data), int(n), c, h, w, ll_h, ll_w, top.ctypes.data,
                                                other.ctypes.data, len(other_slices), out.ctypes.data,
                                                meta.ctypes.data))
    return out, meta


def pyramid(x, ll_h, ll_w):
    x = np.ascontiguousarray(x, dtype=np.int32)
    c, h, w = x.shape
    nh, nw = h // 2, w // 2
    dp = np.zeros((c, nh, nw), np.uint8)
    lp = np.zeros((c, nh, nw), np.uint8)
    dpll = np.zeros((c, ll_h, ll_w), np.uint8)
    lpll = np.zeros((c, ll_h, ll_w), np.uint8)
    lib().spiht_model_pyramid(x.ctypes.data, c, h, w, ll_h, ll_w, dp.ctypes.data, lp.ctypes.data,
                              dpll.ctypes.data, lpll.ctypes.data)
    return dp, lp, dpll, lpll


def geom_ok(h, w, ll_h, ll_w):
    return bool(lib().spiht_model_geom_ok(h, w, ll_h, ll_w))


def get_offspring(i, j, h, w, ll_h, ll_w):
    buf = (ctypes.c_uint64 * 8)()
    if not lib().spiht_ref_get_offspring(i, j, h, w, ll_h, ll_w, buf):
        return None
    return [(buf[2 * q], buf[2 * q + 1]) for q in range(4)]


def set_bit(x, n, bit):
    return lib().spiht_ref_set_bit(x, n, int(bit))


def is_bit_set(x, n):
    return bool(lib().spiht_ref_is_bit_set(x, n))


def is_element_sig(x, n):
    return bool(lib().spiht_ref_is_element_sig(x, n))


def has_descendents_past_offspring(i, j, h, w):
    return bool(lib().spiht_ref_has_descendents_past_offspring(i, j, h, w))


def max_n_of(max_abs):
    return lib().spiht_ref_max_n(int(max_abs))
