"""ctypes binding of libspiht_b200.so (the C ABI declared in include/spiht_b200.h).

There is no CPU fallback: if the library cannot be loaded, or no CUDA device is
usable, every compute entry point raises.
"""
import ctypes
import os
import threading

from . import build as _build

OK, EINVAL, ELL, EGEOM, ESHAPE, ECAP, ECUDA, ENOMEM, ELEVEL = range(9)
MAX_LEVELS = 24

WAVELET_IDS = {"bior2.2": 0, "bior4.4": 1, "bior6.8": 2,
               # the rest of PyWavelets' bior family (csrc/dwt_gen.cu)
               "bior1.1": 3, "bior1.3": 4, "bior1.5": 5, "bior2.4": 6, "bior2.6": 7, "bior2.8": 8,
               "bior3.1": 9, "bior3.3": 10, "bior3.5": 11, "bior3.7": 12, "bior3.9": 13, "bior5.5": 14}
MODE_IDS = {"reflect": 0, "symmetric": 1, "periodization": 2}
COLOR_NONE, COLOR_IPT = 0, 1
F32, F64, U8 = 0, 1, 2   # U8: forward direction only (pixels / 255 in float64, as utils.imload)


class Geom(ctypes.Structure):
    """struct spihtb_geom"""
    _fields_ = [
        ("h", ctypes.c_int32), ("w", ctypes.c_int32),
        ("wavelet", ctypes.c_int32), ("mode", ctypes.c_int32),
        ("levels", ctypes.c_int32),
        ("enc_h", ctypes.c_int32), ("enc_w", ctypes.c_int32),
        ("ll_h", ctypes.c_int32), ("ll_w", ctypes.c_int32),
        ("rec_h", ctypes.c_int32), ("rec_w", ctypes.c_int32),
        ("in_h", ctypes.c_int32 * MAX_LEVELS), ("in_w", ctypes.c_int32 * MAX_LEVELS),
        ("band_h", ctypes.c_int32 * MAX_LEVELS), ("band_w", ctypes.c_int32 * MAX_LEVELS),
        ("off_h", ctypes.c_int32 * MAX_LEVELS), ("off_w", ctypes.c_int32 * MAX_LEVELS),
    ]


class SpihtB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libspiht_b200 error {code}: {msg}")
        self.code = code


_lib = None
_lock = threading.Lock()

# every symbol include/spiht_b200.h declares
OPT_SCRATCH_COEFFS = 1

EXPORTS = (
    "spihtb_version", "spihtb_last_error", "spihtb_create", "spihtb_destroy", "spihtb_set_stream",
    "spihtb_sync", "spihtb_launch_count", "spihtb_plan", "spihtb_encode", "spihtb_decode",
    "spihtb_encode_coeffs", "spihtb_decode_coeffs", "spihtb_forward", "spihtb_inverse",
    "spihtb_encode_images", "spihtb_decode_images", "spihtb_stream_bound",
    "spihtb_profile_enable", "spihtb_profile_read", "spihtb_max_abs", "spihtb_convert_color", "spihtb_forward_path", "spihtb_decode_with_metadata",
    "spihtb_set_option", "spihtb_wavelet_filters",
)


def lib():
    """Load (building first when the sources are newer) libspiht_b200.so."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        try:
            if _build.needs_build():
                _build.build()
        except Exception as e:  # no nvcc on this box: use the prebuilt library if it is there
            if not os.path.exists(path):
                raise ImportError(
                    f"libspiht_b200.so is missing and could not be built ({e}); "
                    "the SPIHT hot path has no CPU fallback") from e
        L = ctypes.CDLL(path)
        vp, i32, u64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64, ctypes.c_double
        P = ctypes.POINTER
        L.spihtb_version.restype = ctypes.c_int
        L.spihtb_last_error.restype = ctypes.c_char_p
        L.spihtb_create.argtypes = [ctypes.c_int, P(vp)]
        L.spihtb_destroy.argtypes = [vp]
        L.spihtb_set_stream.argtypes = [vp, vp]
        L.spihtb_sync.argtypes = [vp]
        L.spihtb_launch_count.argtypes = [vp]
        L.spihtb_launch_count.restype = ctypes.c_int64
        L.spihtb_forward_path.argtypes = [vp]
        L.spihtb_plan.argtypes = [i32, i32, i32, i32, i32, P(Geom)]
        L.spihtb_wavelet_filters.argtypes = [i32, P(i32), P(dbl), P(dbl)]
        L.spihtb_encode.argtypes = [vp, vp, i32, i32, i32, i32, i32, u64, P(vp), P(u64), P(i32)]
        L.spihtb_decode.argtypes = [vp, ctypes.c_char_p, u64, i32, i32, i32, i32, i32, i32, vp]
        L.spihtb_decode_with_metadata.argtypes = [vp, ctypes.c_char_p, u64, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp]
        L.spihtb_encode_coeffs.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, u64, vp, vp, u64, vp, vp, vp]
        L.spihtb_decode_coeffs.argtypes = [vp, vp, u64, vp, vp, i32, i32, i32, i32, i32, i32, vp]
        L.spihtb_forward.argtypes = [vp, vp, i32, i32, i32, P(Geom), i32, P(dbl), dbl, vp]
        L.spihtb_inverse.argtypes = [vp, vp, i32, i32, P(Geom), i32, P(dbl), dbl, vp, i32]
        L.spihtb_encode_images.argtypes = [vp, vp, i32, i32, i32, P(Geom), i32, P(dbl), dbl, u64, vp, vp,
                                           vp, u64, vp, vp, vp]
        L.spihtb_decode_images.argtypes = [vp, vp, u64, vp, vp, i32, i32, P(Geom), i32, P(dbl), dbl, vp, vp, i32]
        L.spihtb_max_abs.argtypes = [vp, vp, i32, u64, vp]
        L.spihtb_convert_color.argtypes = [vp, vp, i32, i32, u64, i32, i32, vp, i32]
        L.spihtb_stream_bound.argtypes = [i32, i32, i32, i32, i32]
        L.spihtb_stream_bound.restype = u64
        L.spihtb_profile_enable.argtypes = [vp, ctypes.c_int]
        L.spihtb_profile_read.argtypes = [vp, P(dbl), P(ctypes.c_int64), ctypes.c_int]
        L.spihtb_set_option.argtypes = [vp, i32, ctypes.c_int64]
        for name in EXPORTS:
            getattr(L, name)   # every declared symbol must resolve
        _lib = L
        return _lib


def check(rc):
    """Map a status code to the exception the reference would raise."""
    if rc == OK:
        return
    msg = lib().spihtb_last_error().decode("utf-8", "replace")
    if rc in (EINVAL, ELEVEL):
        raise ValueError(msg)
    if rc in (ELL, EGEOM):
        # the reference panics here (assert! / index out of bounds -> pyo3 PanicException)
        raise SpihtB200Error(rc, msg)
    if rc == ESHAPE:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise SpihtB200Error(rc, msg)


class Context:
    """Owns a spihtb_ctx (streams + device workspaces) for one CUDA device."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        self.device = device
        check(lib().spihtb_create(int(device), ctypes.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def set_stream(self, stream_ptr):
        check(lib().spihtb_set_stream(self._h, ctypes.c_void_p(stream_ptr or None)))

    def sync(self):
        check(lib().spihtb_sync(self._h))

    def launch_count(self):
        return int(lib().spihtb_launch_count(self._h))

    def set_option(self, option, value):
        check(lib().spihtb_set_option(self._h, int(option), int(value)))

    def forward_path(self):
        """12 when the last forward transform ran levels 1+2 in the fused TMA kernel, else 1"""
        return int(lib().spihtb_forward_path(self._h))

    STAGES = ("dwt_fwd_level1", "dwt_fwd_rest", "pyramid_base", "pyramid_rest", "spiht_encode", "spiht_decode",
              "dwt_inv_coarse", "dwt_inv_level1")

    def profile(self, enable=True):
        check(lib().spihtb_profile_enable(self._h, int(bool(enable))))

    def profile_read(self, reset=True):
        """{stage: (total_ms, intervals)} since the last reset (synchronises)"""
        n = len(self.STAGES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        check(lib().spihtb_profile_read(self._h, ms, cnt, int(bool(reset))))
        return {name: (ms[i], int(cnt[i])) for i, name in enumerate(self.STAGES)}

    def close(self):
        if self._h:
            lib().spihtb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ctxs = {}
_tls = threading.local()


def get_context(device=0):
    """One context per (thread, device)."""
    d = getattr(_tls, "ctxs", None)
    if d is None:
        d = _tls.ctxs = {}
    if device not in d:
        d[device] = Context(device)
    return d[device]


def plan(h, w, wavelet="bior2.2", mode="reflect", level=None):
    """spihtb_plan: band geometry without touching the GPU."""
    if wavelet not in WAVELET_IDS:
        raise ValueError(f"Unknown wavelet name '{wavelet}', supported: {sorted(WAVELET_IDS)}")
    if mode not in MODE_IDS:
        raise ValueError(f"Unknown mode name '{mode}', supported: {sorted(MODE_IDS)}")
    if level is not None and level < 0:
        raise ValueError(f"Level value of {level} is too low . Minimum level is 0.")
    g = Geom()
    check(lib().spihtb_plan(int(h), int(w), WAVELET_IDS[wavelet], MODE_IDS[mode],
                            -1 if level is None else int(level), ctypes.byref(g)))
    return g
