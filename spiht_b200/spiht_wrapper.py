"""Drop-in for the reference's spiht/spiht_wrapper.py: same names, argument
order, defaults and result fields; the colour transform, DWT, quantiser and
SPIHT coder all run on the GPU through libspiht_b200.so.

Added on top of the reference surface: `encode_images` / `decode_images`, the
batched forms of `encode_image` / `decode_image` (one library call per batch of
same-shaped images; mixed shapes are grouped).
"""
from dataclasses import asdict, dataclass
from typing import Any, List, Optional, Sequence, Tuple, Union

import threading

import numpy as np

from . import _lib
from . import spiht as spiht_rs
from .color_models import normalise as _norm_color


def quantize(arr, q_scale=10.):
    """spiht_wrapper.py:9-11"""
    arr = arr * q_scale
    return arr.astype(np.int32)


def dequantize(arr, q_scale=10.):
    """spiht_wrapper.py:13-14"""
    return arr / q_scale


ENCODER_DECODER_VERSION = "0.0.2"

# max_bits=None in the reference (spiht_wrapper.py:174-176)
_VERY_LARGE = 99999999999999999


@dataclass
class SpihtSettings:
    """spiht_wrapper.py:20-63.  Parameters of the codec that are not particular
    to one image: wavelet, quantisation scale, boundary mode, colour model and
    optional per-channel quantisation scales."""
    wavelet: str = 'bior2.2'
    quantization_scale: float = 50.0
    mode: str = 'reflect'
    color_model: Optional[str] = None
    per_channel_quant_scales: Optional[List[float]] = None


@dataclass
class EncodingResult:
    """spiht_wrapper.py:65-89.
    encoded_bytes: bytes returned by the spiht encoder
    h, w, c: height, width, channels of the original image
    max_n: the starting n parameter used in the spiht encoder
    level: optional number of DWT levels
    """
    encoded_bytes: bytes
    h: int
    w: int
    c: int
    max_n: int
    level: Optional[int]
    _encoding_version: str = ENCODER_DECODER_VERSION

    def to_dict(self):
        return {f"encoding_result_{k}": v for k, v in asdict(self).items()}

    @staticmethod
    def from_dict(d):
        d = {k.removeprefix('encoding_result_'): v for k, v in d.items() if k.startswith('encoding_result_')}
        return EncodingResult(**d)


def _geom(h, w, spiht_settings, level):
    return _lib.plan(h, w, spiht_settings.wavelet, spiht_settings.mode, level)


def get_slices_and_h_w(h: int, w: int, spiht_settings: SpihtSettings, level: Optional[int]):
    """spiht_wrapper.py:92-139: the slices pywt.coeffs_to_array would use, and the
    height and width of the coefficient array."""
    g = _geom(h, w, spiht_settings, level)
    start_h, start_w = g.ll_h, g.ll_w
    slices: List[Any] = [(slice(None), slice(start_h), slice(start_w))]
    for l in range(g.levels - 1, -1, -1):
        bh, bw = g.band_h[l], g.band_w[l]
        slices.append({
            "ad": (slice(None), slice(0, bh), slice(start_w, start_w + bw)),
            "da": (slice(None), slice(start_h, start_h + bh), slice(0, bw)),
            "dd": (slice(None), slice(start_h, start_h + bh), slice(start_w, start_w + bw)),
        })
        start_h += bh
        start_w += bw
    return slices, start_h, start_w


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("spiht_b200 needs a CUDA device: the codec has no CPU fallback")
    return torch


def _to_device_pixels(image):
    """numpy / torch image batch -> contiguous CUDA tensor; float32 and uint8 kept, everything else float64.
    uint8 is image data as stored on disk: the library scales it by 1 / 255 in float64, exactly what the
    reference's loader does (utils.py:12-20), so encode_image(raw_uint8) == encode_image(imload-style floats)."""
    torch = _torch()
    if isinstance(image, torch.Tensor):
        t = image
        if t.dtype not in (torch.float32, torch.float64, torch.uint8):
            t = t.to(torch.float64)
        return t.cuda().contiguous()
    a = np.asarray(image)
    if a.dtype not in (np.float32, np.uint8):
        a = a.astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


_PIPE_CHUNK = 32   # images per host->device copy in the pipelined host path, at most
_PIPE_CHUNK_BYTES = 400 << 20   # ... and about this many bytes (32 float32 RGB images of 1024^2)


def _pipe_chunk(per_image_bytes: int) -> int:
    return max(1, min(_PIPE_CHUNK, _PIPE_CHUNK_BYTES // max(1, per_image_bytes)))


_pinned_cache = {}


def _pinned(name, shape, dtype):
    """A reusable pinned host buffer (page-locking memory costs milliseconds; the pipelined path below needs a few per
    call).  One buffer per (name, shape, dtype); the caller copies what it keeps before the next call."""
    torch = _torch()
    key = (threading.get_ident(), name, tuple(shape), dtype)   # one set per calling thread
    buf = _pinned_cache.get(key)
    if buf is None:
        if len(_pinned_cache) > 16:
            _pinned_cache.clear()
        buf = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        _pinned_cache[key] = buf
    return buf


def _encode_host_pipelined(images, g, spiht_settings, budget, level):
    """Host pixels in, host bytes out, for a large batch: the batch goes to the device in chunks on a copy
    stream (two device buffers) while the previous chunk is transformed and coded on the current stream, so
    the PCIe transfer -- by far the longest part of an end-to-end encode -- hides everything else.  The streams of
    a finished chunk come back on a third stream into pinned memory (the link is full duplex) and the host builds
    that chunk's results while the later chunks are still in flight; only the last chunk's return trip is exposed.
    Pinned host memory gives a truly asynchronous copy; pageable memory still works (the copy then blocks)."""
    from . import batch
    torch = _torch()
    t = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
    if t.dtype not in (torch.float32, torch.float64, torch.uint8):
        t = t.to(torch.float64)
    t = t.contiguous()
    B, c, h, w = t.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    n = _pipe_chunk(c * h * w * t.element_size())
    stride = batch.stream_stride(budget, c, g)
    bufs = [torch.empty((n, c, h, w), dtype=t.dtype, device=dev) for _ in range(2)]
    coeffs = torch.empty((n, c, g.enc_h, g.enc_w), dtype=torch.int32, device=dev)
    streams = torch.empty((B, stride), dtype=torch.uint8, device=dev)
    rows_h = _pinned("rows", (B, stride), torch.uint8)
    nbits_h = _pinned("nbits", (B,), torch.int64)
    max_n_h = _pinned("max_n", (B,), torch.int32)
    status_h = _pinned("status", (B,), torch.int32)
    main = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)
    back = torch.cuda.Stream(dev)
    copy.wait_stream(main)
    done = [None, None]   # per device buffer: event after the chunk that last read it
    results: List[Optional[EncodingResult]] = [None] * B
    pending = []          # (lo, hi, event after the chunk's results reached the host)

    def harvest(lo, hi, ev):
        ev.synchronize()
        if int(status_h[lo:hi].max()) != 0:
            raise _lib.SpihtB200Error(_lib.ECAP, "bitstream row too small")
        nb = (nbits_h[lo:hi].numpy() + 7) // 8
        rows = rows_h.numpy()
        mx = max_n_h.numpy()
        for b in range(lo, hi):
            results[b] = EncodingResult(rows[b, :int(nb[b - lo])].tobytes(), h, w, c, int(mx[b]), level)

    for i, lo in enumerate(range(0, B, n)):
        hi = min(B, lo + n)
        buf = bufs[i & 1][:hi - lo]
        with torch.cuda.stream(copy):
            if done[i & 1] is not None:
                copy.wait_event(done[i & 1])
            buf.copy_(t[lo:hi], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        main.wait_event(ready)
        _, nbits, max_n, status, _ = batch.encode_images(buf, g, spiht_settings, budget, out_stride=stride,
                                                         coeffs=coeffs[:hi - lo], out=streams[lo:hi])
        done[i & 1] = torch.cuda.Event()
        done[i & 1].record(main)
        with torch.cuda.stream(back):
            back.wait_event(done[i & 1])
            rows_h[lo:hi].copy_(streams[lo:hi], non_blocking=True)
            nbits_h[lo:hi].copy_(nbits, non_blocking=True)
            max_n_h[lo:hi].copy_(max_n, non_blocking=True)
            status_h[lo:hi].copy_(status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(back)
        # the tensors of this chunk must outlive their copies on the other stream
        nbits.record_stream(back)
        max_n.record_stream(back)
        status.record_stream(back)
        pending.append((lo, hi, ev))
        if len(pending) > 1:             # build the previous chunk's results while this one is in flight
            harvest(*pending.pop(0))
    while pending:
        harvest(*pending.pop(0))
    main.wait_stream(back)
    return results  # type: ignore[return-value]


def encode_image(image: np.ndarray, spiht_settings: SpihtSettings = SpihtSettings(), level: Optional[int] = None,
                 max_bits: Optional[int] = None):
    """spiht_wrapper.py:142-189.  Takes the DWT of the image, quantises the
    coefficients and encodes them.

    image: 3D array (C,H,W) of floating point pixel values (numpy, or a torch tensor)
    Returns EncodingResult
    """
    if image.ndim != 3:
        raise ValueError('image ndim must be 3: c,h,w')
    return encode_images(image[None], spiht_settings, level, max_bits)[0]


def encode_images(images, spiht_settings: SpihtSettings = SpihtSettings(), level: Optional[int] = None,
                  max_bits=None) -> List[EncodingResult]:
    """Batched encode_image.  images: array/tensor (B,C,H,W), or a sequence of
    (C,H,W) images (shapes may differ; equal shapes are encoded together).
    max_bits: None (no limit), one budget for every image, or a sequence with one budget per image."""
    from . import batch
    torch = _torch()
    per_image = None
    if max_bits is not None and not np.isscalar(max_bits):
        per_image = [int(m) for m in max_bits]
        if len(per_image) != len(images):
            raise ValueError("max_bits must be a scalar or have one entry per image")
        if any(m <= 0 for m in per_image):
            raise ValueError("per-image max_bits must be positive")
    if isinstance(images, (list, tuple)):
        for im in images:
            if im.ndim != 3:
                raise ValueError('image ndim must be 3: c,h,w')
        groups = {}
        for idx, im in enumerate(images):
            groups.setdefault((tuple(im.shape), str(im.dtype)), []).append(idx)
        results: List[Optional[EncodingResult]] = [None] * len(images)
        for _, idxs in groups.items():
            ims = [images[i] for i in idxs]
            stack = torch.stack(ims) if isinstance(ims[0], torch.Tensor) else np.stack(ims)
            mb = max_bits if per_image is None else [per_image[i] for i in idxs]
            for i, r in zip(idxs, encode_images(stack, spiht_settings, level, mb)):
                results[i] = r
        return results  # type: ignore[return-value]
    if images.ndim != 4:
        raise ValueError('images ndim must be 4: b,c,h,w')
    B, c, h, w = images.shape
    _norm_color(spiht_settings.color_model)
    g = _geom(h, w, spiht_settings, level)
    if max_bits is None:
        max_bits = _VERY_LARGE
    bound = 8 * int(_lib.lib().spihtb_stream_bound(c, g.enc_h, g.enc_w, g.ll_h, g.ll_w))
    if per_image is not None:
        # one budget per image (device array); rows sized for the largest, never beyond the full-encode bound
        pixels = _to_device_pixels(images)
        budgets = torch.tensor([min(m, bound) for m in per_image], dtype=torch.int64, device=pixels.device)
        stride = batch.stream_stride(min(max(per_image), bound), c, g)
        streams, nbits, max_n, status, _ = batch.encode_images(pixels, g, spiht_settings, budgets, out_stride=stride)
        if max(per_image) >= bound:
            status = torch.zeros_like(status)      # a budget at the bound cannot truncate
    else:
        budget = int(max_bits)
        host_side = not (isinstance(images, torch.Tensor) and images.is_cuda)
        if host_side and 0 < budget <= bound and B >= 2 * _pipe_chunk(
                c * h * w * (images.element_size() if isinstance(images, torch.Tensor) else images.dtype.itemsize)):
            return _encode_host_pipelined(images, g, spiht_settings, budget, level)
        pixels = _to_device_pixels(images)
        if budget == 0 or budget > bound:
            # untruncated: run the transform first, size the rows from the largest coefficient
            coeffs = batch.forward(pixels, g, spiht_settings)
            planes = int(batch.max_abs(coeffs).cpu().numpy().max()).bit_length() + 1
            stride = (bound // 31 * min(planes + 1, 31) // 8 + 64) // 8 * 8
            streams, nbits, max_n, status = batch.encode_coeffs(coeffs, g.ll_h, g.ll_w, budget, out_stride=stride)
        else:
            streams, nbits, max_n, status, _ = batch.encode_images(pixels, g, spiht_settings, budget)
    nbits_h = nbits.cpu().numpy()
    max_n_h = max_n.cpu().numpy()
    if int(status.max().item()) != 0:
        raise _lib.SpihtB200Error(_lib.ECAP, "bitstream row too small for an untruncated encode")
    nbytes = (nbits_h + 7) // 8
    rows = streams[:, :int(nbytes.max()) if B else 0].cpu().numpy()
    return [EncodingResult(rows[b, :int(nbytes[b])].tobytes(), h, w, c, int(max_n_h[b]), level) for b in range(B)]


def decode_image(encoding_result: EncodingResult, spiht_settings: SpihtSettings,
                 return_metadata: bool = False) -> Union[np.ndarray, Tuple[np.ndarray, np.ndarray]]:
    """spiht_wrapper.py:192-216.  Decodes the encoding_result to pixel values
    (float64 array (C,H',W'); H' = H + 1 for odd H, as pywt.waverec2 returns)."""
    if return_metadata:
        # spiht_wrapper.py:203-216: the metadata path goes through the raw coder and the host-side inverse
        out = decode_rec_array(encoding_result, spiht_settings, return_metadata=True)
        image = decode_from_rec_arr(out["rec_arr"], out["h"], out["w"], out["level"], spiht_settings, out["slices"])
        return image, out["spiht_metadata"]
    return decode_images([encoding_result], spiht_settings)[0]


def decode_images(encoding_results: Sequence[EncodingResult], spiht_settings: SpihtSettings, as_numpy: bool = True):
    """Batched decode_image.  Returns a list of float64 (C,H',W') arrays in input order."""
    from . import batch
    torch = _torch()
    _norm_color(spiht_settings.color_model)
    groups = {}
    for idx, er in enumerate(encoding_results):
        if er._encoding_version != ENCODER_DECODER_VERSION:
            raise ValueError(er._encoding_version)
        groups.setdefault((er.c, er.h, er.w, er.level), []).append(idx)
    out: List[Any] = [None] * len(encoding_results)
    for (c, h, w, level), idxs in groups.items():
        g = _geom(h, w, spiht_settings, level)
        ers = [encoding_results[i] for i in idxs]
        lens = np.array([len(e.encoded_bytes) for e in ers], dtype=np.int64)
        stride = (int(lens.max()) + 15) // 8 * 8
        host = np.zeros((len(ers), stride), dtype=np.uint8)
        for r, e in enumerate(ers):
            host[r, :lens[r]] = np.frombuffer(e.encoded_bytes, dtype=np.uint8)
        streams = torch.from_numpy(host).cuda()
        nbytes = torch.from_numpy(lens).cuda()
        max_n = torch.tensor([e.max_n for e in ers], dtype=torch.int32).cuda()
        pix, _ = batch.decode_images(streams, nbytes, max_n, c, g, spiht_settings, dtype=torch.float64,
                                     scratch_coeffs=True)   # only the pixels leave this function
        pix = pix.cpu().numpy() if as_numpy else pix
        for r, i in enumerate(idxs):
            out[i] = pix[r]
    return out


def decode_image_prefixes(encoding_result: EncodingResult, spiht_settings: SpihtSettings, byte_lengths: Sequence[int],
                          return_coeffs: bool = False, as_numpy: bool = True):
    """Progressive (multi-rate) decode: the images an embedded stream gives when it is cut after each of
    `byte_lengths` bytes, all decoded by ONE batched library call (one CTA per prefix).  This is the loop of
    the reference's make_gif.py:46-61 (`encoded.encoded_bytes = original_bytes[:byte_len]; decode_image(...)`
    per frame) as a batch.  Returns a list of float64 (C,H',W') images in the order of byte_lengths; with
    return_coeffs also the int32 coefficient arrays (what make_gif.py:61 gets from spiht.decode)."""
    from . import batch
    torch = _torch()
    er = encoding_result
    if er._encoding_version != ENCODER_DECODER_VERSION:
        raise ValueError(er._encoding_version)
    _norm_color(spiht_settings.color_model)
    lens = np.asarray([min(max(int(n), 0), len(er.encoded_bytes)) for n in byte_lengths], dtype=np.int64)
    g = _geom(er.h, er.w, spiht_settings, er.level)
    if len(lens) == 0:
        return ([], []) if return_coeffs else []
    stride = (int(lens.max()) + 15) // 8 * 8
    row = np.zeros((stride,), dtype=np.uint8)
    row[:int(lens.max())] = np.frombuffer(er.encoded_bytes[:int(lens.max())], dtype=np.uint8)
    dev_row = torch.from_numpy(row).cuda()
    # every row holds the whole (longest) prefix: the decoder stops at 8 * nbytes[b] bits, so rows need no zero tail
    streams = dev_row[None, :].expand(len(lens), stride).contiguous()
    nbytes = torch.from_numpy(lens).cuda()
    max_n = torch.full((len(lens),), int(er.max_n), dtype=torch.int32, device=streams.device)
    pix, coeffs = batch.decode_images(streams, nbytes, max_n, er.c, g, spiht_settings, dtype=torch.float64)
    if as_numpy:
        pix = pix.cpu().numpy()
        coeffs = coeffs.cpu().numpy() if return_coeffs else None
    images = [pix[i] for i in range(len(lens))]
    if return_coeffs:
        return images, [coeffs[i] for i in range(len(lens))]
    return images


def decode_rec_array(encoding_result: EncodingResult, spiht_settings: SpihtSettings, return_metadata: bool = False):
    """spiht_wrapper.py:218-257: bitstream -> int32 coefficient array (+ geometry)"""
    encoded_bytes = encoding_result.encoded_bytes
    h = encoding_result.h
    w = encoding_result.w
    c = encoding_result.c
    max_n = encoding_result.max_n
    level = encoding_result.level

    if encoding_result._encoding_version != ENCODER_DECODER_VERSION:
        raise ValueError(encoding_result._encoding_version)

    slices, enc_h, enc_w = get_slices_and_h_w(h, w, spiht_settings, level)
    ll_h, ll_w = slices[0][1].stop, slices[0][2].stop

    if return_metadata:
        # spiht_wrapper.py:232-250: band rectangles in the order da, ad, dd, coarsest level first
        top_slice = [(slices[0][1].start or 0, slices[0][1].stop), (slices[0][2].start or 0, slices[0][2].stop)]
        other_slices = []
        for slice_level in slices[1:]:
            other_slices.append([[(slice_level[key][1].start, slice_level[key][1].stop),
                                  (slice_level[key][2].start, slice_level[key][2].stop)] for key in ["da", "ad", "dd"]])
        rec_arr, spiht_metadata = spiht_rs.decode_with_metadata(encoded_bytes, max_n, c, enc_h, enc_w, ll_h, ll_w,
                                                                top_slice, other_slices)
    else:
        rec_arr = spiht_rs.decode(encoded_bytes, max_n, c, enc_h, enc_w, ll_h, ll_w)
        spiht_metadata = None
    return dict(rec_arr=rec_arr, slices=slices, spiht_metadata=spiht_metadata, h=h, w=w, level=level)


def decode_from_rec_arr(rec_arr: np.ndarray, h: int, w: int, level, spiht_settings: SpihtSettings, slices=None):
    """spiht_wrapper.py:259-281: int32 coefficient array -> image"""
    from . import batch
    torch = _torch()
    _norm_color(spiht_settings.color_model)
    g = _geom(h, w, spiht_settings, level)
    rec_arr = np.asarray(rec_arr)
    if rec_arr.shape[1:] != (g.enc_h, g.enc_w):
        raise ValueError(f"rec_arr has shape {rec_arr.shape}, expected (c, {g.enc_h}, {g.enc_w})")
    coeffs = torch.from_numpy(np.ascontiguousarray(rec_arr.astype(np.int32)))[None].cuda()
    return batch.inverse(coeffs, g, spiht_settings, dtype=torch.float64)[0].cpu().numpy()
