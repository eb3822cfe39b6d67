#!/usr/bin/env python
"""Encode and decode one image with the B200 SPIHT codec; save the reconstruction.

Command-line counterpart of the reference's encode_decode.py (same flags and defaults, /root/reference
encode_decode.py:16-26): the default level leaves a coarsest band about 8 pixels across (:33-38), the bit budget
is round(bpp * H * W) (:42-43), the reconstruction is clipped to [0, 1] and written as 8-bit (:79-83).

    python encode_decode.py tests/golden/images/zebra.jpg --bpp 0.5 --out reconstructed.png
    python encode_decode.py image.png --container out.spiht          # also write / re-read the stream container

The container (--container) is EncodingResult.to_dict() plus the settings, as a numpy .npz: every field the
decoder needs (bitstream bytes, image shape, max_n, level, wavelet, mode, quantisation scales, colour space).
"""
import math
import time
from argparse import ArgumentParser

import numpy as np

from spiht import EncodingResult, decode_image, encode_image
from spiht.spiht_wrapper import SpihtSettings, get_slices_and_h_w
from spiht.utils import imload


def build_parser():
    ap = ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("image_filename")
    ap.add_argument("--bpp", help="bits per pixel", type=float, default=0.1)
    ap.add_argument("--quantization_scale", default=255.0, type=float)
    ap.add_argument("--level", type=int, default=None,
                    help="wavedec2 level; default: the coarsest band is about 8 pixels across")
    ap.add_argument("--wavelet", help="wavedec2 wavelet", default="bior2.2", type=str)
    ap.add_argument("--mode", help="wavedec2 mode", default="reflect", type=str)
    ap.add_argument("--color_model", default="IPT", type=str)
    ap.add_argument("--per_channel_quant_scales", default="1., 0.2, 0.2", type=str)
    ap.add_argument("--out", help="save the reconstructed image to this path", type=str, default="reconstructed.png")
    ap.add_argument("--container", help="also write the encoded stream + settings to this .npz and decode from it",
                    type=str, default=None)
    return ap


def save_container(path, encoded: EncodingResult, settings: SpihtSettings):
    d = encoded.to_dict()
    d["encoding_result_encoded_bytes"] = np.frombuffer(encoded.encoded_bytes, dtype=np.uint8)
    d["encoding_result_level"] = -1 if encoded.level is None else int(encoded.level)
    d.update(settings_wavelet=settings.wavelet, settings_quantization_scale=settings.quantization_scale,
             settings_mode=settings.mode, settings_color_model=settings.color_model or "",
             settings_per_channel_quant_scales=np.asarray(settings.per_channel_quant_scales or [], dtype=np.float64))
    np.savez(path, **d)


def load_container(path):
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    d = {k: z[k] for k in z.files}
    er = {k: (v.item() if v.ndim == 0 else v) for k, v in d.items() if k.startswith("encoding_result_")}
    er["encoding_result_encoded_bytes"] = er["encoding_result_encoded_bytes"].tobytes()
    er["encoding_result_level"] = None if er["encoding_result_level"] < 0 else int(er["encoding_result_level"])
    encoded = EncodingResult.from_dict(er)
    pcs = d["settings_per_channel_quant_scales"].tolist()
    settings = SpihtSettings(wavelet=str(d["settings_wavelet"]), quantization_scale=float(d["settings_quantization_scale"]),
                             mode=str(d["settings_mode"]), color_model=str(d["settings_color_model"]) or None,
                             per_channel_quant_scales=pcs or None)
    return encoded, settings


def main(args):
    from PIL import Image
    im = imload(args.image_filename)
    c, h, w = im.shape
    level = args.level if args.level is not None else math.floor(min(math.log2(h / 8), math.log2(w / 8)))
    max_bits = round(args.bpp * h * w)
    scales = [float(x) for x in args.per_channel_quant_scales.split(",")]
    color = args.color_model if (args.color_model and c == 3) else None
    settings = SpihtSettings(quantization_scale=args.quantization_scale, mode=args.mode, wavelet=args.wavelet,
                             color_model=color, per_channel_quant_scales=scales[:c] if len(scales) >= c else None)
    print(f"Starting encoding of image {c} {h} {w}")
    st = time.time()
    encoded = encode_image(im, settings, level, max_bits)
    et = time.time()
    print(f"Encoding done in {et - st:.3f}s. Image encoded to {len(encoded.encoded_bytes) / 1024:.2f}kb")
    print(f"   levels: {encoded.level}")
    print(f"    max n: {encoded.max_n}")
    slices, enc_h, enc_w = get_slices_and_h_w(h, w, settings, encoded.level)
    print(f"ll_h ll_w: {(slices[0][1].stop, slices[0][2].stop)}")
    if args.container:
        save_container(args.container, encoded, settings)
        encoded, settings = load_container(args.container)
        print("Stream container written to and re-read from", args.container)
    st = time.time()
    dec_im = decode_image(encoded, settings)[:, :h, :w]
    et = time.time()
    print(f"Decoding done in {et - st:.3f}s. L2 distance: {((im - dec_im) ** 2).mean():.5f}")
    out = dec_im[0] if c == 1 else np.moveaxis(dec_im, 0, -1)
    out = (out.clip(0.0, 1.0) * 255).astype(np.uint8)
    Image.fromarray(out).save(args.out)
    print("Saved to ", args.out)
    return encoded, dec_im


if __name__ == "__main__":
    main(build_parser().parse_args())
