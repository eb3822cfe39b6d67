"""Golden fixtures for the reference's own image fixtures (tests/golden/images/ = /root/reference/images/*.jpg).

The reference's integration tests run its codec over these files without asserting values
(spiht/tests/test_spiht.py:10-17, spiht/tests/test_rust.py:11-56) and the reference cannot run in the build
container (Rust coder, PyWavelets, colour-science absent), so the numbers recorded here come from the CPU
oracle (`source: oracle`): they pin the oracle against drift and give the GPU tests a second, oracle-free
anchor on real photographs.  Scenarios:

  * config1      BASELINE.json configs[0]: zebra.jpg, SpihtSettings() defaults, 1.0 bpp, encode + decode
  * default      test_spiht.py:10-17: every image, SpihtSettings(), max_bits=None (full encode), decode
  * rust_test    test_rust.py:11-56: skiing.jpg, bior4.4, mode='symmetric', q=50, raw encode / decode of the
                 quantised coefficient array at an unlimited budget; `lossless` records whether the reference's
                 assertion array_equal(coeffs_arr, rec_arr) holds (its author notes it fails for some
                 geometries: here ll_w = 19 is odd) and `mismatches` how many coefficients differ
  * ipt          demonstrate.py:17-32 settings (IPT, [100,20,20], q=1) on zebra.jpg at 0.5 bpp

Run from the repo root:  python tests/golden/make_golden_images.py
"""
import glob
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import dwt_ref, spiht_oracle, wrapper_ref  # noqa: E402
from spiht_b200.utils import imload  # noqa: E402


def psnr(a, b):
    return float(10 * np.log10(1.0 / np.mean((a - b) ** 2)))


def stream_entry(enc):
    data = enc["encoded_bytes"]
    return {"nbytes": len(data), "max_n": int(enc["max_n"]), "sha256": hashlib.sha256(data).hexdigest(),
            "head_hex": data[:24].hex()}


def main():
    out = {"source": "oracle (oracle/spiht_ref.c + oracle/dwt_ref.py + oracle/ipt_ref.py)", "images": {}}
    files = sorted(glob.glob(os.path.join(HERE, "images", "*.jpg")))
    for f in files:
        name = os.path.basename(f)
        im = imload(f)
        c, h, w = im.shape
        enc = wrapper_ref.encode_image(im)
        rec = wrapper_ref.decode_image(enc)
        e = {"shape": [c, h, w], "pixel_sha256": hashlib.sha256(np.ascontiguousarray(im).tobytes()).hexdigest(),
             "default": dict(stream_entry(enc), psnr_db=round(psnr(rec[:, :h, :w], im), 6))}
        out["images"][name] = e
    # config 1
    im = imload(os.path.join(HERE, "images", "zebra.jpg"))
    c, h, w = im.shape
    mb = int(h * w * 1.0)
    enc = wrapper_ref.encode_image(im, max_bits=mb)
    rec = wrapper_ref.decode_image(enc)
    out["config1"] = dict(stream_entry(enc), image="zebra.jpg", max_bits=mb, psnr_db=round(psnr(rec[:, :h, :w], im), 6))
    # test_rust.py scenario
    im = imload(os.path.join(HERE, "images", "skiing.jpg"))
    co = dwt_ref.wavedec2(im, "bior4.4", "symmetric", None)
    arr = (dwt_ref.coeffs_to_array(co) * 50).astype(np.int32)
    ll_h, ll_w = co[0].shape[1:]
    data, max_n = spiht_oracle.encode(arr, ll_h, ll_w, 999999999999)
    rec = spiht_oracle.decode(data, max_n, *arr.shape, ll_h, ll_w)
    out["rust_test"] = {"image": "skiing.jpg", "coeff_shape": list(arr.shape), "ll": [int(ll_h), int(ll_w)],
                        "nbytes": len(data), "max_n": int(max_n), "sha256": hashlib.sha256(data).hexdigest(),
                        "coeff_sha256": hashlib.sha256(arr.tobytes()).hexdigest(),
                        "lossless": bool(np.array_equal(arr, rec)), "mismatches": int((arr != rec).sum())}
    # demonstrate.py settings
    im = imload(os.path.join(HERE, "images", "zebra.jpg"))
    kw = dict(quantization_scale=1.0, color_model="IPT", per_channel_quant_scales=[100, 20, 20])
    mb = int(h * w * 0.5)
    enc = wrapper_ref.encode_image(im, max_bits=mb, **kw)
    rec = wrapper_ref.decode_image(enc, **kw)
    out["ipt"] = dict(stream_entry(enc), image="zebra.jpg", max_bits=mb, psnr_db=round(psnr(rec[:, :h, :w], im), 6))
    with open(os.path.join(HERE, "images.json"), "w") as fjs:
        json.dump(out, fjs, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True)[:1500])


if __name__ == "__main__":
    main()
