"""batched progressive (multi-rate) decode of one embedded stream
(/root/reference/make_gif.py:46-61) and the command-line encode / decode with its on-disk container
(/root/reference/encode_decode.py; EncodingResult.to_dict / from_dict, spiht_wrapper.py:83-89)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.gpu
def test_batched_prefix_decode_matches_oracle_per_prefix():
    import spiht
    from spiht.spiht_wrapper import SpihtSettings, get_slices_and_h_w
    from spiht.utils import imload
    from oracle import spiht_oracle, wrapper_ref
    # make_gif.py:14-21 settings without the colour model (bit-exact path), on a reference photograph
    st = SpihtSettings(quantization_scale=75, wavelet="bior4.4", mode="symmetric")
    im = imload(os.path.join(GOLD, "images", "porter.jpg"))
    c, h, w = im.shape
    level = 4
    enc = spiht.encode_image(im, st, level, max_bits=int(0.7 * h * w))
    bpps = np.linspace(0.01, 0.7 ** 0.5, 12) ** 2               # make_gif.py:24-25
    lens = [max(int(b * h * w / 8), 1) for b in bpps] + [0, 3, len(enc.encoded_bytes), 10 ** 9]
    images, coeffs = spiht.decode_image_prefixes(enc, st, lens, return_coeffs=True)
    assert len(images) == len(lens)
    slices, enc_h, enc_w = get_slices_and_h_w(h, w, st, level)
    ll_h, ll_w = slices[0][1].stop, slices[0][2].stop
    prev_err = None
    for n, img, co in zip(lens, images, coeffs):
        data = enc.encoded_bytes[:n]
        want = spiht_oracle.decode(data, enc.max_n, c, enc_h, enc_w, ll_h, ll_w)
        assert np.array_equal(co, want), n                      # bit-exact coefficients for every prefix
        ref = wrapper_ref.inverse_coeffs(want, h, w, "bior4.4", "symmetric", level, 75)
        assert np.abs(img - ref).max() < 1e-9
        # and identical to the one-at-a-time path of the reference's loop
        one = spiht.decode_image(spiht.EncodingResult(data, h, w, c, enc.max_n, level), st)
        assert np.array_equal(one, img)
    errs = [float(((img[:, :h, :w] - im) ** 2).mean()) for img in images[:12]]
    assert errs[-1] < errs[0] and errs[-1] < 0.01               # quality rises with the rate


@pytest.mark.gpu
def test_cli_encode_decode_and_container_roundtrip(tmp_path):
    sys.path.insert(0, ROOT)
    import encode_decode as cli
    from PIL import Image
    out = tmp_path / "rec.png"
    box = tmp_path / "zebra_stream"
    args = cli.build_parser().parse_args([os.path.join(GOLD, "images", "zebra.jpg"), "--bpp", "0.5", "--out", str(out),
                                          "--container", str(box)])
    encoded, dec = cli.main(args)
    assert encoded.level == 5 and (encoded.h, encoded.w, encoded.c) == (256, 384, 3)     # floor(log2(256 / 8)), encode_decode.py:33-38
    assert len(encoded.encoded_bytes) == round(0.5 * 256 * 384) // 8
    rec = np.asarray(Image.open(out))
    assert rec.shape == (256, 384, 3)
    # the container holds everything the decoder needs
    enc2, st2 = cli.load_container(str(box))
    assert enc2 == encoded and st2.color_model == "IPT" and st2.per_channel_quant_scales == [1.0, 0.2, 0.2]
    import spiht
    assert np.array_equal(spiht.decode_image(enc2, st2)[:, :256, :384], dec)
    from spiht_b200.utils import imload
    im = imload(os.path.join(GOLD, "images", "zebra.jpg"))
    assert 10 * np.log10(1.0 / np.mean((dec - im) ** 2)) > 15     # the CLI's defaults (q = 255, IPT, [1, .2, .2]) at 0.5 bpp
