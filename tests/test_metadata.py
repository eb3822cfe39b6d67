"""decode_with_metadata (/root/reference/src/encoder_decoder.rs:616-841,
src/lib.rs:47-56, spiht/spiht_wrapper.py:232-250): the (8 * nbytes + 1) x 8 table of decoder states.

CPU: the oracle's restatement (oracle/spiht_meta.c) is self-consistent with the plain decoder and reproduces
hand-checked rows.  GPU: the CUDA decoder's table equals the oracle's row for row, on every geometry class the
parallel parse treats differently (budgets that end in every kind of record, complete decodes with zero rows
left over, one-byte and empty streams), and the reference's own test holds (spiht/tests/test_spiht.py:19-28)."""
import numpy as np
import pytest

from conftest import synth_image


def _slices_args(h, w, wavelet, mode, level):
    from oracle import dwt_ref
    slices, enc_h, enc_w = dwt_ref.get_slices_and_h_w(h, w, wavelet, mode, level)
    top = [(slices[0][1].start or 0, slices[0][1].stop), (slices[0][2].start or 0, slices[0][2].stop)]
    other = [[[(lv[k][1].start, lv[k][1].stop), (lv[k][2].start, lv[k][2].stop)] for k in ("da", "ad", "dd")]
             for lv in slices[1:]]
    return top, other, enc_h, enc_w, slices[0][1].stop, slices[0][2].stop


def test_oracle_metadata_table_is_consistent(oracle):
    from oracle import wrapper_ref
    img = synth_image(3, 64, 96, 3)
    enc = wrapper_ref.encode_image(img, max_bits=6000)
    top, other, enc_h, enc_w, ll_h, ll_w = _slices_args(64, 96, "bior2.2", "reflect", None)
    rec, meta = oracle.decode_with_metadata(enc["encoded_bytes"], enc["max_n"], 3, enc_h, enc_w, ll_h, ll_w, top, other)
    assert np.array_equal(rec, oracle.decode(enc["encoded_bytes"], enc["max_n"], 3, enc_h, enc_w, ll_h, ll_w))
    assert meta.shape == (6001, 8) and meta.dtype == np.int32
    # the first bits are the LIP pass over the LL band, channel innermost: action 0, filter LL, depth = levels
    assert meta[0].tolist() == [0, -100000, -100000, 0, 0, len(other), enc["max_n"], 0]
    assert meta[1, 3] == 1 and meta[2, 3] == 2 and meta[3, 2] == int(np.float32(1 / ll_w) * np.float32(200000) - np.float32(100000))
    assert set(np.unique(meta[:, 0])) <= set(range(7))
    # a sign row (1 / 4) always follows its significance row (0 / 3) with the same coefficient
    for a_sig, a_sign in ((0, 1), (3, 4)):
        rows = np.nonzero(meta[:, 0] == a_sign)[0]
        assert (meta[rows - 1, 0] == a_sig).all() and (meta[rows - 1, 1:6] == meta[rows, 1:6]).all()
    # n never rises, refinement rows carry a significant coefficient
    assert (np.diff(meta[:, 6]) <= 0).all()
    assert (meta[meta[:, 0] == 6, 7] != 0).all()


CASES = [
    # (c, h, w), wavelet, mode, level, byte budgets
    ((3, 64, 96), "bior2.2", "reflect", None, [0, 1, 2, 7, 100, 750, 3000]),
    ((1, 128, 128), "bior2.2", "periodization", None, [5, 64, 1000, 10 ** 6]),
    ((2, 96, 160), "bior4.4", "symmetric", 2, [33, 500, 4000]),
    ((3, 256, 384), "bior2.2", "reflect", None, [12288]),
    ((1, 40, 56), "bior2.2", "reflect", 1, [10 ** 6]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,wavelet,mode,level,budgets", CASES)
def test_cuda_metadata_table_matches_oracle(oracle, shape, wavelet, mode, level, budgets):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import spiht.spiht as spiht_rs
    from oracle import wrapper_ref
    c, h, w = shape
    img = synth_image(c, h, w, 17)
    arr, ll_h, ll_w = wrapper_ref.forward_coeffs(img, wavelet=wavelet, mode=mode, level=level)
    if (ll_h | ll_w) & 1:
        pytest.skip("odd LL band")
    top, other, enc_h, enc_w, ll_h2, ll_w2 = _slices_args(h, w, wavelet, mode, level)
    assert (ll_h, ll_w) == (ll_h2, ll_w2)
    full, max_n = oracle.encode(arr, ll_h, ll_w, 10 ** 12)
    for nb in budgets:
        data = full[:nb]
        want_rec, want_meta = oracle.decode_with_metadata(data, max_n, c, enc_h, enc_w, ll_h, ll_w, top, other)
        rec, meta = spiht_rs.decode_with_metadata(data, max_n, c, enc_h, enc_w, ll_h, ll_w, top, other)
        assert rec.dtype == np.int32 and meta.dtype == np.int32 and meta.shape == (8 * len(data) + 1, 8)
        assert np.array_equal(rec, want_rec), nb
        bad = np.nonzero((meta != want_meta).any(axis=1))[0]
        assert len(bad) == 0, (nb, len(bad), int(bad[0]), meta[bad[0]].tolist(), want_meta[bad[0]].tolist())
        assert np.array_equal(rec, spiht_rs.decode(data, max_n, c, enc_h, enc_w, ll_h, ll_w))


@pytest.mark.gpu
def test_reference_test_encode_decode_with_metadata():
    """spiht/tests/test_spiht.py:19-28 on the reference's own images (those with an even LL band)"""
    import os
    import spiht
    from spiht.spiht_wrapper import SpihtSettings
    from spiht.utils import imload
    from spiht_b200 import _lib
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "images")
    done = 0
    for name in sorted(os.listdir(gold)):
        image = imload(os.path.join(gold, name))
        st = SpihtSettings()
        g = _lib.plan(image.shape[1], image.shape[2])
        encoded = spiht.encode_image(image, spiht_settings=st, max_bits=40000)
        if (g.ll_h | g.ll_w) & 1:
            with pytest.raises(_lib.SpihtB200Error):
                spiht.decode_image(encoded, st, return_metadata=True)
            continue
        decoded_image, spiht_metadata = spiht.decode_image(encoded, st, return_metadata=True)
        decoded_image_2 = spiht.decode_image(encoded, st, return_metadata=False)
        assert np.allclose(decoded_image, decoded_image_2)
        assert spiht_metadata.shape == (40000 + 1, 8)
        done += 1
    assert done >= 4


@pytest.mark.gpu
def test_metadata_errors():
    import spiht.spiht as spiht_rs
    from spiht_b200 import _lib
    with pytest.raises(_lib.SpihtB200Error):      # odd LL band
        spiht_rs.decode_with_metadata(b"\\x01\\x02", 3, 1, 24, 24, 3, 3, [(0, 3), (0, 3)], [])
    with pytest.raises(_lib.SpihtB200Error):      # the reference: assert!(ll_h > 1)
        spiht_rs.decode_with_metadata(b"\\x01", 3, 1, 8, 8, 1, 2, [(0, 1), (0, 2)], [])
