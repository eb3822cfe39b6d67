"""Pins the CPU oracle (oracle/spiht_ref.c) to every result the reference's own
tests hold for the SPIHT path (src/encoder_decoder.rs:845-1025) and to the
hand-traced known-answer vectors of SURVEY.md section 4."""
import numpy as np
import pytest

from oracle import spiht_oracle as o


# ---- src/encoder_decoder.rs:851-862 test_is_bit_set
def test_is_bit_set():
    assert o.is_bit_set(32, 5)
    for n in range(5):
        assert not o.is_bit_set(32, n)
    assert o.is_bit_set(-69, 6)
    assert not o.is_bit_set(3590854, 8)


# ---- :988-991 test_base_sig
def test_base_sig():
    assert (1 << 5) + (1 << 6) == 96
    assert (1 << 4) + (1 << 5) == 48


# ---- :994-1000 test_set_bit
def test_set_bit():
    assert o.set_bit(-96, 5, False) == -64
    assert o.set_bit(-96, 5, True) == -96
    assert o.set_bit(-64, 5, True) == -96
    assert o.set_bit(96, 5, True) == 96
    assert o.set_bit(96, 5, False) == 64


# ---- :1003-1009 test_is_element_sig
def test_is_element_sig():
    assert not o.is_element_sig(-21, 6)
    assert o.is_element_sig(-64, 6)
    assert o.is_element_sig(64, 6)
    assert not o.is_element_sig(55, 6)


# ---- :1012-1024 test_set_bit_doesnt_change_sign (rand-crate stream not reproducible; same property)
def test_set_bit_doesnt_change_sign():
    rng = np.random.default_rng(420)
    for _ in range(420):
        x = int(rng.integers(-2 ** 31 + 1, 2 ** 31 - 1))
        n = int(rng.integers(0, 16))
        bit = bool(rng.integers(0, 2))
        y = o.set_bit(x, n, bit)
        assert (x >= 0) == (y >= 0)
        assert o.is_bit_set(y, n) == bit


# ---- :864-875 simple_test_encode
def test_simple_encode_max_n():
    arr = np.full((1, 16, 16), 32, np.int32)
    _, max_n = o.encode(arr, 2, 2, 10000)
    assert max_n == 5


# ---- :877-888 simple_test_encode_decode
def test_simple_encode_decode():
    arr = np.full((1, 16, 16), 32, np.int32)
    data, max_n = o.encode(arr, 2, 2, 10000)
    assert np.array_equal(o.decode(data, max_n, 1, 16, 16, 2, 2), arr)


# ---- :890-909 simple_test_encode_decode_w_negative
def test_simple_encode_decode_w_negative():
    arr = np.full((1, 16, 16), 32, np.int32)
    arr[:, 0::2, :] *= -1      # sign = (i%2 != 0)*2-1
    data, max_n = o.encode(arr, 2, 2, 10000)
    assert np.array_equal(o.decode(data, max_n, 1, 16, 16, 2, 2), arr)


# ---- :911-927 / :968-985 random lossless round trips (Normal(0,16) truncated to i32)
@pytest.mark.parametrize("c,h,w,reps", [(4, 32, 32, 20), (1, 8, 8, 20)])
def test_encode_decode_many_random(c, h, w, reps):
    rng = np.random.default_rng(42)
    for _ in range(reps):
        arr = rng.normal(0.0, 16.0, (c, h, w)).astype(np.int32)
        data, max_n = o.encode(arr, 2, 2, 10000000)
        assert np.array_equal(o.decode(data, max_n, c, h, w, 2, 2), arr)


# ---- SURVEY.md section 4 known answers
KAT1 = np.array([[[5, -3, 1, 0], [2, -7, 0, 1], [0, 1, -1, 0], [3, 0, 0, -2]]], np.int32)


def test_kat1_bitstream():
    data, max_n, nbits = o.encode_nbits(KAT1, 2, 2, 10 ** 9)
    assert (max_n, nbits) == (2, 49)
    assert data.hex() == "135a166971be00"
    bits = "".join(str(b) for b in np.unpackbits(np.frombuffer(data, np.uint8), bitorder="little")[:49])
    assert bits == "110010000" "1011" "0" "100110" "100010" "01" "01101000" "1110011" "111010"
    assert np.array_equal(o.decode(data, max_n, 1, 4, 4, 2, 2), KAT1)


def test_kat2_all32():
    arr = np.full((1, 16, 16), 32, np.int32)
    data, max_n, nbits = o.encode_nbits(arr, 2, 2, 10000)
    assert (max_n, nbits, len(data)) == (5, 1870, 234)
    bits = np.unpackbits(np.frombuffer(data, np.uint8), bitorder="little")[:nbits]
    assert bits[:590].all() and not bits[590:].any()


def test_kat3_odd_dims_lose_last_row_col():
    arr = np.random.default_rng(0).integers(-100, 100, (1, 17, 17)).astype(np.int32)
    arr[:, 16, :] |= 1
    arr[:, :, 16] |= 1
    data, max_n = o.encode(arr, 2, 2, 10 ** 9)
    rec = o.decode(data, max_n, 1, 17, 17, 2, 2)
    assert np.array_equal(rec[:, :16, :16], arr[:, :16, :16])
    assert not rec[:, 16, :].any() and not rec[:, :, 16].any()


def test_truncation_is_a_prefix_and_max_bits_zero_is_unlimited():
    rng = np.random.default_rng(3)
    arr = rng.normal(0, 40, (3, 24, 40)).astype(np.int32)
    full, max_n, nfull = o.encode_nbits(arr, 4, 6, 10 ** 12)
    fbits = np.unpackbits(np.frombuffer(full, np.uint8), bitorder="little")[:nfull]
    z, _, nz = o.encode_nbits(arr, 4, 6, 0)
    assert nz == nfull and z == full
    for mb in (1, 2, 7, 8, 9, 63, 64, 65, 500, 1001, nfull - 1, nfull, nfull + 5):
        d, mn, nb = o.encode_nbits(arr, 4, 6, mb)
        assert nb == min(mb, nfull) and mn == max_n
        bits = np.unpackbits(np.frombuffer(d, np.uint8), bitorder="little")
        assert np.array_equal(bits[:nb], fbits[:nb]) and not bits[nb:].any()


def test_every_byte_prefix_decodes():
    """make_gif.py:46-61 decodes byte prefixes of one stream (embedded code)."""
    rng = np.random.default_rng(5)
    arr = rng.normal(0, 30, (2, 16, 16)).astype(np.int32)
    data, max_n = o.encode(arr, 2, 2, 10 ** 9)
    prev_err = None
    for cut in range(0, len(data) + 1, 7):
        rec = o.decode(data[:cut], max_n, 2, 16, 16, 2, 2)
        err = np.abs(rec.astype(np.int64) - arr).sum()
        if prev_err is not None:
            assert err <= prev_err * 1.5 + 64
        prev_err = err
    assert np.array_equal(o.decode(data, max_n, 2, 16, 16, 2, 2), arr)


def test_asserts_and_panics():
    arr = np.zeros((1, 8, 8), np.int32)
    with pytest.raises(o.OraclePanic):
        o.encode(arr, 1, 2, 100)          # assert!(ll_h > 1)
    with pytest.raises(o.OraclePanic):
        o.decode(b"\x00", 3, 1, 8, 8, 2, 1)
    with pytest.raises(TypeError):
        o.encode(arr.astype(np.int64), 2, 2, 100)
    big = np.ones((1, 8, 8), np.int32)
    with pytest.raises(o.OraclePanic):
        o.encode(big, 6, 6, 10 ** 6)      # LL-root offspring out of bounds


def test_all_zero_array():
    arr = np.zeros((2, 8, 8), np.int32)
    data, max_n, nbits = o.encode_nbits(arr, 2, 2, 10 ** 6)
    assert max_n == 0
    assert not np.frombuffer(data, np.uint8).any()
    assert not o.decode(data, 0, 2, 8, 8, 2, 2).any()


def test_max_n_is_integer_log2_below_2_19():
    for m in list(range(1, 5000)) + [2 ** k + d for k in range(12, 19) for d in (-1, 0, 1)]:
        assert o.max_n_of(m) == m.bit_length() - 1


# ---- the pyramid / generation model the GPU coder follows == the faithful coder
def test_model_equals_reference_random():
    rng = np.random.default_rng(1)
    checked = 0
    for t in range(400):
        c = int(rng.integers(1, 5)); h = int(rng.integers(4, 48)); w = int(rng.integers(4, 48))
        llh = int(rng.integers(2, max(3, h // 2 + 1))); llw = int(rng.integers(2, max(3, w // 2 + 1)))
        if not o.geom_ok(h, w, llh, llw):
            continue
        x = rng.normal(0, 16 * rng.random() ** 2 * 20 + 0.5, (c, h, w)).astype(np.int32)
        if rng.random() < 0.1:
            x[:] = 0
        mb = int(rng.integers(0, 4000)) if rng.random() < 0.7 else 10 ** 9
        d1, m1, n1 = o.encode_nbits(x, llh, llw, mb)
        d2, m2 = o.model_encode(x, llh, llw, mb)
        assert d1 == d2 and m1 == m2, (t, c, h, w, llh, llw, mb)
        checked += 1
    assert checked > 150


def test_model_equals_reference_reflect_geometry():
    """coefficient-array shapes of mode=reflect (trees straddle bands and zero gaps)"""
    from oracle import wrapper_ref
    from conftest import synth_image
    img = synth_image(3, 96, 80, 7)
    arr, ll_h, ll_w = wrapper_ref.forward_coeffs(img)
    for mb in (0, 5000, 96 * 80):
        d1, m1, _ = o.encode_nbits(arr, ll_h, ll_w, mb)
        d2, m2 = o.model_encode(arr, ll_h, ll_w, mb)
        assert d1 == d2 and m1 == m2


def test_pyramid_matches_recursive_significance():
    rng = np.random.default_rng(9)
    x = rng.normal(0, 50, (2, 21, 26)).astype(np.int32)
    ll_h, ll_w = 3, 4
    dp, lp, dpll, lpll = o.pyramid(x, ll_h, ll_w)

    def dmax(k, i, j):
        off = o.get_offspring(i, j, 21, 26, ll_h, ll_w)
        if off is None:
            return 0
        return max(max(abs(int(x[k, a, b])), dmax(k, a, b)) for a, b in off)

    for k in range(2):
        for i in range(10):
            for j in range(13):
                if i < ll_h and j < ll_w:
                    d = dmax(k, i, j)
                    assert dpll[k, i, j] == (d.bit_length())
                else:
                    assert dp[k, i, j] == dmax(k, i, j).bit_length()
