"""GPU parity: the CUDA SPIHT coder (through the C ABI / the `spiht.spiht` shim)
against the CPU oracle on identical int32 coefficient arrays.  Bit-exact."""
import numpy as np
import pytest

from conftest import synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rs():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from spiht_b200 import spiht
    return spiht


def _bits(data, n=None):
    b = np.unpackbits(np.frombuffer(data, np.uint8), bitorder="little")
    return b if n is None else b[:n]


def assert_stream_equal(got, want, what=""):
    if got == want:
        return
    gb, wb = _bits(got), _bits(want)
    m = min(len(gb), len(wb))
    diff = np.nonzero(gb[:m] != wb[:m])[0]
    first = int(diff[0]) if len(diff) else m
    lo = max(0, first - 16)
    raise AssertionError(
        f"{what}: streams differ: len got {len(got)} want {len(want)} bytes; first differing bit {first}; "
        f"got[{lo}:{first + 16}]={''.join(map(str, gb[lo:first + 16]))} "
        f"want={''.join(map(str, wb[lo:first + 16]))}")


KAT1 = np.array([[[5, -3, 1, 0], [2, -7, 0, 1], [0, 1, -1, 0], [3, 0, 0, -2]]], np.int32)


def test_kat1(rs, oracle):
    data, max_n = rs.encode(KAT1, 2, 2, 10 ** 9)
    assert max_n == 2
    assert data.hex() == "135a166971be00"
    assert np.array_equal(rs.decode(data, max_n, 1, 4, 4, 2, 2), KAT1)


def test_kat2_reference_simple_tests(rs, oracle):
    """src/encoder_decoder.rs:864-909"""
    arr = np.full((1, 16, 16), 32, np.int32)
    data, max_n = rs.encode(arr, 2, 2, 10000)
    assert max_n == 5 and len(data) == 234
    assert_stream_equal(data, oracle.encode(arr, 2, 2, 10000)[0], "all-32")
    assert np.array_equal(rs.decode(data, max_n, 1, 16, 16, 2, 2), arr)
    arr[:, 0::2, :] *= -1
    data, max_n = rs.encode(arr, 2, 2, 10000)
    assert np.array_equal(rs.decode(data, max_n, 1, 16, 16, 2, 2), arr)


@pytest.mark.parametrize("c,h,w,reps", [(4, 32, 32, 20), (1, 8, 8, 20)])
def test_reference_random_roundtrips(rs, oracle, c, h, w, reps):
    """src/encoder_decoder.rs:911-927, 968-985"""
    rng = np.random.default_rng(42)
    for _ in range(reps):
        arr = rng.normal(0.0, 16.0, (c, h, w)).astype(np.int32)
        data, max_n = rs.encode(arr, 2, 2, 10000000)
        want, want_n = oracle.encode(arr, 2, 2, 10000000)
        assert max_n == want_n
        assert_stream_equal(data, want, f"random {c}x{h}x{w}")
        assert np.array_equal(rs.decode(data, max_n, c, h, w, 2, 2), arr)


def test_kat3_odd_dims(rs, oracle):
    arr = np.random.default_rng(0).integers(-100, 100, (1, 17, 17)).astype(np.int32)
    data, max_n = rs.encode(arr, 2, 2, 10 ** 9)
    assert_stream_equal(data, oracle.encode(arr, 2, 2, 10 ** 9)[0], "17x17")
    rec = rs.decode(data, max_n, 1, 17, 17, 2, 2)
    assert np.array_equal(rec, oracle.decode(data, max_n, 1, 17, 17, 2, 2))
    assert np.array_equal(rec[:, :16, :16], arr[:, :16, :16])


def test_random_shapes_truncations(rs, oracle):
    rng = np.random.default_rng(11)
    n_checked = 0
    for t in range(120):
        c = int(rng.integers(1, 5)); h = int(rng.integers(4, 70)); w = int(rng.integers(4, 70))
        llh = int(rng.integers(2, max(3, h // 2 + 1))); llw = int(rng.integers(2, max(3, w // 2 + 1)))
        if not oracle.geom_ok(h, w, llh, llw):
            continue
        x = rng.normal(0, 16 * rng.random() ** 2 * 40 + 0.5, (c, h, w)).astype(np.int32)
        if rng.random() < 0.08:
            x[:] = 0
        mb = int(rng.integers(0, 6000)) if rng.random() < 0.7 else 10 ** 9
        want, want_n = oracle.encode(x, llh, llw, mb)
        got, got_n = rs.encode(x, llh, llw, mb)
        assert got_n == want_n, (t, c, h, w, llh, llw, mb)
        assert_stream_equal(got, want, f"case {t} c={c} h={h} w={w} ll={llh}x{llw} max_bits={mb}")
        rec = rs.decode(got, got_n, c, h, w, llh, llw)
        ref = oracle.decode(got, got_n, c, h, w, llh, llw)
        assert np.array_equal(rec, ref), (t, "decode", c, h, w, llh, llw, mb, int((rec != ref).sum()))
        n_checked += 1
    assert n_checked > 50


def test_truncation_sweep_every_length(rs, oracle):
    """every budget from 1 bit up, incl. mid-record cuts and non-multiples of 8"""
    rng = np.random.default_rng(3)
    arr = rng.normal(0, 40, (3, 12, 20)).astype(np.int32)
    full, max_n, nfull = oracle.encode_nbits(arr, 2, 4, 0)
    budgets = list(range(1, 200)) + list(range(200, nfull + 40, 37)) + [nfull - 1, nfull, nfull + 1]
    for mb in budgets:
        got, got_n = rs.encode(arr, 2, 4, mb)
        want, _ = oracle.encode(arr, 2, 4, mb)
        assert_stream_equal(got, want, f"max_bits={mb}")
        assert got_n == max_n
    got, _ = rs.encode(arr, 2, 4, 0)   # 0 never truncates
    assert_stream_equal(got, full, "max_bits=0")


def test_decode_byte_prefixes_and_pad_bits(rs, oracle):
    """make_gif.py:46-61 decodes byte prefixes; lib.rs:15-21 feeds pad bits to the decoder"""
    rng = np.random.default_rng(5)
    arr = rng.normal(0, 30, (2, 24, 24)).astype(np.int32)
    data, max_n = oracle.encode(arr, 2, 2, 10 ** 9)
    for cut in list(range(0, 40)) + list(range(40, len(data) + 1, 11)) + [len(data)]:
        rec = rs.decode(data[:cut], max_n, 2, 24, 24, 2, 2)
        ref = oracle.decode(data[:cut], max_n, 2, 24, 24, 2, 2)
        assert np.array_equal(rec, ref), (cut, int((rec != ref).sum()))
    # a budget that is not a multiple of 8: the last byte carries pad bits
    for mb in (13, 101, 1003, 2001):
        d, n = oracle.encode(arr, 2, 2, mb)
        assert np.array_equal(rs.decode(d, n, 2, 24, 24, 2, 2), oracle.decode(d, n, 2, 24, 24, 2, 2)), mb


@pytest.mark.parametrize("c,h,w,llh,llw", [(2, 16, 20, 2, 2), (3, 47, 12, 3, 5), (2, 40, 40, 5, 7), (1, 30, 34, 4, 5)])
def test_decode_garbage_streams(rs, oracle, c, h, w, llh, llw):
    """the decoder is a pure function of (bytes, n, geometry): random bytes must decode identically.
    Odd LL sizes duplicate whole subtrees in the reference's lists; writes to those cells must keep list order."""
    rng = np.random.default_rng(17)
    for t in range(10):
        nbytes = int(rng.integers(1, 600))
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        if t % 2:
            data[rng.random(nbytes) < 0.7] = 0      # sparser streams reach deeper planes
        data = data.tobytes()
        n = int(rng.integers(0, 9))
        rec = rs.decode(data, n, c, h, w, llh, llw)
        ref = oracle.decode(data, n, c, h, w, llh, llw)
        assert np.array_equal(rec, ref), (t, nbytes, n, int((rec != ref).sum()))


def test_odd_ll_unaligned_budgets(rs, oracle):
    """odd LL band + budgets that are not multiples of 8: pad bits reach duplicated cells"""
    rng = np.random.default_rng(29)
    for (c, h, w, llh, llw) in [(3, 47, 12, 3, 5), (2, 45, 52, 5, 7), (3, 84, 84, 13, 13)]:
        x = rng.normal(0, 40, (c, h, w)).astype(np.int32)
        for mb in [int(v) for v in rng.integers(50, 9000, 25)]:
            data, n = oracle.encode(x, llh, llw, mb)
            got, got_n = rs.encode(x, llh, llw, mb)
            assert_stream_equal(got, data, f"odd ll {llh}x{llw} mb={mb}")
            rec = rs.decode(data, n, c, h, w, llh, llw)
            ref = oracle.decode(data, n, c, h, w, llh, llw)
            assert np.array_equal(rec, ref), (c, h, w, llh, llw, mb, int((rec != ref).sum()))


def test_tight_odd_ll_geometry_is_accepted(rs, oracle):
    """h == 2 ll_h (and 2 ll_h - 1) with odd ll_h: every one-level decomposition with an odd band.  The lowest
    LL-root offspring row is 2 ll_h - 2 there, so the reference accepts these shapes."""
    rng = np.random.default_rng(31)
    for (c, h, w, llh, llw) in [(1, 22, 22, 11, 11), (2, 21, 22, 11, 11), (3, 106, 106, 53, 53), (1, 21, 25, 11, 13)]:
        x = rng.normal(0, 30, (c, h, w)).astype(np.int32)
        for mb in (10 ** 9, 1500, 333):
            want, want_n = oracle.encode(x, llh, llw, mb)
            got, got_n = rs.encode(x, llh, llw, mb)
            assert got_n == want_n
            assert_stream_equal(got, want, f"tight odd ll {(c, h, w, llh, llw)} mb={mb}")
            assert np.array_equal(rs.decode(got, got_n, c, h, w, llh, llw),
                                  oracle.decode(want, want_n, c, h, w, llh, llw))
    # 18x18 and 17x17 images at default settings: level 1, ll = 11 (the shapes ADVICE names)
    import spiht_b200 as spiht
    from oracle import wrapper_ref
    for n in (18, 17, 101, 102):
        img = synth_image(3, n, n, n)
        lvl = None if n < 30 else 1
        enc = spiht.encode_image(img, spiht.SpihtSettings(), level=lvl)
        ref = wrapper_ref.encode_image(img, level=lvl)
        assert enc.max_n == ref["max_n"] and enc.encoded_bytes == ref["encoded_bytes"], n
        rec = spiht.decode_image(enc, spiht.SpihtSettings())
        assert np.abs(rec - wrapper_ref.decode_image(ref)).max() < 1e-9


@pytest.mark.parametrize("shape,wavelet,mode", [((3, 96, 80), "bior2.2", "reflect"),
                                                ((3, 256, 384), "bior2.2", "reflect"),
                                                ((1, 130, 70), "bior4.4", "symmetric"),
                                                ((3, 128, 128), "bior2.2", "periodization")])
def test_wavelet_geometry_arrays(rs, oracle, shape, wavelet, mode):
    """coefficient arrays with the reference's real geometry (odd ll included), several budgets"""
    from oracle import wrapper_ref
    c, h, w = shape
    img = synth_image(c, h, w, 7)
    arr, ll_h, ll_w = wrapper_ref.forward_coeffs(img, wavelet=wavelet, mode=mode)
    for bpp in (0.075, 0.5, 1.0, None):
        mb = 10 ** 12 if bpp is None else int(h * w * bpp)
        want, want_n = oracle.model_encode(arr, ll_h, ll_w, mb)
        got, got_n = rs.encode(arr, ll_h, ll_w, mb)
        assert got_n == want_n
        assert_stream_equal(got, want, f"{shape} {wavelet} {mode} bpp={bpp}")
        rec = rs.decode(got, got_n, c, arr.shape[1], arr.shape[2], ll_h, ll_w)
        assert np.array_equal(rec, oracle.decode(got, got_n, c, arr.shape[1], arr.shape[2], ll_h, ll_w))


def test_batched_device_api_per_image_budgets(oracle):
    import torch
    from spiht_b200 import batch
    rng = np.random.default_rng(23)
    B, c, h, w, llh, llw = 9, 3, 40, 56, 4, 6
    x = rng.normal(0, 60, (B, c, h, w)).astype(np.int32)
    x[4] = 0
    budgets = np.array([0, 1, 77, 800, 4000, 12345, 10 ** 9, 31, 2048], dtype=np.int64)
    stride = 8 * 4096
    streams, nbits, max_n, status = batch.encode_coeffs(torch.from_numpy(x).cuda(), llh, llw,
                                                        torch.from_numpy(budgets).cuda(), out_stride=stride)
    streams, nbits, max_n = streams.cpu().numpy(), nbits.cpu().numpy(), max_n.cpu().numpy()
    assert not status.cpu().numpy().any()
    nbytes = (nbits + 7) // 8
    for b in range(B):
        want, want_n, want_bits = oracle.encode_nbits(x[b], llh, llw, int(budgets[b]))
        assert (int(nbits[b]), int(max_n[b])) == (want_bits, want_n), b
        assert_stream_equal(streams[b, :nbytes[b]].tobytes(), want, f"image {b}")
    # batched decode of the same rows (rows are NOT zero padded past their length on purpose)
    rec = batch.decode_coeffs(torch.from_numpy(streams).cuda(), torch.from_numpy(nbytes), torch.from_numpy(max_n),
                              c, h, w, llh, llw).cpu().numpy()
    for b in range(B):
        ref = oracle.decode(streams[b, :nbytes[b]].tobytes(), int(max_n[b]), c, h, w, llh, llw)
        assert np.array_equal(rec[b], ref), b


def test_capacity_flag(oracle):
    import torch
    from spiht_b200 import batch
    x = np.random.default_rng(1).normal(0, 60, (2, 1, 32, 32)).astype(np.int32)
    streams, nbits, max_n, status = batch.encode_coeffs(torch.from_numpy(x).cuda(), 2, 2, 0, out_stride=64)
    assert status.cpu().numpy().all() and (nbits.cpu().numpy() == 512).all()
    for b in range(2):
        want, _ = oracle.encode(x[b], 2, 2, 512)
        assert_stream_equal(streams[b].cpu().numpy().tobytes(), want, "cap-limited")


def test_errors(rs):
    from spiht_b200 import _lib
    arr = np.zeros((1, 8, 8), np.int32)
    with pytest.raises(_lib.SpihtB200Error):      # reference: assert!(ll_h > 1) -> PanicException
        rs.encode(arr, 1, 2, 100)
    with pytest.raises(_lib.SpihtB200Error):      # reference: index out of bounds -> PanicException
        rs.encode(np.ones((1, 8, 8), np.int32), 6, 6, 100)
    with pytest.raises(TypeError):                # reference: PyReadonlyArray3<i32>
        rs.encode(arr.astype(np.int64), 2, 2, 100)
    with pytest.raises(TypeError):
        rs.encode(arr[0], 2, 2, 100)
    with pytest.raises(_lib.SpihtB200Error):
        rs.decode(b"\x01", 3, 1, 8, 8, 2, 1)


def test_full_size_config2_against_model_oracle(oracle):
    """BASELINE config 2 shape (3x1024x1024, bior2.2 reflect -> 3x1053x1053, 0.5 bpp):
    two images bit-exact against the oracle, the rest of the batch through a
    size-independent property of the decoded arrays."""
    import torch
    from spiht_b200 import batch
    rng = np.random.default_rng(2)
    B, c, H, W, ll = 6, 3, 1053, 1053, 12
    # heavy-tailed synthetic coefficients with a decaying scale towards fine bands
    yy, xx = np.mgrid[0:H, 0:W]
    scale = 4000.0 / (1.0 + np.maximum(yy, xx)) ** 1.2 + 0.4
    x = (rng.standard_t(3, (B, c, H, W)) * scale).astype(np.int32)
    mb = 1024 * 1024 // 2
    xd = torch.from_numpy(x).cuda()
    streams, nbits, max_n, status = batch.encode_coeffs(xd, ll, ll, mb)
    s_h, nb_h, mn_h = streams.cpu().numpy(), nbits.cpu().numpy(), max_n.cpu().numpy()
    assert (nb_h == mb).all() and not status.cpu().numpy().any()
    for b in range(2):
        want, want_n = oracle.model_encode(x[b], ll, ll, mb)
        assert int(mn_h[b]) == want_n
        assert_stream_equal(s_h[b, :mb // 8].tobytes(), want, f"1053^2 image {b}")
    nbytes = torch.from_numpy((nb_h + 7) // 8)
    rec = batch.decode_coeffs(streams, nbytes, max_n, c, H, W, ll, ll)
    rec_h = rec.cpu().numpy()
    for b in range(B):
        ref = oracle.decode(s_h[b, :mb // 8].tobytes(), int(mn_h[b]), c, H, W, ll, ll)
        bad = np.argwhere(rec_h[b] != ref)
        assert len(bad) == 0, (b, len(bad), bad[:5].tolist(), [(int(rec_h[b][tuple(t)]), int(ref[tuple(t)])) for t in bad[:5]])
    # size-independent property over the whole batch: every decoded coefficient has the
    # sign and the top bit-plane of the original (significance planes are exact)
    nz = rec != 0
    assert int(nz.sum()) > 0
    assert torch.equal(torch.sign(rec[nz]), torch.sign(xd[nz]))
    top = lambda t: torch.floor(torch.log2(t.abs().double() + 0.5))   # +0.5: exact powers of two stay exact
    assert torch.equal(top(rec[nz]), top(xd[nz]))


@pytest.mark.parametrize("cl", [2, 4, 8, 16])
def test_cluster_coder_matches_oracle(rs, oracle, monkeypatch, cl):
    """the cluster coder (csrc/spiht_enc_cl.cu: one thread-block cluster per image, chunk totals exchanged through
    distributed shared memory) gives the oracle's stream at every budget -- lists several super-chunks long, budgets
    that end in every pass, word-unaligned chunk boundaries, batches with more images than clusters"""
    import torch
    from spiht_b200 import batch
    monkeypatch.setenv("SPIHTB_ENC_CLUSTER", str(cl))
    rng = np.random.default_rng(100 + cl)
    # a large flat-spectrum array: long lists (hundreds of thousands of entries) at modest size
    c, h, w, llh, llw = 3, 264, 392, 8, 12
    x = (rng.laplace(0, 6, (c, h, w)) * (1 + 40 * (rng.random((c, h, w)) < 0.01))).astype(np.int32)
    for mb in [10 ** 9, 1, 33, 4097, 65536, 300001, 1234567]:
        want, want_n = oracle.model_encode(x, llh, llw, mb)
        got, got_n = rs.encode(x, llh, llw, mb)
        assert got_n == want_n
        assert_stream_equal(got, want, f"cluster {cl} mb={mb}")
    # batch: per-image budgets, more images than clusters can be resident
    B = 11
    xb = (rng.laplace(0, 5, (B, 2, 72, 88))).astype(np.int32)
    xb[3] = 0
    budgets = np.array([0, 5, 77, 800, 4000, 12345, 10 ** 9, 31, 2048, 9999, 64], dtype=np.int64)
    streams, nbits, max_n, status = batch.encode_coeffs(torch.from_numpy(xb).cuda(), 4, 6, torch.from_numpy(budgets).cuda(),
                                                        out_stride=8 * 8192)
    streams, nbits, max_n = streams.cpu().numpy(), nbits.cpu().numpy(), max_n.cpu().numpy()
    for b in range(B):
        want, want_n, want_bits = oracle.encode_nbits(xb[b], 4, 6, int(budgets[b]))
        assert (int(nbits[b]), int(max_n[b])) == (want_bits, want_n), b
        assert_stream_equal(streams[b, :(want_bits + 7) // 8].tobytes(), want, f"cluster {cl} image {b}")


def test_decoder_pipelined_rounds_and_walk_variants(rs, oracle, monkeypatch):
    """the decoder's LIS rounds run pipelined (warps apply round r-1 on a named barrier while one thread walks round
    r) with the branch-free walk; SPIHTB_DEC_PIPE=0 / SPIHTB_WALK=0 select the serial order and the round-1 loop.
    All combinations must decode every prefix to the oracle's array: lists several rounds (2048 entries) long per
    generation, prefixes that end inside a walk, inside an apply, inside the first and the last round of a
    generation, one-byte and empty streams, a batch with per-image lengths"""
    import torch
    from spiht_b200 import batch
    rng = np.random.default_rng(77)
    c, h, w, llh, llw = 3, 264, 392, 8, 12            # even LL: the pipelined path
    x = (rng.laplace(0, 9, (c, h, w)) * (1 + 30 * (rng.random((c, h, w)) < 0.02))).astype(np.int32)
    data, n = rs.encode(x, llh, llw, 0)               # untruncated: ~1.5 Mbit, generations of tens of thousands of entries
    assert len(data) > 100000
    lengths = [0, 1, 2, 7, 100, 4099, 20000, 65537, 100003, len(data) // 2, len(data) - 1, len(data)]
    lengths += [int(v) for v in rng.integers(1000, len(data), 12)]
    want = {L: oracle.decode(data[:L], n, c, h, w, llh, llw) for L in lengths}
    for pipe, walk in [("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")]:
        monkeypatch.setenv("SPIHTB_DEC_PIPE", pipe)
        monkeypatch.setenv("SPIHTB_WALK", walk)
        for L in lengths:
            got = rs.decode(data[:L], n, c, h, w, llh, llw)
            assert np.array_equal(got, want[L]), (pipe, walk, L, int((got != want[L]).sum()))
    monkeypatch.delenv("SPIHTB_DEC_PIPE")
    monkeypatch.delenv("SPIHTB_WALK")
    # a batch: one row per prefix, more rows than CTAs fit on the device at once would need > 296 rows; 24 rows with
    # their own lengths check the per-image state (pending round, counters handed back through shared memory)
    B = len(lengths)
    stride = (len(data) + 15) // 8 * 8
    rows = np.zeros((B, stride), np.uint8)
    for r, L in enumerate(lengths):
        rows[r, :L] = np.frombuffer(data[:L], np.uint8)
    rec = batch.decode_coeffs(torch.from_numpy(rows).cuda(), torch.tensor(lengths, dtype=torch.int64),
                              torch.full((B,), n, dtype=torch.int32), c, h, w, llh, llw).cpu().numpy()
    for r, L in enumerate(lengths):
        assert np.array_equal(rec[r], want[L]), (r, L)
    # repeated runs are bit-identical (a race between the walker and the applying warps would not be)
    rec2 = batch.decode_coeffs(torch.from_numpy(rows).cuda(), torch.tensor(lengths, dtype=torch.int64),
                               torch.full((B,), n, dtype=torch.int32), c, h, w, llh, llw).cpu().numpy()
    assert np.array_equal(rec, rec2)
