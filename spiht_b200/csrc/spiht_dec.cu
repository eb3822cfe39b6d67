// SPIHT decoder (replaces src/encoder_decoder.rs:307-454), one CTA per stream.
//
// The decoder's control flow depends on the bits it reads, so record
// boundaries cannot be found by a scan alone.  Each list pass is split in two:
//   chain   : one thread walks the pass's bit region, skipping runs of zero
//             bits 64 at a time (a zero is "entry stays"), and records an
//             event (entry index, bit position) at every set bit that starts a
//             record -- a newly significant pixel, a fired A set (whose 4..8
//             child bits it skips) or a fired B set.  The chain touches only
//             the bitstream (read-only cache) and two shared-memory bitmasks.
//   apply   : all threads process the events in parallel (child bits, value
//             writes, list appends through a CTA-wide scan) and compact the
//             retained entries in place.
// The refinement pass is a plain parallel map (entry e reads bit p0 + e).
// Values are written straight into the coefficient array, as the reference
// does, so pad bits and duplicated coordinates behave identically.  Decoding
// stops at the first bit position >= 8 * nbytes (pop_bit!, :314-325): an
// action is applied only if every bit it needs lies below that.
//
// Odd LL sizes: the reference's LL-root offspring rule (:50-62) then gives some
// cells two parents, so whole subtrees sit in the lists twice and two list
// entries write the same cell.  Those writes must land in list order (pad bits
// or foreign streams can make the two entries disagree), so writes to cells of
// duplicated subtrees are queued in list order (ordered compaction through the
// CTA scan) and applied by one thread after each parallel step.  Even LL sizes
// (all BASELINE configs) never take that path.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace spihtb {

constexpr int DEC_NT = 256;
constexpr int DEC_CH = 8192;  // list entries per chain round
constexpr int DEC_EV = 1024;  // event buffer
constexpr int DEC_SLACK = 8 * DEC_NT + 64;
constexpr int DEC_DQ = 4 * DEC_NT;  // ordered write queue (odd LL sizes only)

struct DecK {
    const uint32_t *in;
    uint64_t in_stride_words;
    const uint64_t *nbytes;
    const int32_t *n;
    int B, C, H, W, ll_h, ll_w;
    KeyFmt kf;
    int32_t *out;
    uint32_t *lip, *lsp, *lis;  // lis: 3 buffers per slot
    size_t pix_cap, lis_cap;
    unsigned int *counter;
};

struct BitRow {
    const uint32_t *row;
    uint64_t nwords;
    uint64_t nbits;
    __device__ __forceinline__ uint32_t word(uint64_t idx) const { return idx < nwords ? __ldg(row + idx) : 0u; }
    __device__ __forceinline__ uint64_t get64(uint64_t p) const
    {
        const uint64_t idx = p >> 5;
        const int sh = (int)(p & 31);
        const uint32_t w0 = word(idx), w1 = word(idx + 1), w2 = word(idx + 2);
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        return (uint64_t)lo | ((uint64_t)hi << 32);
    }
    __device__ __forceinline__ uint32_t get32(uint64_t p) const
    {
        const uint64_t idx = p >> 5;
        return __funnelshift_r(word(idx), word(idx + 1), (int)(p & 31));
    }
    __device__ __forceinline__ uint32_t bit(uint64_t p) const { return (word(p >> 5) >> (p & 31)) & 1u; }
};

// Is (y,x) inside a subtree that the reference's lists hold twice?  Cells with
// two LL-root parents: row ll_h (odd ll_h) x cols [ll_w, 2 ll_w) and col ll_w
// (odd ll_w) x rows [ll_h, 2 ll_h); their descendants follow the dyadic rule.
__device__ __forceinline__ bool in_dup_subtree(uint32_t y, uint32_t x, uint32_t ll_h, uint32_t ll_w)
{
    for (int t = 0; t < 32; ++t) {
        const uint32_t yy = y >> t, xx = x >> t;
        if (yy < ll_h && xx < ll_w) return false;
        if ((ll_h & 1u) && yy == ll_h && xx >= ll_w && xx < 2 * ll_w) return true;
        if ((ll_w & 1u) && xx == ll_w && yy >= ll_h && yy < 2 * ll_h) return true;
    }
    return false;
}

struct ChainState {
    uint64_t p;      // next bit position
    uint32_t e;      // next entry (relative to the round)
    uint32_t nev;    // events recorded in this sub-round
    uint32_t ended;  // the stream is exhausted
};

// Thread 0 only.  Entries [st.e, cnt) of the round; stops when the event
// buffer is full.  lis_mode: record kind taken from tmask (1 = A, 0 = B);
// otherwise LIP records ("0" | "1 sign").
__device__ void run_chain(const BitRow &br, bool lis_mode, const uint32_t *tmask, uint32_t *fmask, uint2 *ev,
                          uint32_t cnt, uint64_t p_base, ChainState &st)
{
    uint64_t p = st.p;
    uint32_t e = st.e, nev = 0;
    while (e < cnt && nev < DEC_EV) {
        if (p >= br.nbits) break;
        const uint64_t w = br.get64(p);
        uint64_t lim = cnt - e;
        if (lim > 64) lim = 64;
        const uint64_t rem = br.nbits - p;
        if (rem < lim) lim = rem;  // bits past the end are not data
        const uint64_t wl = lim < 64 ? (w & ((1ull << lim) - 1ull)) : w;
        if (wl == 0) {
            e += (uint32_t)lim;
            p += lim;
            continue;
        }
        const int z = __ffsll((long long)wl) - 1;
        e += z;
        p += z;
        fmask[e >> 5] |= 1u << (e & 31);
        uint32_t len;
        if (!lis_mode) {
            len = 1;
        } else if ((tmask[e >> 5] >> (e & 31)) & 1u) {
            uint32_t cb = (z + 9 <= 64) ? (uint32_t)(w >> (z + 1)) : br.get32(p + 1);
            len = 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (cb & 1u) {
                    cb >>= 2;
                    ++len;
                } else {
                    cb >>= 1;
                }
            }
        } else {
            len = 0;
        }
        ev[nev++] = make_uint2(e, (uint32_t)(p - p_base));
        p += 1 + len;
        e += 1;
    }
    st.p = p;
    st.e = e;
    st.nev = nev;
    st.ended = p >= br.nbits ? 1u : 0u;
}

// set_bit (encoder_decoder.rs:14-29): set / clear bit n of the magnitude, keep the sign
__device__ __forceinline__ void refine_cell(int32_t *cell, int n, uint32_t bit)
{
    const int32_t x = *cell;
    const uint32_t m = 1u << n;
    uint32_t mag = absu(x);
    mag = bit ? (mag | m) : (mag & ~m);
    *cell = x >= 0 ? (int32_t)mag : -(int32_t)mag;
}

// One thread: apply queued writes in list order.
__device__ void apply_queue(const uint2 *dq, uint32_t cnt, const KeyFmt &kf, int32_t *rec, uint32_t H, uint32_t W,
                            int n)
{
    for (uint32_t q = 0; q < cnt; ++q) {
        const uint2 op = dq[q];
        uint32_t k, i, j;
        key_unpack(kf, op.x, k, i, j);
        int32_t *cell = rec + ((size_t)k * H + i) * W + j;
        if (op.x >> 31)
            refine_cell(cell, n, op.y);
        else
            *cell = (int32_t)op.y;
    }
}

__global__ void __launch_bounds__(DEC_NT) spiht_decode_kernel(const DecK p)
{
    __shared__ uint32_t s_tmask[DEC_CH / 32];
    __shared__ uint32_t s_fmask[DEC_CH / 32];
    __shared__ uint2 s_ev[DEC_EV];
    __shared__ uint64_t s_wtot[DEC_NT / 32];
    __shared__ ChainState s_st;
    __shared__ int s_img;
    __shared__ uint2 s_dq[DEC_DQ];  // {cell key | refine flag << 31, value or bit}

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const KeyFmt kf = p.kf;
    const uint32_t H = p.H, W = p.W, ll_h = p.ll_h, ll_w = p.ll_w, C = p.C;
    uint32_t *lip = p.lip + (size_t)blockIdx.x * p.pix_cap;
    uint32_t *lsp = p.lsp + (size_t)blockIdx.x * p.pix_cap;
    uint32_t *R = p.lis + (size_t)blockIdx.x * 3 * p.lis_cap;
    uint32_t *G0 = R + p.lis_cap;
    uint32_t *G1 = G0 + p.lis_cap;
    const bool has_dups = ((ll_h | ll_w) & 1u) != 0;

    for (;;) {
        if (tid == 0) s_img = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int b = s_img;
        __syncthreads();
        if (b >= p.B) break;

        BitRow br;
        br.row = p.in + (size_t)b * p.in_stride_words;
        br.nwords = p.in_stride_words;
        br.nbits = p.nbytes[b] * 8ull;
        if (br.nbits > br.nwords * 32ull) br.nbits = br.nwords * 32ull;
        int32_t *rec = p.out + (size_t)b * C * H * W;
        int n = p.n[b];
        n = n < 0 ? 0 : (n > 31 ? 31 : n);

        // ---- list initialisation (encoder_decoder.rs:329-348)
        const uint32_t T0 = ll_h * ll_w * C;
        uint32_t lip_len = T0, lsp_len = 0, r_len = 0;
        for (uint32_t base = 0; base < T0; base += DEC_NT) {
            const uint32_t t = base + tid;
            const bool valid = t < T0;
            uint32_t key = 0;
            bool root = false;
            if (valid) {
                const uint32_t k = t % C, ij = t / C;
                const uint32_t j = ij % ll_w, i = ij / ll_w;
                key = key_pack(kf, k, i, j);
                lip[t] = key;
                root = ((i | j) & 1u) != 0;
            }
            uint64_t tot;
            const uint64_t ex = block_exscan<DEC_NT>(root ? 1ull : 0ull, s_wtot, tot);
            if (root) R[r_len + (uint32_t)ex] = 0x80000000u | key;
            r_len += (uint32_t)tot;
        }
        __syncthreads();

        uint64_t pos = 0;  // uniform: next unread bit
        bool ended = br.nbits == 0;
        for (; !ended; --n) {
            const int32_t basev = n == 0 ? 1 : (int32_t)((1u << (n - 1)) + (1u << n));
            const uint32_t lsp_len0 = lsp_len;

            // ================= LIP pass (encoder_decoder.rs:355-377)
            {
                uint32_t keep = 0;
                for (uint32_t ebase = 0; ebase < lip_len && !ended; ebase += DEC_CH) {
                    const uint32_t cnt = min((uint32_t)DEC_CH, lip_len - ebase);
                    for (uint32_t i = tid; i < (cnt + 31) / 32; i += DEC_NT) s_fmask[i] = 0;
                    if (tid == 0) {
                        s_st.p = pos;
                        s_st.e = 0;
                        s_st.ended = 0;
                    }
                    __syncthreads();
                    uint32_t eprev = 0;
                    for (;;) {
                        const uint64_t p_base = pos;
                        if (tid == 0) run_chain(br, false, nullptr, s_fmask, s_ev, cnt, p_base, s_st);
                        __syncthreads();
                        const uint32_t nev = s_st.nev, ecur = s_st.e;
                        const bool end_now = s_st.ended != 0;
                        pos = s_st.p;
                        // events: newly significant pixels, in list order
                        for (uint32_t rb = 0; rb < nev; rb += DEC_NT) {
                            const uint32_t r = rb + tid;
                            bool defer = false;
                            uint32_t key = 0;
                            int32_t val = 0;
                            if (r < nev) {
                                const uint2 ev = s_ev[r];
                                const uint64_t ps = p_base + ev.y + 1;  // sign bit
                                if (ps < br.nbits) {
                                    key = lip[ebase + ev.x];
                                    uint32_t k, i, j;
                                    key_unpack(kf, key, k, i, j);
                                    val = br.bit(ps) ? basev : -basev;
                                    defer = has_dups && in_dup_subtree(i, j, ll_h, ll_w);
                                    if (!defer) rec[((size_t)k * H + i) * W + j] = val;
                                    lsp[lsp_len + r] = key;
                                }
                            }
                            if (has_dups) {
                                uint64_t tot;
                                const uint64_t ex = block_exscan<DEC_NT>(defer ? 1ull : 0ull, s_wtot, tot);
                                if (defer) s_dq[(uint32_t)ex] = make_uint2(key, (uint32_t)val);
                                __syncthreads();
                                if (tid == 0) apply_queue(s_dq, (uint32_t)tot, kf, rec, H, W, n);
                                __syncthreads();
                            }
                        }
                        lsp_len += nev;
                        // retained entries [eprev, ecur): in-place stable compaction
                        for (uint32_t cb = eprev; cb < ecur; cb += DEC_NT) {
                            const uint32_t e = cb + tid;
                            const bool valid = e < ecur;
                            const bool keepit = valid && !((s_fmask[e >> 5] >> (e & 31)) & 1u);
                            const uint32_t key = keepit ? lip[ebase + e] : 0u;
                            uint64_t tot;
                            const uint64_t ex = block_exscan<DEC_NT>(keepit ? 1ull : 0ull, s_wtot, tot);
                            if (keepit) lip[keep + (uint32_t)ex] = key;
                            keep += (uint32_t)tot;
                        }
                        eprev = ecur;
                        __syncthreads();
                        if (end_now) ended = true;
                        if (ended || ecur >= cnt) break;
                    }
                }
                if (ended) break;
                lip_len = keep;
            }

            // ================= LIS pass (encoder_decoder.rs:379-436), generation by generation
            {
                uint32_t *cur = R, *nxt = G0;
                uint32_t cur_len = r_len, rkeep = 0;
                int gen = 0;
                while (cur_len > 0 && !ended) {
                    uint32_t nxt_len = 0;
                    for (uint32_t ebase = 0; ebase < cur_len && !ended; ebase += DEC_CH) {
                        const uint32_t cnt = min((uint32_t)DEC_CH, cur_len - ebase);
                        // set-type mask of this round's entries
                        // (bit set = A set that has offspring, i.e. a fired record carries child bits)
                        for (uint32_t w = wid; w < (cnt + 31) / 32; w += DEC_NT / 32) {
                            const uint32_t e = w * 32 + lane;
                            const uint32_t key = e < cnt ? cur[ebase + e] : 0u;
                            bool a_with_children = false;
                            if (key >> 31) {
                                uint32_t k, i, j, ci, cj;
                                key_unpack(kf, key, k, i, j);
                                a_with_children = offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj);
                            }
                            const uint32_t m = __ballot_sync(0xffffffffu, a_with_children);
                            if (lane == 0) {
                                s_tmask[w] = m;
                                s_fmask[w] = 0;
                            }
                        }
                        if (tid == 0) {
                            s_st.p = pos;
                            s_st.e = 0;
                            s_st.ended = 0;
                        }
                        __syncthreads();
                        uint32_t eprev = 0;
                        for (;;) {
                            const uint64_t p_base = pos;
                            if (tid == 0) run_chain(br, true, s_tmask, s_fmask, s_ev, cnt, p_base, s_st);
                            __syncthreads();
                            const uint32_t nev = s_st.nev, ecur = s_st.e;
                            const bool end_now = s_st.ended != 0;
                            pos = s_st.p;
                            // events: fired sets, in list order
                            for (uint32_t rb = 0; rb < nev; rb += DEC_NT) {
                                const uint32_t r = rb + tid;
                                const bool valid = r < nev;
                                uint32_t key = 0, k = 0, i = 0, j = 0, ci = 0, cj = 0;
                                uint32_t nlsp = 0, nlip = 0, nnext = 0, sigmask = 0, sgnmask = 0, nread = 0;
                                uint32_t ndef = 0, defmask = 0;
                                bool isA = false, has = false;
                                if (valid) {
                                    const uint2 ev = s_ev[r];
                                    key = cur[ebase + ev.x];
                                    isA = (key >> 31) != 0;
                                    key_unpack(kf, key, k, i, j);
                                    has = offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj);
                                    if (isA) {
                                        uint64_t q = p_base + ev.y + 1;
                                        uint32_t cbits = br.get32(q);
                                        bool cut = false;
                                        if (has) {
#pragma unroll
                                            for (int c4 = 0; c4 < 4; ++c4) {
                                                if (cut) break;
                                                if (q >= br.nbits) { cut = true; break; }
                                                const uint32_t sg = cbits & 1u;
                                                cbits >>= 1;
                                                ++q;
                                                if (sg) {
                                                    if (q >= br.nbits) { cut = true; break; }
                                                    sigmask |= 1u << c4;
                                                    sgnmask |= (cbits & 1u) << c4;
                                                    cbits >>= 1;
                                                    ++q;
                                                    ++nlsp;
                                                } else {
                                                    ++nlip;
                                                }
                                                ++nread;
                                            }
                                        }
                                        if (!cut && has_desc_past_offspring(i, j, H, W)) nnext = 1;
                                        if (has_dups && nlsp) {
                                            for (uint32_t c4 = 0; c4 < nread; ++c4)
                                                if ((sigmask & (1u << c4)) &&
                                                    in_dup_subtree(ci + (c4 >> 1), cj + (c4 & 1), ll_h, ll_w)) {
                                                    defmask |= 1u << c4;
                                                    ++ndef;
                                                }
                                        }
                                    } else {
                                        nnext = has ? 4 : 0;
                                    }
                                }
                                const uint64_t pack = (uint64_t)nlsp | ((uint64_t)nlip << 16) | ((uint64_t)nnext << 32) |
                                                      ((uint64_t)ndef << 48);
                                uint64_t tot;
                                const uint64_t ex = block_exscan<DEC_NT>(pack, s_wtot, tot);
                                if (valid) {
                                    uint32_t os = lsp_len + (uint32_t)(ex & 0xffff);
                                    uint32_t oi = lip_len + (uint32_t)((ex >> 16) & 0xffff);
                                    const uint32_t on = nxt_len + (uint32_t)((ex >> 32) & 0xffff);
                                    uint32_t od = (uint32_t)(ex >> 48);
                                    if (isA) {
                                        for (uint32_t c4 = 0; c4 < nread; ++c4) {
                                            const uint32_t y = ci + (c4 >> 1), x = cj + (c4 & 1);
                                            const uint32_t ck = key_pack(kf, k, y, x);
                                            if (sigmask & (1u << c4)) {
                                                const int32_t val = (sgnmask & (1u << c4)) ? basev : -basev;
                                                if (defmask & (1u << c4))
                                                    s_dq[od++] = make_uint2(ck, (uint32_t)val);
                                                else
                                                    rec[((size_t)k * H + y) * W + x] = val;
                                                lsp[os++] = ck;
                                            } else {
                                                lip[oi++] = ck;
                                            }
                                        }
                                        if (nnext) nxt[on] = key & 0x7fffffffu;
                                    } else if (nnext) {
#pragma unroll
                                        for (int c4 = 0; c4 < 4; ++c4)
                                            nxt[on + c4] = 0x80000000u | key_pack(kf, k, ci + (c4 >> 1), cj + (c4 & 1));
                                    }
                                }
                                lsp_len += (uint32_t)(tot & 0xffff);
                                lip_len += (uint32_t)((tot >> 16) & 0xffff);
                                nxt_len += (uint32_t)((tot >> 32) & 0xffff);
                                if (has_dups) {
                                    __syncthreads();
                                    if (tid == 0) apply_queue(s_dq, (uint32_t)(tot >> 48), kf, rec, H, W, n);
                                    __syncthreads();
                                }
                            }
                            // retained sets [eprev, ecur)
                            for (uint32_t cb = eprev; cb < ecur; cb += DEC_NT) {
                                const uint32_t e = cb + tid;
                                const bool valid = e < ecur;
                                const bool keepit = valid && !((s_fmask[e >> 5] >> (e & 31)) & 1u);
                                const uint32_t key = keepit ? cur[ebase + e] : 0u;
                                uint64_t tot;
                                const uint64_t ex = block_exscan<DEC_NT>(keepit ? 1ull : 0ull, s_wtot, tot);
                                if (keepit) R[rkeep + (uint32_t)ex] = key;
                                rkeep += (uint32_t)tot;
                            }
                            eprev = ecur;
                            __syncthreads();
                            if (end_now) ended = true;
                            if (ended || ecur >= cnt) break;
                        }
                    }
                    uint32_t *old = cur;
                    cur = nxt;
                    nxt = gen == 0 ? G1 : old;
                    cur_len = nxt_len;
                    ++gen;
                }
                r_len = rkeep;
            }
            if (ended) break;

            // ================= refinement (encoder_decoder.rs:439-444)
            if (!has_dups) {
                for (uint32_t e = tid; e < lsp_len0; e += DEC_NT) {
                    const uint64_t q = pos + e;
                    if (q < br.nbits) {
                        uint32_t k, i, j;
                        key_unpack(kf, lsp[e], k, i, j);
                        refine_cell(rec + ((size_t)k * H + i) * W + j, n, br.bit(q));
                    }
                }
            } else {
                for (uint32_t eb = 0; eb < lsp_len0; eb += DEC_NT) {
                    const uint32_t e = eb + tid;
                    const uint64_t q = pos + e;
                    bool defer = false;
                    uint32_t key = 0, bit = 0;
                    if (e < lsp_len0 && q < br.nbits) {
                        key = lsp[e];
                        bit = br.bit(q);
                        uint32_t k, i, j;
                        key_unpack(kf, key, k, i, j);
                        defer = in_dup_subtree(i, j, ll_h, ll_w);
                        if (!defer) refine_cell(rec + ((size_t)k * H + i) * W + j, n, bit);
                    }
                    uint64_t tot;
                    const uint64_t ex = block_exscan<DEC_NT>(defer ? 1ull : 0ull, s_wtot, tot);
                    if (defer) s_dq[(uint32_t)ex] = make_uint2(key | 0x80000000u, bit);
                    __syncthreads();
                    if (tid == 0) apply_queue(s_dq, (uint32_t)tot, kf, rec, H, W, n);
                    __syncthreads();
                }
            }
            pos += lsp_len0;
            if (pos >= br.nbits) ended = true;
            __syncthreads();
            if (n == 0) break;
        }
        __syncthreads();
    }
}

int launch_decode(spihtb_ctx *ctx, const DecArgs &a)
{
    DecK k;
    if (!make_keyfmt(a.C, a.H, a.W, &k.kf)) {
        set_error("shape c=%d h=%d w=%d does not fit a 31-bit packed list entry", a.C, a.H, a.W);
        return SPIHTB_ESHAPE;
    }
    if ((a.in_stride & 7) != 0 || a.in_stride == 0) {
        set_error("in_stride must be a positive multiple of 8 bytes");
        return SPIHTB_EINVAL;
    }
    k.in = reinterpret_cast<const uint32_t *>(a.in);
    k.in_stride_words = a.in_stride / 4;
    k.nbytes = a.nbytes;
    k.n = a.n;
    k.B = a.B; k.C = a.C; k.H = a.H; k.W = a.W; k.ll_h = a.ll_h; k.ll_w = a.ll_w;
    k.out = a.out;

    int occ = 1;
    SPIHTB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spiht_decode_kernel, DEC_NT, 0));
    if (occ < 1) occ = 1;
    const int slots = std::min(a.B, ctx->sm_count * occ);

    const uint64_t T0 = (uint64_t)a.ll_h * a.ll_w * a.C;
    const uint64_t chw = (uint64_t)a.C * a.H * a.W;
    const uint64_t budget = a.in_stride * 8;
    const uint64_t pix_cap = std::min<uint64_t>(chw + T0, T0 + budget) + DEC_SLACK;
    const uint64_t lis_shape = (uint64_t)a.C * (a.H / 2 + 2) * (a.W / 2 + 2) * 5 / 4 + T0;
    const uint64_t lis_cap = std::min<uint64_t>(lis_shape, T0 + budget) + DEC_SLACK;
    k.pix_cap = pix_cap;
    k.lis_cap = lis_cap;
    const size_t per_slot = (pix_cap * 2 + lis_cap * 3) * sizeof(uint32_t);
    int rc = ctx->ensure(ctx->lists, per_slot * slots + 256);
    if (rc) return rc;
    rc = ctx->ensure(ctx->misc, 256);
    if (rc) return rc;
    uint32_t *base = static_cast<uint32_t *>(ctx->lists.p);
    k.lis = base;
    k.lip = base + (size_t)slots * lis_cap * 3;
    k.lsp = k.lip + (size_t)slots * pix_cap;
    k.counter = static_cast<unsigned int *>(ctx->misc.p);

    ctx->stage_begin(5);
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(a.out, 0, sizeof(int32_t) * (size_t)a.B * a.C * a.H * a.W, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(k.counter, 0, sizeof(unsigned int), ctx->stream));
    spiht_decode_kernel<<<slots, DEC_NT, 0, ctx->stream>>>(k);
    ctx->launches++;
    ctx->stage_end(5);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
