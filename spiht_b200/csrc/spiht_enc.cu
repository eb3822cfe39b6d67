// SPIHT encoder (replaces src/encoder_decoder.rs:155-303), one CTA per image.
//
// The reference walks three FIFO lists one entry and one bit at a time.  Here
// each pass over a list is a data-parallel map + CTA-wide exclusive scan:
//   map   : every thread takes one list entry and works out its record
//           (1..9 bits) and what it appends to the other lists;
//   scan  : one packed 64-bit exclusive scan gives the record's bit offset and
//           the append offsets (order-preserving stream compaction);
//   write : bits are OR-ed into a shared-memory staging window and flushed to
//           HBM as whole 32-bit words; list appends go out compacted.
// The FIFO LIS pass is run generation by generation (generation g+1 = what
// generation g pushed), which reproduces the reference's queue order.
// Subtree significance comes from the pyramid (pyramid.cu): an A entry carries
// the plane at which its D-set fires, a B entry the plane of its L-set, LIP
// entries carry the coefficient and LSP entries its magnitude, so a pass never
// re-reads the coefficient array.  Truncation is exact: bits at stream
// positions >= max_bits are dropped (push_bit!, encoder_decoder.rs:192-201).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "spiht_enc.cuh"

namespace spihtb {

// Phase cycle counters of image 0's CTA: debug builds only (-DSPIHTB_PROF, python -m spiht_b200.build -DSPIHTB_PROF).
// Release builds compile them out entirely; the symbols below are not part of the C ABI.
#ifdef SPIHTB_PROF
__device__ unsigned long long g_enc_prof[16];
extern "C" int spihtb_debug_enc_prof(unsigned long long *out16)
{
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out16, g_enc_prof, sizeof(z)) != cudaSuccess) return SPIHTB_ECUDA;
    if (cudaMemcpyToSymbol(g_enc_prof, z, sizeof(z)) != cudaSuccess) return SPIHTB_ECUDA;
    return SPIHTB_OK;
}
#define ENC_T0() const long long _t0 = clock64(); [[maybe_unused]] long long _tl = _t0
#define ENC_LAP(slot)                                                   \
    do {                                                                \
        const long long _n = clock64();                                 \
        if (tid == 0 && b == 0) g_enc_prof[slot] += (unsigned long long)(_n - _tl); \
        _tl = _n;                                                       \
    } while (0)
#define ENC_ADD(slot, cnt)                                                              \
    do {                                                                                \
        if (tid == 0 && b == 0) {                                                       \
            g_enc_prof[slot] += (unsigned long long)(clock64() - _t0);                  \
            g_enc_prof[slot + 8] += (cnt);                                              \
        }                                                                               \
    } while (0)
#define ENC_IMG_T0() const long long _timg = clock64()
#define ENC_IMG_ADD()                                                                        \
    do {                                                                                     \
        if (b == 0) g_enc_prof[3] += (unsigned long long)(clock64() - _timg);                \
    } while (0)
#else
#define ENC_T0() do { } while (0)
#define ENC_LAP(slot) do { } while (0)
#define ENC_ADD(slot, cnt) do { } while (0)
#define ENC_IMG_T0() do { } while (0)
#define ENC_IMG_ADD() do { } while (0)
#endif

__global__ void __launch_bounds__(ENC_NT, 2) spiht_encode_kernel(const EncK p)
{
    __shared__ uint32_t s_ring[ENC_RING];
    __shared__ uint64_t s_scan[2][ENC_NT / 32 + 1];
    __shared__ uint2 s_wa[ENC_CHUNK];          // fired D-sets of the chunk: {key, e | leaves-B << 11 | fired B before << 12}
    __shared__ uint2 s_wb[ENC_CHUNK];          // fired L-sets of the chunk: {key, e | fired A before << 11}
    __shared__ uint16_t s_pab[ENC_CHUNK + 1];  // by fired-A rank: fired A sets before it that leave a B set
    __shared__ uint16_t s_psig[ENC_CHUNK + 1]; // by fired-A rank: significant offspring before it
    __shared__ int s_img;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const KeyFmt kf = p.kf;
    const uint32_t H = p.H, W = p.W, NH = p.NH, NW = p.NW, ll_h = p.ll_h, ll_w = p.ll_w, C = p.C;
    int parity = 0;

    int32_t *lip = p.lip + (size_t)blockIdx.x * p.pix_cap;
    uint32_t *lsp = p.lsp + (size_t)blockIdx.x * p.pix_cap;
    uint2 *R = p.lis + (size_t)blockIdx.x * 3 * p.lis_cap;
    uint2 *G0 = R + p.lis_cap;
    uint2 *G1 = G0 + p.lis_cap;

    for (;;) {
        if (tid == 0) s_img = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int b = s_img;
        __syncthreads();
        if (b >= p.B) break;

        const int32_t *img = p.coeffs + (size_t)b * C * H * W;
        const uint8_t *dp = p.dp + (size_t)b * C * NH * NW;
        const uint8_t *lp = p.lp + (size_t)b * C * NH * NW;
        const uint8_t *dpll = p.dpll + (size_t)b * C * ll_h * ll_w;
        const uint8_t *lpll = p.lpll + (size_t)b * C * ll_h * ll_w;
        uint32_t *outrow = p.out + (size_t)b * p.out_stride_words;

        const int max_n = max_n_of(p.maxabs[b]);
        uint64_t want = p.dev_max_bits ? p.dev_max_bits[b] : p.max_bits;
        if (want == 0) want = ~0ull;  // the reference's length check never matches 0
        const uint64_t cap_bits = p.out_stride_words * 32ull;
        const uint64_t limit = want < cap_bits ? want : cap_bits;

        for (int i = tid; i < ENC_RING; i += ENC_NT) s_ring[i] = 0;
        uint64_t wflushed = 0, bitpos = 0;
        ENC_IMG_T0();

        // ---- list initialisation (encoder_decoder.rs:170-190): i, j, channel innermost
        const uint32_t T0 = ll_h * ll_w * C;
        uint32_t lip_len = T0, lsp_len = 0, r_len = 0;
        for (uint32_t base = 0; base < T0; base += ENC_NT) {
            const uint32_t t = base + tid;
            const bool valid = t < T0;
            uint32_t k = 0, i = 0, j = 0;
            bool root = false;
            if (valid) {
                k = t % C;
                const uint32_t ij = t / C;
                j = ij % ll_w;
                i = ij / ll_w;
                lip[t] = img[((size_t)k * H + i) * W + j];
                root = ((i | j) & 1u) != 0;
            }
            uint64_t tot;
            const uint64_t ex = block_exscan2<ENC_NT, uint64_t>(root ? 1ull : 0ull, s_scan, parity, tot);
            if (root)
                R[r_len + (uint32_t)ex] =
                    make_uint2(0x80000000u | key_pack(kf, k, i, j), dpll[((size_t)k * ll_h + i) * ll_w + j]);
            r_len += (uint32_t)tot;
        }
        __syncthreads();

        bool done = false;
        for (int n = max_n;; --n) {
            const uint32_t thr = 1u << n;
            const uint32_t lsp_len0 = lsp_len;

            // ---- LIP pass (encoder_decoder.rs:207-222): in-place stable compaction.
            // A thread takes ENC_ITEMS consecutive entries: record = "0" | "1 sign".
            uint32_t keep = 0;
            for (uint32_t base = 0; base < lip_len && !done; base += ENC_CHUNK) {
                ENC_T0();
                const uint32_t e0 = base + tid * ENC_ITEMS;
                const uint32_t nval = e0 < lip_len ? min((uint32_t)ENC_ITEMS, lip_len - e0) : 0u;
                int32_t v[ENC_ITEMS];
                if (nval == ENC_ITEMS) {
                    const int4 q = *reinterpret_cast<const int4 *>(lip + e0);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int t = 0; t < ENC_ITEMS; ++t) v[t] = (uint32_t)t < nval ? lip[e0 + t] : 0;
                }
                uint32_t sigm = 0, nb = 0;
                uint64_t val = 0;
#pragma unroll
                for (int t = 0; t < ENC_ITEMS; ++t) {
                    if ((uint32_t)t < nval) {
                        if (absu(v[t]) >= thr) {
                            sigm |= 1u << t;
                            val |= (uint64_t)(1u | ((v[t] >= 0) ? 2u : 0u)) << nb;
                            nb += 2;
                        } else {
                            nb += 1;
                        }
                    }
                }
                const uint32_t nsig = __popc(sigm);
                // one scan: entries before this thread (valid ones are dense) and significant ones before it
                uint64_t tot;
                const uint64_t ex = block_exscan2<ENC_NT, uint64_t>((uint64_t)nsig | ((uint64_t)nval << 32), s_scan,
                                                                    parity, tot);
                const uint32_t sig_before = (uint32_t)ex, val_before = (uint32_t)(ex >> 32);
                bw_emit(s_ring, limit, bitpos + val_before + sig_before, val, (int)nb);
                uint32_t ok = keep + val_before - sig_before, os = lsp_len + sig_before;
#pragma unroll
                for (int t = 0; t < ENC_ITEMS; ++t) {
                    if ((uint32_t)t < nval) {
                        if (sigm & (1u << t))
                            lsp[os++] = absu(v[t]);
                        else
                            lip[ok++] = v[t];
                    }
                }
                const uint32_t tsig = (uint32_t)tot, tval = (uint32_t)(tot >> 32);
                keep += tval - tsig;
                lsp_len += tsig;
                bitpos += tval + tsig;
                bw_flush(s_ring, outrow, wflushed, bitpos < limit ? bitpos : limit);
                done = bitpos >= limit;
                ENC_ADD(0, 1);
            }
            if (done) break;
            lip_len = keep;

            // ---- LIS pass (encoder_decoder.rs:224-284), generation by generation.
            // A chunk of the queue is handled in two steps so that warps never diverge over the set type:
            //   classify: every thread looks at ENC_ITEMS consecutive entries (fires <=> firing plane > n);
            //             one packed scan ranks the fired A sets, the fired B sets and the retained ones;
            //             retained entries go straight back to R, fired ones into two dense shared work lists;
            //   dense A : one fired D-set per thread -- its four offspring coefficients (and the firing plane
            //             of the L-set it leaves) are loaded by all lanes at once; a second scan over the
            //             number of significant offspring gives the record's bit offset and the LSP/LIP slots;
            //   dense B : one fired L-set per thread -- "1" bit, four child D-sets with their firing planes.
            // An entry's bit position is  bitpos + e + 4 (fired A before it) + (significant offspring before it)
            // (e = its index in the chunk: every entry costs one bit); "0" bits are never written.
            {
                uint2 *cur = R, *nxt = G0;
                uint32_t cur_len = r_len, rkeep = 0;
                int gen = 0;
                while (cur_len > 0 && !done) {
                    uint32_t nxt_len = 0;
                    for (uint32_t base = 0; base < cur_len && !done; base += ENC_CHUNK) {
                        ENC_T0();
                        const uint32_t chunk_n = min((uint32_t)ENC_CHUNK, cur_len - base);
                        const uint32_t el = tid * ENC_ITEMS;  // first entry of this thread, chunk-relative
                        const uint32_t nval = el < chunk_n ? min((uint32_t)ENC_ITEMS, chunk_n - el) : 0u;
                        uint2 ent[ENC_ITEMS];
                        if (nval == ENC_ITEMS) {
                            const uint4 q0 = *reinterpret_cast<const uint4 *>(cur + base + el);
                            const uint4 q1 = *reinterpret_cast<const uint4 *>(cur + base + el + 2);
                            ent[0] = make_uint2(q0.x, q0.y); ent[1] = make_uint2(q0.z, q0.w);
                            ent[2] = make_uint2(q1.x, q1.y); ent[3] = make_uint2(q1.z, q1.w);
                        } else {
#pragma unroll
                            for (int t = 0; t < ENC_ITEMS; ++t)
                                ent[t] = (uint32_t)t < nval ? cur[base + el + t] : make_uint2(0u, 0u);
                        }
                        // ---- classify
                        uint32_t firem = 0, amask = 0, bnext = 0;  // bnext: fired A sets that leave a B set
#pragma unroll
                        for (int t = 0; t < ENC_ITEMS; ++t) {
                            if ((uint32_t)t < nval && ent[t].y >= (uint32_t)(n + 1)) {
                                firem |= 1u << t;
                                if (ent[t].x >> 31) {
                                    amask |= 1u << t;
                                    uint32_t k, i, j;
                                    key_unpack(kf, ent[t].x, k, i, j);
                                    if (has_desc_past_offspring(i, j, H, W)) bnext |= 1u << t;
                                }
                            }
                        }
                        ENC_LAP(4);
                        uint64_t tot;
                        const uint64_t pack = (uint64_t)__popc(amask) | ((uint64_t)__popc(bnext) << 16) |
                                              ((uint64_t)__popc(firem & ~amask) << 32);
                        const uint64_t ex = block_exscan2<ENC_NT, uint64_t>(pack, s_scan, parity, tot);
                        const uint32_t t_fa = (uint32_t)(tot & 0xffff), t_fab = (uint32_t)((tot >> 16) & 0xffff),
                                       t_fb = (uint32_t)((tot >> 32) & 0xffff);
                        {
                            uint32_t x_fa = (uint32_t)(ex & 0xffff), x_fab = (uint32_t)((ex >> 16) & 0xffff),
                                     x_fb = (uint32_t)((ex >> 32) & 0xffff);
                            uint32_t ok = rkeep + el - x_fa - x_fb;
#pragma unroll
                            for (int t = 0; t < ENC_ITEMS; ++t) {
                                if ((uint32_t)t < nval) {
                                    const uint32_t e = el + t;
                                    if (!(firem & (1u << t))) {
                                        R[ok++] = ent[t];
                                    } else if (amask & (1u << t)) {
                                        const uint32_t bn = (bnext >> t) & 1u;
                                        s_wa[x_fa] = make_uint2(ent[t].x, e | (bn << 11) | (x_fb << 12));
                                        s_pab[x_fa] = (uint16_t)x_fab;
                                        ++x_fa;
                                        x_fab += bn;
                                    } else {
                                        s_wb[x_fb] = make_uint2(ent[t].x, e | (x_fa << 11));
                                        ++x_fb;
                                    }
                                }
                            }
                            if (tid == 0) s_pab[t_fa] = (uint16_t)t_fab;
                        }
                        __syncthreads();
                        ENC_LAP(5);
                        // ---- dense A
                        uint32_t sig_carry = 0;
                        for (uint32_t a0 = 0; a0 < t_fa; a0 += ENC_NT) {
                            const uint32_t a = a0 + tid;
                            const bool valid = a < t_fa;
                            uint2 wa = make_uint2(0u, 0u);
                            int32_t xs[4] = {0, 0, 0, 0};
                            uint32_t fnext = 0;
                            if (valid) {
                                wa = s_wa[a];
                                uint32_t k, i, j, ci = 0, cj = 0;
                                key_unpack(kf, wa.x, k, i, j);
                                offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj);
                                const int32_t *x = img + ((size_t)k * H + ci) * W + cj;
                                xs[0] = x[0];
                                xs[1] = x[1];
                                xs[2] = x[W];
                                xs[3] = x[W + 1];
                                if (wa.y & (1u << 11)) {
                                    if (i < ll_h && j < ll_w)
                                        fnext = lpll[((size_t)k * ll_h + i) * ll_w + j];
                                    else if (i < NH && j < NW)
                                        fnext = lp[((size_t)k * NH + i) * NW + j];
                                }
                            }
                            uint32_t sigm = 0, rb = 1;
                            uint32_t rec = valid ? 1u : 0u;
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const bool sg = absu(xs[r]) >= thr;
                                rec |= (uint32_t)sg << rb;
                                ++rb;
                                if (sg) {
                                    rec |= (uint32_t)(xs[r] >= 0) << rb;
                                    ++rb;
                                    sigm |= 1u << r;
                                }
                            }
                            const uint32_t nl = valid ? __popc(sigm) : 0u;
                            ENC_LAP(12);
                            uint64_t tot2;
                            const uint32_t x_sig =
                                sig_carry + (uint32_t)block_exscan2<ENC_NT, uint64_t>((uint64_t)nl, s_scan, parity, tot2);
                            ENC_LAP(13);
                            if (valid) {
                                s_psig[a] = (uint16_t)x_sig;
                                const uint32_t e = wa.y & 0x7ffu;
                                bw_emit(s_ring, limit, bitpos + e + 4 * a + x_sig, rec, (int)rb);
                                uint32_t os = lsp_len + x_sig, oi = lip_len + 4 * a - x_sig;
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    if (sigm & (1u << r))
                                        lsp[os++] = absu(xs[r]);
                                    else
                                        lip[oi++] = xs[r];
                                }
                                if (wa.y & (1u << 11))
                                    nxt[nxt_len + s_pab[a] + 4 * (wa.y >> 12)] = make_uint2(wa.x & 0x7fffffffu, fnext);
                            }
                            sig_carry += (uint32_t)tot2;
                            ENC_LAP(14);
                        }
                        if (tid == 0) s_psig[t_fa] = (uint16_t)sig_carry;
                        __syncthreads();
                        ENC_LAP(6);
                        // ---- dense B
                        for (uint32_t b0 = 0; b0 < t_fb; b0 += ENC_NT) {
                            const uint32_t bi = b0 + tid;
                            if (bi < t_fb) {
                                const uint2 wb = s_wb[bi];
                                const uint32_t e = wb.y & 0x7ffu, xa = wb.y >> 11;
                                uint32_t k, i, j, ci = 0, cj = 0;
                                key_unpack(kf, wb.x, k, i, j);
                                offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj);
                                uint32_t f[4];
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    const uint32_t y = ci + (r >> 1), xx = cj + (r & 1);
                                    f[r] = (y < NH && xx < NW) ? (uint32_t)dp[((size_t)k * NH + y) * NW + xx] : 0u;
                                }
                                bw_emit(s_ring, limit, bitpos + e + 4 * xa + s_psig[xa], 1ull, 1);
                                uint2 *o = nxt + nxt_len + s_pab[xa] + 4 * bi;
#pragma unroll
                                for (int r = 0; r < 4; ++r)
                                    o[r] = make_uint2(0x80000000u | key_pack(kf, k, ci + (r >> 1), cj + (r & 1)), f[r]);
                            }
                        }
                        ENC_LAP(7);
                        rkeep += chunk_n - t_fa - t_fb;
                        lsp_len += sig_carry;
                        lip_len += 4 * t_fa - sig_carry;
                        nxt_len += t_fab + 4 * t_fb;
                        bitpos += chunk_n + 4 * t_fa + sig_carry;
                        bw_flush(s_ring, outrow, wflushed, bitpos < limit ? bitpos : limit);
                        ENC_LAP(11);
                        done = bitpos >= limit;
                        ENC_ADD(1, 1);
                    }
                    uint2 *old = cur;
                    cur = nxt;
                    nxt = gen == 0 ? G1 : old;
                    cur_len = nxt_len;
                    ++gen;
                }
                r_len = rkeep;
            }
            if (done) break;

            // ---- refinement (encoder_decoder.rs:287-292): one bit per older LSP entry
            for (uint32_t base = 0; base < lsp_len0 && !done; base += ENC_NT * ENC_REF_ITEMS) {
                ENC_T0();
#pragma unroll
                for (int r = 0; r < ENC_REF_ITEMS; ++r) {
                    const uint32_t e = base + r * ENC_NT + tid;
                    const uint32_t v = e < lsp_len0 ? lsp[e] : 0u;
                    const uint32_t word = __ballot_sync(0xffffffffu, (v >> n) & 1u);
                    const uint32_t eb = e - lane;
                    if (lane == 0 && eb < lsp_len0) {
                        const uint32_t cnt = min(32u, lsp_len0 - eb);
                        bw_emit(s_ring, limit, bitpos + (eb - base), cnt < 32 ? (word & ((1u << cnt) - 1u)) : word,
                                (int)cnt);
                    }
                }
                bitpos += min((uint32_t)(ENC_NT * ENC_REF_ITEMS), lsp_len0 - base);
                bw_flush(s_ring, outrow, wflushed, bitpos < limit ? bitpos : limit);
                done = bitpos >= limit;
                ENC_ADD(2, 1);
            }
            if (done || n == 0) break;
        }

        // ---- finish the stream
        __syncthreads();
        const uint64_t end = bitpos < limit ? bitpos : limit;
        if (tid == 0) {
            if (end & 31) outrow[wflushed] = s_ring[wflushed & (ENC_RING - 1)];
            p.nbits[b] = end;
            p.max_n[b] = max_n;
            if (p.status) p.status[b] = (bitpos >= limit && want > cap_bits) ? 1 : 0;
            ENC_IMG_ADD();
        }
        __syncthreads();
    }
}

// List capacities: bounded by the shape and by the bit budget (every entry created costs at least one
// emitted bit), see DESIGN.md.  max_slots: CTAs of the coder that can be resident at once.
int plan_encode(spihtb_ctx *ctx, const EncArgs &a, EncPlan *pl)
{
    int occ = 1;
    SPIHTB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spiht_encode_kernel, ENC_NT, 0));
    if (occ < 1) occ = 1;
    pl->max_slots = ctx->sm_count * occ;
    const uint64_t T0 = (uint64_t)a.ll_h * a.ll_w * a.C;
    const uint64_t chw = (uint64_t)a.C * a.H * a.W;
    uint64_t budget = a.dev_max_bits ? a.out_stride * 8 : (a.max_bits == 0 ? ~0ull : a.max_bits);
    budget = std::min<uint64_t>(budget, a.out_stride * 8);
    pl->pix_cap = (std::min<uint64_t>(chw + T0, T0 + budget) + ENC_SLACK + 3) / 4 * 4;
    const uint64_t lis_shape = (uint64_t)a.C * (a.H / 2 + 2) * (a.W / 2 + 2) * 5 / 4 + T0;
    pl->lis_cap = (std::min<uint64_t>(lis_shape, T0 + budget) + ENC_SLACK + 3) / 4 * 4;
    pl->per_slot = pl->pix_cap * 8 + pl->lis_cap * 3 * sizeof(uint2);
    return SPIHTB_OK;
}

// `lists`: a region of per_slot * slots bytes (256-byte aligned); `counter`: one zero-initialised-here word.
int launch_encode(spihtb_ctx *ctx, const EncArgs &a, const EncPlan &pl, void *lists, int slots, unsigned int *counter)
{
    EncK k;
    if (!make_keyfmt(a.C, a.H, a.W, &k.kf)) {
        set_error("shape c=%d h=%d w=%d does not fit a 31-bit packed list entry", a.C, a.H, a.W);
        return SPIHTB_ESHAPE;
    }
    if ((a.out_stride & 7) != 0 || a.out_stride == 0) {
        set_error("out_stride must be a positive multiple of 8 bytes");
        return SPIHTB_EINVAL;
    }
    k.coeffs = a.coeffs;
    k.B = a.B; k.C = a.C; k.H = a.H; k.W = a.W; k.NH = a.H / 2; k.NW = a.W / 2;
    k.ll_h = a.ll_h; k.ll_w = a.ll_w;
    k.dp = a.dp; k.lp = a.lp; k.dpll = a.dpll; k.lpll = a.lpll;
    k.maxabs = a.maxabs;
    k.max_bits = a.max_bits;
    k.dev_max_bits = a.dev_max_bits;
    k.out = reinterpret_cast<uint32_t *>(a.out);
    k.out_stride_words = a.out_stride / 4;
    k.nbits = a.nbits; k.max_n = a.max_n; k.status = a.status;
    k.pix_cap = pl.pix_cap;
    k.lis_cap = pl.lis_cap;
    uint8_t *base = static_cast<uint8_t *>(lists);
    k.lis = reinterpret_cast<uint2 *>(base);
    k.lip = reinterpret_cast<int32_t *>(base + (size_t)slots * pl.lis_cap * 3 * sizeof(uint2));
    k.lsp = reinterpret_cast<uint32_t *>(base + (size_t)slots * (pl.lis_cap * 3 * sizeof(uint2) + pl.pix_cap * 4));
    k.counter = counter;
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(k.counter, 0, sizeof(unsigned int), ctx->stream));
    ctx->stage_begin(4);
    spiht_encode_kernel<<<slots, ENC_NT, 0, ctx->stream>>>(k);
    ctx->launches++;
    ctx->stage_end(4);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

// Cluster size for the cluster coder (spiht_enc_cl.cu), or 0 for one CTA per image.  A cluster pays two cluster
// barriers per 2048-entry chunk, so it only wins when the lists are long (large budgets) and the batch leaves
// SMs idle: budgets of at least 4 Mbit and at most sm_count / 2 images.  SPIHTB_ENC_CLUSTER=N forces N (0: never).
static int cluster_size_for(spihtb_ctx *ctx, const EncArgs &a)
{
    if (const char *e = getenv("SPIHTB_ENC_CLUSTER")) {
        const int v = atoi(e);
        return v >= 2 ? std::min(v, 16) : 0;
    }
    uint64_t budget = a.dev_max_bits ? a.out_stride * 8 : (a.max_bits == 0 ? ~0ull : a.max_bits);
    budget = std::min<uint64_t>(budget, a.out_stride * 8);
    if (budget < (1ull << 22) || a.B * 2 > ctx->sm_count) return 0;
    int cl = 16;
    while (cl > 2 && a.B * cl > ctx->sm_count) cl >>= 1;
    return cl;
}

int launch_encode_cluster(spihtb_ctx *ctx, const EncArgs &a, const EncPlan &pl, int CL);

// single launch over the whole batch with the context's workspace
int launch_encode(spihtb_ctx *ctx, const EncArgs &a)
{
    EncPlan pl;
    int rc = plan_encode(ctx, a, &pl);
    if (rc) return rc;
    if (const int cl = cluster_size_for(ctx, a)) return launch_encode_cluster(ctx, a, pl, cl);
    const int slots = std::min(a.B, pl.max_slots);
    rc = ctx->ensure(ctx->lists, pl.per_slot * slots + 256);
    if (rc) return rc;
    rc = ctx->ensure(ctx->misc, 256);
    if (rc) return rc;
    return launch_encode(ctx, a, pl, ctx->lists.p, slots, static_cast<unsigned int *>(ctx->misc.p));
}

}  // namespace spihtb
