// SPIHT encoder, one thread-block CLUSTER per image (replaces src/encoder_decoder.rs:155-303 for large images).
//
// spiht_enc.cu gives an image to one CTA: fine for a batch of hundreds of megapixel images, hopeless for a few
// 8192 x 8192 images at 1 bpp (BASELINE.json configs[3]): 67 Mbit of stream, lists of tens of millions of entries,
// and 4 of 148 SMs busy.  Here the CTAs of a cluster (up to 16, one GPC) share an image.  Every list pass is cut
// into super-chunks of CL x 2048 entries; CTA r takes the r-th chunk and runs the same map + block scan as the
// single-CTA coder; the chunk totals are exchanged through distributed shared memory (every CTA writes its totals
// into the exchange slots of all CTAs, one cluster barrier) and each CTA adds the totals of the lower ranks to its
// offsets.  All pass state (list lengths, bit position, the truncation flag) is replicated: every CTA derives it
// from the same totals, so all CTAs take the same branches and the cluster barriers always match.
//   - lists (LIP, LSP, LIS generations) live in global memory, shared by the cluster;
//   - a CTA stages the bits of its chunk in shared memory and writes whole words to the stream; the two words it
//     may share with its neighbours in the stream go out with atomicOr (the stream row is zeroed first);
//   - a fired D-set needs the number of significant offspring of all sets before it: the dense-A step is split
//     into gather + count (results parked in shared memory), a second exchange, then emit + append.
// Bit-exact with the single-CTA coder and the CPU oracle (tests/test_gpu_spiht.py runs the suite in both modes).
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "spiht_enc.cuh"

namespace cg = cooperative_groups;

namespace spihtb {

constexpr int CL_MAX = 16;  // largest cluster (non-portable size, one GPC)
constexpr int CL_XK = 4;    // 64-bit values per exchange

struct ClSmem {
    uint32_t ring[ENC_RING];
    uint64_t scan[2][ENC_NT / 32 + 1];
    uint64_t xch[2][CL_MAX][CL_XK];   // exchange slots: [parity][rank][value]
    uint2 wa[ENC_CHUNK];              // fired D-sets of the chunk: {key, e | leaves-B << 11 | fired B before << 12}
    uint2 wb[ENC_CHUNK];              // fired L-sets of the chunk: {key, e | fired A before << 11}
    int4 xs[ENC_CHUNK];               // by fired-A rank: the four offspring coefficients
    uint32_t rec[ENC_CHUNK];          // by fired-A rank: record bits | length << 16 | significance mask << 24
    uint32_t fnext[ENC_CHUNK];        // by fired-A rank: firing plane of the L-set it leaves
    uint16_t pab[ENC_CHUNK + 2];      // by fired-A rank: fired A sets before it that leave a B set
    uint16_t psig[ENC_CHUNK + 2];     // by fired-A rank: significant offspring before it (within the chunk)
};

// Write out the words of stream range [start, end) staged in `ring` (and clear them).  Words shared with a
// neighbouring range (the first when start is not word-aligned, the last when end is not) are OR-ed in.
__device__ __forceinline__ void cl_flush_range(uint32_t *ring, uint32_t *outrow, uint64_t start, uint64_t end)
{
    __syncthreads();
    if (end > start) {
        const uint64_t w0 = start >> 5, w1 = (end + 31) >> 5;
        for (uint64_t w = w0 + threadIdx.x; w < w1; w += ENC_NT) {
            const uint32_t v = ring[w & (ENC_RING - 1)];
            ring[w & (ENC_RING - 1)] = 0u;
            if (v) {
                const bool shared_word = (w == w0 && (start & 31)) || (w == w1 - 1 && (end & 31));
                if (shared_word)
                    atomicOr(outrow + w, v);
                else
                    outrow[w] = v;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(ENC_NT, 1) spiht_encode_cluster_kernel(const EncK p)
{
    extern __shared__ __align__(16) unsigned char cl_smem_raw[];
    ClSmem &S = *reinterpret_cast<ClSmem *>(cl_smem_raw);
    cg::cluster_group cl = cg::this_cluster();
    const unsigned CL = cl.num_blocks(), rank = cl.block_rank();
    const int tid = threadIdx.x, lane = tid & 31;
    const KeyFmt kf = p.kf;
    const uint32_t H = p.H, W = p.W, NH = p.NH, NW = p.NW, ll_h = p.ll_h, ll_w = p.ll_w, C = p.C;
    int parity = 0, xpar = 0;
    const unsigned cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;

    int32_t *lip = p.lip + (size_t)cluster_id * p.pix_cap;
    uint32_t *lsp = p.lsp + (size_t)cluster_id * p.pix_cap;
    uint2 *R = p.lis + (size_t)cluster_id * 3 * p.lis_cap;
    uint2 *G0 = R + p.lis_cap;
    uint2 *G1 = G0 + p.lis_cap;

    // every CTA writes `mine` into slot [rank] of all CTAs, one cluster barrier, then sums the lower ranks
    auto xchg = [&](const uint64_t (&mine)[CL_XK], uint64_t (&before)[CL_XK], uint64_t (&total)[CL_XK]) {
        if ((unsigned)tid < CL) {
            uint64_t *dst = cl.map_shared_rank(&S.xch[xpar][rank][0], tid);
#pragma unroll
            for (int k = 0; k < CL_XK; ++k) dst[k] = mine[k];
        }
        cl.sync();
#pragma unroll
        for (int k = 0; k < CL_XK; ++k) before[k] = total[k] = 0;
        for (unsigned r = 0; r < CL; ++r) {
#pragma unroll
            for (int k = 0; k < CL_XK; ++k) {
                const uint64_t v = S.xch[xpar][r][k];
                total[k] += v;
                if (r < rank) before[k] += v;
            }
        }
        xpar ^= 1;
    };

    for (int i = tid; i < ENC_RING; i += ENC_NT) S.ring[i] = 0;
    __syncthreads();

    for (unsigned b = cluster_id; b < (unsigned)p.B; b += n_clusters) {
        const int32_t *img = p.coeffs + (size_t)b * C * H * W;
        const uint8_t *dp = p.dp + (size_t)b * C * NH * NW;
        const uint8_t *lp = p.lp + (size_t)b * C * NH * NW;
        const uint8_t *dpll = p.dpll + (size_t)b * C * ll_h * ll_w;
        const uint8_t *lpll = p.lpll + (size_t)b * C * ll_h * ll_w;
        uint32_t *outrow = p.out + (size_t)b * p.out_stride_words;

        const int max_n = max_n_of(p.maxabs[b]);
        uint64_t want = p.dev_max_bits ? p.dev_max_bits[b] : p.max_bits;
        if (want == 0) want = ~0ull;
        const uint64_t cap_bits = p.out_stride_words * 32ull;
        const uint64_t limit = want < cap_bits ? want : cap_bits;
        uint64_t bitpos = 0;

        // ---- list initialisation (encoder_decoder.rs:170-190) by rank 0; everyone knows the lengths
        const uint32_t T0 = ll_h * ll_w * C;
        uint32_t lip_len = T0, lsp_len = 0;
        uint32_t r_len = C * (ll_h * ll_w - ((ll_h + 1) / 2) * ((ll_w + 1) / 2));
        if (rank == 0) {
            uint32_t rl = 0;
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
                const uint64_t ex = block_exscan2<ENC_NT, uint64_t>(root ? 1ull : 0ull, S.scan, parity, tot);
                if (root)
                    R[rl + (uint32_t)ex] = make_uint2(0x80000000u | key_pack(kf, k, i, j), dpll[((size_t)k * ll_h + i) * ll_w + j]);
                rl += (uint32_t)tot;
            }
        }
        cl.sync();

        bool done = false;
        for (int n = max_n;; --n) {
            const uint32_t thr = 1u << n;
            const uint32_t lsp_len0 = lsp_len;

            // ---- LIP pass (encoder_decoder.rs:207-222)
            {
                uint32_t keep = 0;
                for (uint32_t sbase = 0; sbase < lip_len && !done; sbase += CL * ENC_CHUNK) {
                    const uint32_t e0 = sbase + rank * ENC_CHUNK + tid * ENC_ITEMS;
                    const uint32_t cend_e = min(lip_len, sbase + (rank + 1) * ENC_CHUNK);  // end of this CTA's chunk
                    const uint32_t nval = e0 < cend_e ? min((uint32_t)ENC_ITEMS, cend_e - e0) : 0u;
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
                    uint64_t tot;
                    const uint64_t ex = block_exscan2<ENC_NT, uint64_t>((uint64_t)nsig | ((uint64_t)nval << 32), S.scan, parity, tot);
                    const uint64_t mine[CL_XK] = {tot & 0xffffffffull, tot >> 32, 0, 0};
                    uint64_t bf[CL_XK], tt[CL_XK];
                    xchg(mine, bf, tt);   // cluster barrier: every chunk of the super-chunk is loaded
                    const uint32_t sig_before = (uint32_t)ex + (uint32_t)bf[0], val_before = (uint32_t)(ex >> 32) + (uint32_t)bf[1];
                    bw_emit(S.ring, limit, bitpos + val_before + sig_before, val, (int)nb);
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
                    const uint64_t cstart = bitpos + bf[0] + bf[1], cend = cstart + mine[0] + mine[1];
                    keep += (uint32_t)(tt[1] - tt[0]);
                    lsp_len += (uint32_t)tt[0];
                    bitpos += tt[1] + tt[0];
                    cl_flush_range(S.ring, outrow, min(cstart, limit), min(cend, limit));
                    done = bitpos >= limit;
                }
                if (!done) lip_len = keep;
                cl.sync();   // the compacted LIP / appended LSP are complete before anyone reads them
            }
            if (done) break;

            // ---- LIS pass (encoder_decoder.rs:224-284), generation by generation
            {
                uint2 *cur = R, *nxt = G0;
                uint32_t cur_len = r_len, rkeep = 0;
                int gen = 0;
                while (cur_len > 0 && !done) {
                    uint32_t nxt_len = 0;
                    for (uint32_t sbase = 0; sbase < cur_len && !done; sbase += CL * ENC_CHUNK) {
                        const uint32_t base = sbase + rank * ENC_CHUNK;
                        const uint32_t chunk_n = base < cur_len ? min((uint32_t)ENC_CHUNK, cur_len - base) : 0u;
                        const uint32_t el = tid * ENC_ITEMS;
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
                        uint32_t firem = 0, amask = 0, bnext = 0;
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
                        uint64_t tot;
                        const uint64_t pack = (uint64_t)__popc(amask) | ((uint64_t)__popc(bnext) << 16) |
                                              ((uint64_t)__popc(firem & ~amask) << 32);
                        const uint64_t ex = block_exscan2<ENC_NT, uint64_t>(pack, S.scan, parity, tot);
                        const uint32_t t_fa = (uint32_t)(tot & 0xffff), t_fab = (uint32_t)((tot >> 16) & 0xffff),
                                       t_fb = (uint32_t)((tot >> 32) & 0xffff);
                        const uint64_t mine1[CL_XK] = {t_fa, t_fab, t_fb, chunk_n};
                        uint64_t b1[CL_XK], t1[CL_XK];
                        xchg(mine1, b1, t1);   // cluster barrier: every chunk of the super-chunk is loaded
                        {
                            uint32_t x_fa = (uint32_t)(ex & 0xffff), x_fab = (uint32_t)((ex >> 16) & 0xffff),
                                     x_fb = (uint32_t)((ex >> 32) & 0xffff);
                            uint32_t ok = rkeep + (uint32_t)(b1[3] - b1[0] - b1[2]) + el - x_fa - x_fb;
#pragma unroll
                            for (int t = 0; t < ENC_ITEMS; ++t) {
                                if ((uint32_t)t < nval) {
                                    const uint32_t e = el + t;
                                    if (!(firem & (1u << t))) {
                                        R[ok++] = ent[t];
                                    } else if (amask & (1u << t)) {
                                        const uint32_t bn = (bnext >> t) & 1u;
                                        S.wa[x_fa] = make_uint2(ent[t].x, e | (bn << 11) | (x_fb << 12));
                                        S.pab[x_fa] = (uint16_t)x_fab;
                                        ++x_fa;
                                        x_fab += bn;
                                    } else {
                                        S.wb[x_fb] = make_uint2(ent[t].x, e | (x_fa << 11));
                                        ++x_fb;
                                    }
                                }
                            }
                            if (tid == 0) S.pab[t_fa] = (uint16_t)t_fab;
                        }
                        __syncthreads();
                        // ---- dense A, step 1: gather, count, park
                        uint32_t sig_carry = 0;
                        for (uint32_t a0 = 0; a0 < t_fa; a0 += ENC_NT) {
                            const uint32_t a = a0 + tid;
                            const bool valid = a < t_fa;
                            int32_t xs[4] = {0, 0, 0, 0};
                            uint32_t fn = 0;
                            if (valid) {
                                const uint2 wa = S.wa[a];
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
                                        fn = lpll[((size_t)k * ll_h + i) * ll_w + j];
                                    else if (i < NH && j < NW)
                                        fn = lp[((size_t)k * NH + i) * NW + j];
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
                            uint64_t tot2;
                            const uint32_t x_sig = sig_carry + (uint32_t)block_exscan2<ENC_NT, uint64_t>((uint64_t)nl, S.scan, parity, tot2);
                            if (valid) {
                                S.psig[a] = (uint16_t)x_sig;
                                S.xs[a] = make_int4(xs[0], xs[1], xs[2], xs[3]);
                                S.rec[a] = rec | (rb << 16) | (sigm << 24);
                                S.fnext[a] = fn;
                            }
                            sig_carry += (uint32_t)tot2;
                        }
                        if (tid == 0) S.psig[t_fa] = (uint16_t)sig_carry;
                        const uint64_t mine2[CL_XK] = {sig_carry, 0, 0, 0};
                        uint64_t b2[CL_XK], t2[CL_XK];
                        xchg(mine2, b2, t2);   // (a CTA barrier as well: the parked results are visible)
                        const uint64_t cstart = bitpos + b1[3] + 4 * b1[0] + b2[0];
                        // ---- dense A, step 2: emit, append
                        for (uint32_t a = tid; a < t_fa; a += ENC_NT) {
                            const uint2 wa = S.wa[a];
                            const uint32_t e = wa.y & 0x7ffu, x_sig = S.psig[a];
                            const uint32_t rw = S.rec[a], sigm = rw >> 24;
                            const int4 xq = S.xs[a];
                            const int32_t xs[4] = {xq.x, xq.y, xq.z, xq.w};
                            bw_emit(S.ring, limit, cstart + e + 4 * a + x_sig, rw & 0xffffu, (int)((rw >> 16) & 0xffu));
                            uint32_t os = lsp_len + (uint32_t)b2[0] + x_sig;
                            uint32_t oi = lip_len + 4 * ((uint32_t)b1[0] + a) - ((uint32_t)b2[0] + x_sig);
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                if (sigm & (1u << r))
                                    lsp[os++] = absu(xs[r]);
                                else
                                    lip[oi++] = xs[r];
                            }
                            if (wa.y & (1u << 11))
                                nxt[nxt_len + (uint32_t)b1[1] + S.pab[a] + 4 * ((uint32_t)b1[2] + (wa.y >> 12))] =
                                    make_uint2(wa.x & 0x7fffffffu, S.fnext[a]);
                        }
                        // ---- dense B
                        for (uint32_t bi = tid; bi < t_fb; bi += ENC_NT) {
                            const uint2 wb = S.wb[bi];
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
                            bw_emit(S.ring, limit, cstart + e + 4 * xa + S.psig[xa], 1ull, 1);
                            uint2 *o = nxt + nxt_len + (uint32_t)b1[1] + S.pab[xa] + 4 * ((uint32_t)b1[2] + bi);
#pragma unroll
                            for (int r = 0; r < 4; ++r)
                                o[r] = make_uint2(0x80000000u | key_pack(kf, k, ci + (r >> 1), cj + (r & 1)), f[r]);
                        }
                        const uint64_t cend = cstart + chunk_n + 4 * t_fa + sig_carry;
                        rkeep += (uint32_t)(t1[3] - t1[0] - t1[2]);
                        lsp_len += (uint32_t)t2[0];
                        lip_len += 4 * (uint32_t)t1[0] - (uint32_t)t2[0];
                        nxt_len += (uint32_t)t1[1] + 4 * (uint32_t)t1[2];
                        bitpos += t1[3] + 4 * t1[0] + t2[0];
                        cl_flush_range(S.ring, outrow, min(cstart, limit), min(cend, limit));
                        done = bitpos >= limit;
                    }
                    cl.sync();   // the next generation's list is complete before it is read
                    uint2 *old = cur;
                    cur = nxt;
                    nxt = gen == 0 ? G1 : old;
                    cur_len = nxt_len;
                    ++gen;
                }
                r_len = rkeep;
            }
            if (done) break;

            // ---- refinement (encoder_decoder.rs:287-292)
            constexpr uint32_t RCH = ENC_NT * ENC_REF_ITEMS;
            for (uint32_t sbase = 0; sbase < lsp_len0 && !done; sbase += CL * RCH) {
                const uint32_t base = sbase + rank * RCH;
#pragma unroll
                for (int r = 0; r < ENC_REF_ITEMS; ++r) {
                    const uint32_t e = base + r * ENC_NT + tid;
                    const uint32_t v = e < lsp_len0 ? lsp[e] : 0u;
                    const uint32_t word = __ballot_sync(0xffffffffu, (v >> n) & 1u);
                    const uint32_t eb = e - lane;
                    if (lane == 0 && eb < lsp_len0) {
                        const uint32_t cnt = min(32u, lsp_len0 - eb);
                        bw_emit(S.ring, limit, bitpos + (eb - sbase), cnt < 32 ? (word & ((1u << cnt) - 1u)) : word, (int)cnt);
                    }
                }
                const uint64_t cstart = bitpos + (min(base, lsp_len0) - sbase), cend = bitpos + (min(base + RCH, lsp_len0) - sbase);
                bitpos += min(CL * RCH, lsp_len0 - sbase);
                cl_flush_range(S.ring, outrow, min(cstart, limit), min(cend, limit));
                done = bitpos >= limit;
            }
            if (done || n == 0) break;
        }

        // ---- finish the stream
        if (rank == 0 && tid == 0) {
            p.nbits[b] = bitpos < limit ? bitpos : limit;
            p.max_n[b] = max_n;
            if (p.status) p.status[b] = (bitpos >= limit && want > cap_bits) ? 1 : 0;
        }
        cl.sync();   // the lists are reused by the cluster's next image
    }
}

// One cluster of `CL` CTAs per image.  `lists`: per_slot bytes per cluster.
int launch_encode_cluster(spihtb_ctx *ctx, const EncArgs &a, const EncPlan &pl, int CL)
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
    const int clusters = std::max(1, std::min(a.B, ctx->sm_count / CL));
    int rc = ctx->ensure(ctx->lists, pl.per_slot * clusters + 256);
    if (rc) return rc;
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
    uint8_t *base = static_cast<uint8_t *>(ctx->lists.p);
    k.lis = reinterpret_cast<uint2 *>(base);
    k.lip = reinterpret_cast<int32_t *>(base + (size_t)clusters * pl.lis_cap * 3 * sizeof(uint2));
    k.lsp = reinterpret_cast<uint32_t *>(base + (size_t)clusters * (pl.lis_cap * 3 * sizeof(uint2) + pl.pix_cap * 4));
    k.counter = nullptr;

    static bool attr_set = false;
    if (!attr_set) {
        SPIHTB_CUDA_CHECK(cudaFuncSetAttribute(spiht_encode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)sizeof(ClSmem)));
        SPIHTB_CUDA_CHECK(cudaFuncSetAttribute(spiht_encode_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_set = true;
    }
    // shared words of neighbouring chunks are OR-ed into the stream: the rows start from zero
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(a.out, 0, (size_t)a.B * a.out_stride, ctx->stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    cfg.blockDim = dim3(ENC_NT);
    cfg.dynamicSmemBytes = sizeof(ClSmem);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ctx->stage_begin(4);
    cudaError_t e = cudaLaunchKernelEx(&cfg, spiht_encode_cluster_kernel, k);
    ctx->launches++;
    ctx->stage_end(4);
    if (e != cudaSuccess) {
        set_error("cluster launch (%d CTAs per image) failed: %s", CL, cudaGetErrorString(e));
        cudaGetLastError();
        return SPIHTB_ECUDA;
    }
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
