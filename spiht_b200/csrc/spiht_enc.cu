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
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace spihtb {

constexpr int ENC_NT = 512;
constexpr int ENC_REF_ITEMS = 8;                      // refinement: entries per thread per flush
constexpr int ENC_STAGE_WORDS = ENC_NT * 9 / 32 + 4;  // a chunk emits at most 9 bits per thread
constexpr int ENC_SLACK = 16 * ENC_NT;                // list slack for the chunk that crosses the budget

struct EncK {
    const int32_t *coeffs;
    int B, C, H, W, NH, NW, ll_h, ll_w;
    KeyFmt kf;
    const uint8_t *dp, *lp, *dpll, *lpll;
    const uint32_t *maxabs;
    uint64_t max_bits;
    const uint64_t *dev_max_bits;
    uint32_t *out;
    uint64_t out_stride_words;
    uint64_t *nbits;
    int32_t *max_n;
    int32_t *status;
    // per-slot list storage
    int32_t *lip;
    uint32_t *lsp;
    uint2 *lis;  // 3 buffers per slot: R, G0, G1
    size_t pix_cap, lis_cap;
    unsigned int *counter;
};

__device__ __forceinline__ void bw_emit(uint32_t *stage, uint64_t wbase, uint64_t limit, uint64_t off, uint32_t val,
                                        int nb)
{
    if (nb == 0 || off >= limit) return;
    if (off + (uint64_t)nb > limit) {
        nb = (int)(limit - off);
        val &= (1u << nb) - 1u;  // here 1 <= nb < 32
    }
    const uint32_t rel = (uint32_t)((off >> 5) - wbase);
    const int sh = (int)(off & 31);
    atomicOr(&stage[rel], val << sh);
    if (sh + nb > 32) atomicOr(&stage[rel + 1], val >> (32 - sh));
}

// Write out the complete words of the staging window; keep the partial one.
// Callers must place a __syncthreads() (or a block_exscan) before the next emit.
__device__ __forceinline__ void bw_flush(uint32_t *stage, uint32_t *outrow, uint64_t &wbase, uint64_t end)
{
    __syncthreads();
    const uint32_t nfull = (uint32_t)((end >> 5) - wbase);
    for (uint32_t i = threadIdx.x; i < nfull; i += ENC_NT) outrow[wbase + i] = stage[i];
    const uint32_t carry = stage[nfull];
    __syncthreads();
    if (nfull) {
        for (uint32_t i = threadIdx.x; i <= nfull; i += ENC_NT) stage[i] = i ? 0u : carry;
    }
    wbase += nfull;
}

__global__ void __launch_bounds__(ENC_NT) spiht_encode_kernel(const EncK p)
{
    __shared__ uint32_t s_stage[ENC_STAGE_WORDS];
    __shared__ uint64_t s_wtot[ENC_NT / 32];
    __shared__ int s_img;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const KeyFmt kf = p.kf;
    const uint32_t H = p.H, W = p.W, NH = p.NH, NW = p.NW, ll_h = p.ll_h, ll_w = p.ll_w, C = p.C;

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

        for (int i = tid; i < ENC_STAGE_WORDS; i += ENC_NT) s_stage[i] = 0;
        uint64_t wbase = 0, bitpos = 0;

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
            const uint64_t ex = block_exscan<ENC_NT>(root ? 1ull : 0ull, s_wtot, tot);
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

            // ---- LIP pass (encoder_decoder.rs:207-222): in-place stable compaction
            uint32_t keep = 0;
            for (uint32_t base = 0; base < lip_len && !done; base += ENC_NT) {
                const uint32_t e = base + tid;
                const bool valid = e < lip_len;
                const int32_t v = valid ? lip[e] : 0;
                const bool sig = valid && absu(v) >= thr;
                const int nb = valid ? (sig ? 2 : 1) : 0;
                const uint32_t val = sig ? (1u | ((v >= 0) ? 2u : 0u)) : 0u;
                const uint64_t pack = (uint64_t)nb | ((uint64_t)(valid && !sig) << 14) | ((uint64_t)sig << 25);
                uint64_t tot;
                const uint64_t ex = block_exscan<ENC_NT>(pack, s_wtot, tot);
                bw_emit(s_stage, wbase, limit, bitpos + (ex & 0x3fff), val, nb);
                if (valid && !sig) lip[keep + (uint32_t)((ex >> 14) & 0x7ff)] = v;
                if (sig) lsp[lsp_len + (uint32_t)((ex >> 25) & 0x1fff)] = absu(v);
                keep += (uint32_t)((tot >> 14) & 0x7ff);
                lsp_len += (uint32_t)((tot >> 25) & 0x1fff);
                bitpos += tot & 0x3fff;
                bw_flush(s_stage, outrow, wbase, bitpos < limit ? bitpos : limit);
                done = bitpos >= limit;
            }
            if (done) break;
            lip_len = keep;

            // ---- LIS pass (encoder_decoder.rs:224-284), generation by generation
            {
                uint2 *cur = R, *nxt = G0;
                uint32_t cur_len = r_len, rkeep = 0;
                int gen = 0;
                while (cur_len > 0 && !done) {
                    uint32_t nxt_len = 0;
                    for (uint32_t base = 0; base < cur_len && !done; base += ENC_NT) {
                        const uint32_t e = base + tid;
                        const bool valid = e < cur_len;
                        const uint2 ent = valid ? cur[e] : make_uint2(0u, 0u);
                        const uint32_t key = ent.x;
                        const bool isA = (key >> 31) != 0;
                        uint32_t k, i, j;
                        key_unpack(kf, key, k, i, j);
                        const bool fire = valid && ent.y >= (uint32_t)(n + 1);
                        uint32_t val = fire ? 1u : 0u;
                        int nb = valid ? 1 : 0;
                        uint32_t nlsp = 0, nlip = 0, nnext = 0, sigmask = 0;
                        int32_t x[4] = {0, 0, 0, 0};
                        uint32_t nfp[4] = {0, 0, 0, 0};
                        uint32_t ci = 0, cj = 0;
                        if (fire) {
                            offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj);
                            if (isA) {
                                const int32_t *a = img + ((size_t)k * H + ci) * W + cj;
                                x[0] = a[0];
                                x[1] = a[1];
                                x[2] = a[W];
                                x[3] = a[W + 1];
                                if (has_desc_past_offspring(i, j, H, W)) {
                                    nnext = 1;
                                    if (i < ll_h && j < ll_w)
                                        nfp[0] = lpll[((size_t)k * ll_h + i) * ll_w + j];
                                    else if (i < NH && j < NW)
                                        nfp[0] = lp[((size_t)k * NH + i) * NW + j];
                                }
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    const bool sg = absu(x[r]) >= thr;
                                    val |= (uint32_t)sg << nb;
                                    ++nb;
                                    if (sg) {
                                        val |= (uint32_t)(x[r] >= 0) << nb;
                                        ++nb;
                                        ++nlsp;
                                        sigmask |= 1u << r;
                                    }
                                }
                                nlip = 4 - nlsp;
                            } else {
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    const uint32_t y = ci + (r >> 1), xx = cj + (r & 1);
                                    if (y < NH && xx < NW) nfp[r] = dp[((size_t)k * NH + y) * NW + xx];
                                }
                                nnext = 4;
                            }
                        }
                        const uint64_t pack = (uint64_t)nb | ((uint64_t)(valid && !fire) << 14) |
                                              ((uint64_t)nlsp << 25) | ((uint64_t)nlip << 38) |
                                              ((uint64_t)nnext << 51);
                        uint64_t tot;
                        const uint64_t ex = block_exscan<ENC_NT>(pack, s_wtot, tot);
                        bw_emit(s_stage, wbase, limit, bitpos + (ex & 0x3fff), val, nb);
                        if (valid && !fire) R[rkeep + (uint32_t)((ex >> 14) & 0x7ff)] = ent;
                        if (fire) {
                            uint32_t on = nxt_len + (uint32_t)(ex >> 51);
                            if (isA) {
                                uint32_t os = lsp_len + (uint32_t)((ex >> 25) & 0x1fff);
                                uint32_t oi = lip_len + (uint32_t)((ex >> 38) & 0x1fff);
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    if (sigmask & (1u << r))
                                        lsp[os++] = absu(x[r]);
                                    else
                                        lip[oi++] = x[r];
                                }
                                if (nnext) nxt[on] = make_uint2(key & 0x7fffffffu, nfp[0]);
                            } else {
#pragma unroll
                                for (int r = 0; r < 4; ++r)
                                    nxt[on + r] = make_uint2(
                                        0x80000000u | key_pack(kf, k, ci + (r >> 1), cj + (r & 1)), nfp[r]);
                            }
                        }
                        rkeep += (uint32_t)((tot >> 14) & 0x7ff);
                        lsp_len += (uint32_t)((tot >> 25) & 0x1fff);
                        lip_len += (uint32_t)((tot >> 38) & 0x1fff);
                        nxt_len += (uint32_t)(tot >> 51);
                        bitpos += tot & 0x3fff;
                        bw_flush(s_stage, outrow, wbase, bitpos < limit ? bitpos : limit);
                        done = bitpos >= limit;
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
                __syncthreads();  // staging window reset by the previous flush is complete
#pragma unroll
                for (int r = 0; r < ENC_REF_ITEMS; ++r) {
                    const uint32_t e = base + r * ENC_NT + tid;
                    const uint32_t v = e < lsp_len0 ? lsp[e] : 0u;
                    const uint32_t word = __ballot_sync(0xffffffffu, (v >> n) & 1u);
                    const uint32_t e0 = e - lane;
                    if (lane == 0 && e0 < lsp_len0) {
                        const uint32_t cnt = min(32u, lsp_len0 - e0);
                        bw_emit(s_stage, wbase, limit, bitpos + (e0 - base),
                                cnt < 32 ? (word & ((1u << cnt) - 1u)) : word, (int)cnt);
                    }
                }
                bitpos += min((uint32_t)(ENC_NT * ENC_REF_ITEMS), lsp_len0 - base);
                bw_flush(s_stage, outrow, wbase, bitpos < limit ? bitpos : limit);
                done = bitpos >= limit;
            }
            if (done || n == 0) break;
        }

        // ---- finish the stream
        __syncthreads();
        const uint64_t end = bitpos < limit ? bitpos : limit;
        if (tid == 0) {
            if (end & 31) outrow[wbase] = s_stage[0];
            p.nbits[b] = end;
            p.max_n[b] = max_n;
            if (p.status) p.status[b] = (bitpos >= limit && want > cap_bits) ? 1 : 0;
        }
        __syncthreads();
    }
}

int launch_encode(spihtb_ctx *ctx, const EncArgs &a)
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

    int occ = 1;
    SPIHTB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spiht_encode_kernel, ENC_NT, 0));
    if (occ < 1) occ = 1;
    const int slots = std::min(a.B, ctx->sm_count * occ);

    // list capacities: bounded by the shape and by the bit budget (every entry
    // created costs at least one emitted bit), see DESIGN.md
    const uint64_t T0 = (uint64_t)a.ll_h * a.ll_w * a.C;
    const uint64_t chw = (uint64_t)a.C * a.H * a.W;
    uint64_t budget = a.dev_max_bits ? a.out_stride * 8 : (a.max_bits == 0 ? ~0ull : a.max_bits);
    budget = std::min<uint64_t>(budget, a.out_stride * 8);
    const uint64_t pix_cap = std::min<uint64_t>(chw + T0, T0 + budget) + ENC_SLACK;
    const uint64_t lis_shape = (uint64_t)a.C * (a.H / 2 + 2) * (a.W / 2 + 2) * 5 / 4 + T0;
    const uint64_t lis_cap = std::min<uint64_t>(lis_shape, T0 + budget) + ENC_SLACK;
    k.pix_cap = pix_cap;
    k.lis_cap = lis_cap;
    const size_t per_slot = pix_cap * 8 + lis_cap * 3 * sizeof(uint2);
    int rc = ctx->ensure(ctx->lists, per_slot * slots + 256);
    if (rc) return rc;
    rc = ctx->ensure(ctx->misc, 256);
    if (rc) return rc;
    uint8_t *base = static_cast<uint8_t *>(ctx->lists.p);
    k.lis = reinterpret_cast<uint2 *>(base);
    k.lip = reinterpret_cast<int32_t *>(base + (size_t)slots * lis_cap * 3 * sizeof(uint2));
    k.lsp = reinterpret_cast<uint32_t *>(base + (size_t)slots * (lis_cap * 3 * sizeof(uint2) + pix_cap * 4));
    k.counter = static_cast<unsigned int *>(ctx->misc.p);
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(k.counter, 0, sizeof(unsigned int), ctx->stream));
    ctx->stage_begin(4);
    spiht_encode_kernel<<<slots, ENC_NT, 0, ctx->stream>>>(k);
    ctx->launches++;
    ctx->stage_end(4);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
