/*
 * oracle/spiht_model.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Sequential C model of the formulation the CUDA coder uses
 * (spiht_b200/csrc/spiht_enc.cu, spiht_dec.cu), checked bit-for-bit against the
 * faithful restatement in spiht_ref.c by tests/test_oracle_spiht.py:
 *
 *   - a bottom-up descendant-max pyramid replaces the recursive subtree scans
 *     is_set_sig / is_l_sig (encoder_decoder.rs:78-121):
 *        DP(p) = 1 + floor(log2 D(p)),  D(p) = max |x| over all descendants of p
 *        LP(p) = 1 + floor(log2 L(p)),  L(p) = max over grand-descendants
 *     (0 when the set is empty or all zero); a type-A entry fires at the first
 *     plane n <= DP-1, a type-B entry at n <= LP-1;
 *   - lists carry values: LIP holds the coefficient, LSP the magnitude, LIS the
 *     firing plane, so a pass never goes back to the coefficient array;
 *   - the FIFO LIS pass (encoder_decoder.rs:225-283) is run generation by
 *     generation (generation g+1 = what generation g pushed), which yields the
 *     same order as the reference's single queue;
 *
 * (The CUDA decoder writes reconstructed values straight into the coefficient
 *  array like the reference does, so it is checked against spiht_ref_decode.)
 *
 * It also serves as the fast oracle for shapes where the recursive coder is
 * too slow for a unit test.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int spiht_ref_max_n(int32_t max_abs);
int spiht_ref_has_descendents_past_offspring(uint64_t i, uint64_t j, uint64_t h, uint64_t w);
int spiht_ref_get_offspring(uint64_t i, uint64_t j, uint64_t h, uint64_t w,
                            uint64_t ll_h, uint64_t ll_w, uint64_t out[4][2]);

#define MODEL_OK 0
#define MODEL_EBADARG 1
#define MODEL_EGEOM 2   /* LL-rule offspring would fall outside the array */
#define MODEL_ENOMEM 3

static inline uint32_t absu(int32_t x) { return x < 0 ? (uint32_t)(-(int64_t)x) : (uint32_t)x; }
static inline uint8_t plane1(uint32_t v) { return v ? (uint8_t)(32 - __builtin_clz(v)) : 0; }

/* geometry guard shared with the CUDA path: every LL-rule offspring in bounds */
int spiht_model_geom_ok(uint64_t h, uint64_t w, uint64_t ll_h, uint64_t ll_w)
{
    if (ll_h < 2 || ll_w < 2) return 0;
    if (2 * ll_h > h + (ll_h & 1)) return 0;   /* lowest root offspring row: 2 ll_h - 1 (even) or 2 ll_h - 2 (odd) */
    if (2 * ll_w > w + (ll_w & 1)) return 0;
    return 1;
}

/* Pyramid.  dp/lp are [c][nh][nw] with nh=h/2, nw=w/2 (dyadic nodes that have
 * offspring); dpll/lpll are [c][ll_h][ll_w] for the LL roots (LL rule). */
int spiht_model_pyramid(const int32_t *arr, uint64_t c, uint64_t h, uint64_t w,
                        uint64_t ll_h, uint64_t ll_w,
                        uint8_t *dp, uint8_t *lp, uint8_t *dpll, uint8_t *lpll)
{
    uint64_t nh = h / 2, nw = w / 2;
    for (uint64_t k = 0; k < c; ++k) {
        const int32_t *a = arr + k * h * w;
        uint8_t *d = dp + k * nh * nw, *l = lp + k * nh * nw;
        for (uint64_t ii = nh; ii-- > 0;) {
            for (uint64_t jj = nw; jj-- > 0;) {
                if (ii == 0 && jj == 0) { d[0] = l[0] = 0; continue; }
                uint32_t m = absu(a[(2 * ii) * w + 2 * jj]);
                uint32_t t;
                t = absu(a[(2 * ii) * w + 2 * jj + 1]); if (t > m) m = t;
                t = absu(a[(2 * ii + 1) * w + 2 * jj]); if (t > m) m = t;
                t = absu(a[(2 * ii + 1) * w + 2 * jj + 1]); if (t > m) m = t;
                uint8_t dd = plane1(m), ll = 0;
                for (int q = 0; q < 4; ++q) {
                    uint64_t ci = 2 * ii + (q >> 1), cj = 2 * jj + (q & 1);
                    if (ci < nh && cj < nw) { uint8_t x = d[ci * nw + cj]; if (x > ll) ll = x; }
                }
                l[ii * nw + jj] = ll;
                d[ii * nw + jj] = dd > ll ? dd : ll;
            }
        }
        for (uint64_t i = 0; i < ll_h; ++i)
            for (uint64_t j = 0; j < ll_w; ++j) {
                uint64_t off[4][2];
                uint8_t dd = 0, ll = 0;
                if (spiht_ref_get_offspring(i, j, h, w, ll_h, ll_w, off)) {
                    for (int q = 0; q < 4; ++q) {
                        uint64_t ci = off[q][0], cj = off[q][1];
                        uint8_t x = plane1(absu(a[ci * w + cj]));
                        if (x > dd) dd = x;
                        if (ci < nh && cj < nw) { uint8_t y = d[ci * nw + cj]; if (y > ll) ll = y; }
                    }
                    if (ll > dd) dd = ll;
                }
                dpll[(k * ll_h + i) * ll_w + j] = dd;
                lpll[(k * ll_h + i) * ll_w + j] = ll;
            }
    }
    return 0;
}

typedef struct { uint32_t t, k, i, j, fp; } lis_t;
typedef struct { lis_t *d; size_t len, cap; } lisv_t;
typedef struct { int32_t *d; size_t len, cap; } i32v_t;

static int lis_push(lisv_t *v, lis_t e)
{
    if (v->len == v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 1024;
        lis_t *nd = (lis_t *)realloc(v->d, nc * sizeof(lis_t));
        if (!nd) return -1;
        v->d = nd; v->cap = nc;
    }
    v->d[v->len++] = e;
    return 0;
}
static int i32_push(i32v_t *v, int32_t e)
{
    if (v->len == v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 1024;
        int32_t *nd = (int32_t *)realloc(v->d, nc * sizeof(int32_t));
        if (!nd) return -1;
        v->d = nd; v->cap = nc;
    }
    v->d[v->len++] = e;
    return 0;
}

typedef struct { uint8_t *d; uint64_t nbits, capbytes, max_bits; int full; } mbits_t;
static void mb_push(mbits_t *b, int bit)
{
    if (b->full) return;
    if ((b->nbits >> 3) >= b->capbytes) {
        uint64_t nc = b->capbytes ? b->capbytes * 2 : 4096;
        b->d = (uint8_t *)realloc(b->d, nc);
        memset(b->d + b->capbytes, 0, nc - b->capbytes);
        b->capbytes = nc;
    }
    if (bit) b->d[b->nbits >> 3] |= (uint8_t)(1u << (b->nbits & 7));
    b->nbits++;
    if (b->nbits == b->max_bits) b->full = 1;
}

int spiht_model_encode(const int32_t *arr, uint64_t c, uint64_t h, uint64_t w,
                       uint64_t ll_h, uint64_t ll_w, uint64_t max_bits,
                       uint8_t **out_bytes, uint64_t *out_nbits, int *out_max_n)
{
    if (!(ll_h > 1) || !(ll_w > 1)) return MODEL_EBADARG;
    if (!spiht_model_geom_ok(h, w, ll_h, ll_w)) return MODEL_EGEOM;
    uint64_t nh = h / 2, nw = w / 2;
    uint8_t *dp = (uint8_t *)calloc(c * nh * nw + 1, 1), *lp = (uint8_t *)calloc(c * nh * nw + 1, 1);
    uint8_t *dpll = (uint8_t *)calloc(c * ll_h * ll_w, 1), *lpll = (uint8_t *)calloc(c * ll_h * ll_w, 1);
    spiht_model_pyramid(arr, c, h, w, ll_h, ll_w, dp, lp, dpll, lpll);

    uint32_t max = 0;
    for (uint64_t t = 0; t < c * h * w; ++t) { uint32_t a = absu(arr[t]); if (a > max) max = a; }
    int max_n = spiht_ref_max_n((int32_t)max);

    i32v_t lip = {0}, lsp = {0};
    lisv_t R = {0}, Rnew = {0}, G = {0}, Gn = {0};
    mbits_t out = {0, 0, 0, max_bits, 0};

    for (uint64_t i = 0; i < ll_h; ++i)
        for (uint64_t j = 0; j < ll_w; ++j)
            for (uint64_t k = 0; k < c; ++k)
                i32_push(&lip, arr[(k * h + i) * w + j]);
    for (uint64_t i = 0; i < ll_h; ++i)
        for (uint64_t j = 0; j < ll_w; ++j) {
            if (i % 2 == 0 && j % 2 == 0) continue;
            for (uint64_t k = 0; k < c; ++k) {
                lis_t e = {1, (uint32_t)k, (uint32_t)i, (uint32_t)j, dpll[(k * ll_h + i) * ll_w + j]};
                lis_push(&R, e);
            }
        }

    for (int n = max_n; n >= 0 && !out.full; --n) {
        uint32_t thr = 1u << n;
        size_t lsp_len = lsp.len;

        /* LIP pass: in-place stable compaction */
        size_t keep = 0;
        for (size_t q = 0; q < lip.len && !out.full; ++q) {
            int32_t x = lip.d[q];
            int sig = absu(x) >= thr;
            mb_push(&out, sig);
            if (sig) { i32_push(&lsp, (int32_t)absu(x)); mb_push(&out, x >= 0); }
            else lip.d[keep++] = x;
        }
        if (out.full) break;
        lip.len = keep;

        /* LIS pass by generations */
        Rnew.len = 0;
        lisv_t *cur = &R, *nxt = &G, *spare = &Gn;
        while (cur->len && !out.full) {
            nxt->len = 0;
            for (size_t q = 0; q < cur->len && !out.full; ++q) {
                lis_t e = cur->d[q];
                int fire = (int)e.fp - 1 >= n;
                mb_push(&out, fire);
                if (!fire) { lis_push(&Rnew, e); continue; }
                uint64_t off[4][2];
                int has = spiht_ref_get_offspring(e.i, e.j, h, w, ll_h, ll_w, off);
                int is_ll = e.i < ll_h && e.j < ll_w;
                if (e.t) {
                    /* fired => has offspring */
                    for (int r = 0; r < 4 && has; ++r) {
                        int32_t x = arr[(e.k * h + off[r][0]) * w + off[r][1]];
                        int sig = absu(x) >= thr;
                        mb_push(&out, sig);
                        if (sig) { i32_push(&lsp, (int32_t)absu(x)); mb_push(&out, x >= 0); }
                        else i32_push(&lip, x);
                    }
                    if (spiht_ref_has_descendents_past_offspring(e.i, e.j, h, w)) {
                        uint8_t f = is_ll ? lpll[(e.k * ll_h + e.i) * ll_w + e.j]
                                          : ((e.i < nh && e.j < nw) ? lp[(e.k * nh + e.i) * nw + e.j] : 0);
                        lis_t b = {0, e.k, e.i, e.j, f};
                        lis_push(nxt, b);
                    }
                } else {
                    for (int r = 0; r < 4 && has; ++r) {
                        uint64_t ci = off[r][0], cj = off[r][1];
                        uint8_t f = (ci < nh && cj < nw) ? dp[(e.k * nh + ci) * nw + cj] : 0;
                        lis_t a = {1, e.k, (uint32_t)ci, (uint32_t)cj, f};
                        lis_push(nxt, a);
                    }
                }
            }
            if (cur == &R) { cur = nxt; nxt = spare; }
            else { lisv_t *t = cur; cur = nxt; nxt = t; }
        }
        if (out.full) break;
        { lisv_t t = R; R = Rnew; Rnew = t; }

        /* refinement */
        for (size_t q = 0; q < lsp_len && !out.full; ++q)
            mb_push(&out, ((uint32_t)lsp.d[q] >> n) & 1);
    }

    free(dp); free(lp); free(dpll); free(lpll);
    free(lip.d); free(lsp.d); free(R.d); free(Rnew.d); free(G.d); free(Gn.d);
    if (!out.d) out.d = (uint8_t *)calloc(1, 1);
    *out_bytes = out.d;
    *out_nbits = out.nbits;
    *out_max_n = max_n;
    return MODEL_OK;
}
