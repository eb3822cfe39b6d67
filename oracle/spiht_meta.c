/*
 * oracle/spiht_meta.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's decode_with_metadata (theAdamColton/spiht,
 * src/encoder_decoder.rs:123-151 CoefficientMetadata, :457-462 Filter, :464-527 Slices::from_vec,
 * :593-613 get_local_position, :616-841 decode_with_metadata; src/lib.rs:47-56 binding): the decoder of
 * spiht_ref.c plus, for every bit, the decoder state just before that bit is read -- a row of 8 int32:
 *   [0] action id 0..6   [1] [2] position of the coefficient inside its band, scaled to -100000..100000
 *   [3] channel          [4] filter 0..3 (LL, DA, AD, DD)   [5] depth   [6] n   [7] value of the coefficient
 * Quirks kept: the table has 8*nbytes + 1 rows (the last one describes the bit that was never read); the
 * reference's bound check `cur_i >= metadata_arr.len()` compares against the element count, so it never fires;
 * positions are computed in float32 and converted with a saturating truncation; depth is a u8 that wraps;
 * band rectangles come from the caller in the order da, ad, dd (spiht_wrapper.py:240), coarsest level first.
 * Pinning: the reference's own tests only assert that the coefficient array equals the plain decoder's
 * (encoder_decoder.rs:929-966, spiht/tests/test_spiht.py:19-28); the table itself is unpinned by the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SPIHT_REF_OK 0
#define SPIHT_REF_EBADARG 1
#define SPIHT_REF_EPANIC 2
#define SPIHT_REF_ENOMEM 3

int spiht_ref_get_offspring(uint64_t i, uint64_t j, uint64_t h, uint64_t w, uint64_t ll_h, uint64_t ll_w,
                            uint64_t out[4][2]);
int spiht_ref_has_descendents_past_offspring(uint64_t i, uint64_t j, uint64_t h, uint64_t w);
int32_t spiht_ref_set_bit(int32_t x, unsigned n, int bit);

typedef struct { uint8_t t, depth, filter; uint32_t k, i, j; } ment_t;
typedef struct { ment_t *d; size_t head, len, cap; } mfifo_t;

static int mpush(mfifo_t *f, ment_t e)
{
    if (f->head + f->len == f->cap) {
        if (f->head > 0 && f->head >= f->len) {
            memmove(f->d, f->d + f->head, f->len * sizeof(ment_t));
            f->head = 0;
        } else {
            size_t nc = f->cap ? f->cap * 2 : 1024;
            ment_t *nd = (ment_t *)realloc(f->d, nc * sizeof(ment_t));
            if (!nd) return -1;
            f->d = nd; f->cap = nc;
        }
    }
    f->d[f->head + f->len++] = e;
    return 0;
}
static inline int mpop(mfifo_t *f, ment_t *e)
{
    if (!f->len) return 0;
    *e = f->d[f->head++]; f->len--;
    return 1;
}

/* `as i32` of an f32 (Rust: saturating, NaN -> 0) */
static int32_t f32_as_i32(float v)
{
    if (isnan(v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}

/* encoder_decoder.rs:134-149 */
static uint8_t offspring_filter(const ment_t *c)
{
    if (c->filter == 0) {
        if (c->i % 2 == 1 && c->j % 2 == 1) return 3;       /* DD */
        if (c->i % 2 == 0 && c->j % 2 != 0) return 2;       /* AD */
        return 1;                                            /* DA */
    }
    return c->filter;
}

/* top: {start_i, end_i, start_j, end_j}; other: [level][3][4] in the caller's order (da, ad, dd) */
int spiht_ref_decode_with_metadata(const uint8_t *data, uint64_t nbytes, unsigned n, uint64_t c, uint64_t h, uint64_t w,
                                   uint64_t ll_h, uint64_t ll_w, const int64_t *top, const int64_t *other,
                                   uint64_t nlevels, int32_t *out, int32_t *meta)
{
    if (!(ll_h > 1) || !(ll_w > 1)) return SPIHT_REF_EBADARG;
    const uint64_t nbits = nbytes * 8;
    memset(out, 0, (size_t)(c * h * w) * sizeof(int32_t));
    memset(meta, 0, (size_t)(nbits + 1) * 8 * sizeof(int32_t));
    const uint8_t level = (uint8_t)nlevels;
    uint64_t cur = 0;
    int rc = SPIHT_REF_OK;
    mfifo_t lip = {0}, lis = {0}, lsp = {0}, lip_retain = {0}, lis_retain = {0};

#define REC(k, i, j) out[((uint64_t)(k) * h + (i)) * w + (j)]
#define CHECK_IDX(k, i, j) do { if ((k) >= c || (i) >= h || (j) >= w) { rc = SPIHT_REF_EPANIC; goto done; } } while (0)
#define POP_BIT(dst)                                              \
    do {                                                          \
        if (cur >= nbits) goto done;                              \
        (dst) = (data[cur >> 3] >> (cur & 7)) & 1;                \
        cur++;                                                    \
    } while (0)
    /* encoder_decoder.rs:593-613 + :663-682 */
#define ASSIGN(action, co)                                                                            \
    do {                                                                                              \
        float lh, lw;                                                                                 \
        if ((co).depth == level) {                                                                    \
            lh = (float)(co).i / (float)top[1];                                                       \
            lw = (float)(co).j / (float)top[3];                                                       \
        } else {                                                                                      \
            const uint8_t depth_i = (uint8_t)(level - 1 - (co).depth);                                \
            const unsigned filter_i = (unsigned)(co).filter - 1u;                                     \
            if (depth_i >= nlevels || filter_i > 2u) { rc = SPIHT_REF_EPANIC; goto done; }            \
            const int64_t *sl = other + ((size_t)depth_i * 3 + filter_i) * 4;                         \
            lh = ((float)(co).i - (float)sl[0]) / (float)(uint64_t)(sl[1] - sl[0]);                   \
            lw = ((float)(co).j - (float)sl[2]) / (float)(uint64_t)(sl[3] - sl[2]);                   \
        }                                                                                             \
        CHECK_IDX((co).k, (co).i, (co).j);                                                            \
        int32_t *row = meta + cur * 8;                                                                \
        row[0] = (action);                                                                            \
        row[1] = f32_as_i32(lh * 200000.0f - 100000.0f);                                              \
        row[2] = f32_as_i32(lw * 200000.0f - 100000.0f);                                              \
        row[3] = (int32_t)(co).k;                                                                     \
        row[4] = (co).filter;                                                                         \
        row[5] = (co).depth;                                                                          \
        row[6] = (int32_t)n;                                                                          \
        row[7] = REC((co).k, (co).i, (co).j);                                                         \
    } while (0)

    for (uint64_t i = 0; i < ll_h; ++i)
        for (uint64_t j = 0; j < ll_w; ++j)
            for (uint64_t k = 0; k < c; ++k) {
                ment_t e = {1, level, 0, (uint32_t)k, (uint32_t)i, (uint32_t)j};
                mpush(&lip, e);
            }
    for (uint64_t i = 0; i < ll_h; ++i)
        for (uint64_t j = 0; j < ll_w; ++j) {
            if (i % 2 == 0 && j % 2 == 0) continue;
            for (uint64_t k = 0; k < c; ++k) {
                ment_t e = {1, level, 0, (uint32_t)k, (uint32_t)i, (uint32_t)j};
                mpush(&lis, e);
            }
        }

    for (;;) {
        const size_t lsp_len = lsp.len;
        ment_t e;
        int bit;
        lip_retain.head = lip_retain.len = 0;
        while (mpop(&lip, &e)) {
            ASSIGN(0, e);
            POP_BIT(bit);
            if (bit) {
                ASSIGN(1, e);
                POP_BIT(bit);
                const int32_t sign = bit * 2 - 1;
                const int32_t base = n == 0 ? 1 : (int32_t)((1u << (n - 1)) + (1u << n));
                REC(e.k, e.i, e.j) = base * sign;
                mpush(&lsp, e);
            } else {
                mpush(&lip_retain, e);
            }
        }
        { mfifo_t t = lip; lip = lip_retain; lip_retain = t; }

        lis_retain.head = lis_retain.len = 0;
        while (mpop(&lis, &e)) {
            uint64_t off[4][2];
            if (e.t) {
                ASSIGN(2, e);
                POP_BIT(bit);
                if (bit) {
                    if (spiht_ref_get_offspring(e.i, e.j, h, w, ll_h, ll_w, off)) {
                        for (int q = 0; q < 4; ++q) {
                            ment_t ch = {1, (uint8_t)(e.depth - 1), offspring_filter(&e), e.k, (uint32_t)off[q][0],
                                         (uint32_t)off[q][1]};
                            ASSIGN(3, ch);
                            POP_BIT(bit);
                            if (bit) {
                                ASSIGN(4, ch);
                                POP_BIT(bit);
                                const int32_t sign = bit * 2 - 1;
                                const int32_t base = n == 0 ? 1 : (int32_t)((1u << (n - 1)) + (1u << n));
                                REC(ch.k, ch.i, ch.j) = sign * base;
                                mpush(&lsp, ch);
                            } else {
                                mpush(&lip, ch);
                            }
                        }
                    }
                    if (spiht_ref_has_descendents_past_offspring(e.i, e.j, h, w)) {
                        ment_t b = e;
                        b.t = 0;
                        mpush(&lis, b);
                    }
                } else {
                    mpush(&lis_retain, e);
                }
            } else {
                ASSIGN(5, e);
                POP_BIT(bit);
                if (bit) {
                    if (spiht_ref_get_offspring(e.i, e.j, h, w, ll_h, ll_w, off)) {
                        for (int q = 0; q < 4; ++q) {
                            ment_t a = {1, (uint8_t)(e.depth - 1), offspring_filter(&e), e.k, (uint32_t)off[q][0],
                                        (uint32_t)off[q][1]};
                            mpush(&lis, a);
                        }
                    }
                } else {
                    mpush(&lis_retain, e);
                }
            }
        }
        { mfifo_t t = lis; lis = lis_retain; lis_retain = t; }

        for (size_t q = 0; q < lsp_len; ++q) {
            ment_t s = lsp.d[lsp.head + q];
            ASSIGN(6, s);
            POP_BIT(bit);
            REC(s.k, s.i, s.j) = spiht_ref_set_bit(REC(s.k, s.i, s.j), n, bit);
        }
        if (n == 0) break;
        n -= 1;
    }
done:
    free(lip.d); free(lis.d); free(lsp.d); free(lip_retain.d); free(lis_retain.d);
    return rc;
#undef ASSIGN
#undef POP_BIT
#undef REC
#undef CHECK_IDX
}
