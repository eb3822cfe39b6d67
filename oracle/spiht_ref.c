/*
 * oracle/spiht_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference SPIHT bit-plane coder
 * (theAdamColton/spiht, src/encoder_decoder.rs + src/lib.rs), written from the
 * algorithm's behaviour, in plain C.  It is the checker the CUDA path is
 * compared against.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product never does.
 *
 * The reference is Rust (pyo3); no Rust toolchain exists in this image, so the
 * reference itself cannot be compiled (no oracle/_ref).  Pinning: this file is
 * checked against every result the reference's own tests hold for this path --
 * the helper known-answer tests (encoder_decoder.rs:851-862, 988-1024), the
 * `max_n == 5` test (864-875) and the lossless round-trip tests (877-985) --
 * see tests/test_oracle_spiht.py.  The reference holds no stored bitstream, so
 * bitstream BYTES are pinned only by hand-traced vectors (SURVEY.md section 4,
 * KAT-1/2/3).
 *
 * Two coders live here:
 *   spiht_ref_encode / spiht_ref_decode    "faithful": same control flow as the
 *       reference (recursive subtree scans, FIFO lists, one bit at a time).
 *       This is what the CPU baseline times.
 *   (oracle/spiht_model.c holds the scan/pyramid formulation the GPU uses and
 *    is checked bit-for-bit against the faithful coder.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SPIHT_REF_OK 0
#define SPIHT_REF_EBADARG 1   /* reference: assert!(ll_h > 1), assert!(ll_w > 1) */
#define SPIHT_REF_EPANIC 2    /* reference would panic on an out-of-bounds index */
#define SPIHT_REF_ENOMEM 3

/* ---------- scalar helpers (encoder_decoder.rs:7-41) ---------- */

/* encoder_decoder.rs:7-12 */
int spiht_ref_has_descendents_past_offspring(uint64_t i, uint64_t j, uint64_t h, uint64_t w)
{
    if ((i * 2 + 1) * 2 + 1 >= h || (j * 2 + 1) * 2 + 1 >= w) return 0;
    return 1;
}

/* encoder_decoder.rs:14-29 -- set/clear bit n of the magnitude, keep the sign */
int32_t spiht_ref_set_bit(int32_t x, unsigned n, int bit)
{
    int sign = x >= 0;
    int32_t m = (int32_t)(1u << n);
    if (bit) return sign ? (x | m) : -((-x) | m);
    return sign ? (x & ~m) : -((-x) & ~m);
}

/* encoder_decoder.rs:31-34 */
int spiht_ref_is_bit_set(int32_t x, unsigned n)
{
    int32_t a = x < 0 ? -x : x;
    return (a & (int32_t)(1u << n)) != 0;
}

/* encoder_decoder.rs:37-41 */
int spiht_ref_is_element_sig(int32_t x, unsigned n)
{
    int32_t a = x < 0 ? -x : x;
    return a >= (int32_t)(1u << n);
}

/* encoder_decoder.rs:43-75.  Returns 1 and fills out[4][2] when the node has
 * offspring, 0 otherwise.  No bounds check for LL roots, as in the reference. */
int spiht_ref_get_offspring(uint64_t i, uint64_t j, uint64_t h, uint64_t w,
                            uint64_t ll_h, uint64_t ll_w, uint64_t out[4][2])
{
    if (i < ll_h && j < ll_w) {
        if (i % 2 == 0 && j % 2 == 0) return 0;
        uint64_t si = i / 2 * 2, sj = j / 2 * 2;
        uint64_t ci = i % 2, cj = j % 2;
        uint64_t bi = ci * ll_h + si, bj = cj * ll_w + sj;
        out[0][0] = bi;     out[0][1] = bj;
        out[1][0] = bi;     out[1][1] = bj + 1;
        out[2][0] = bi + 1; out[2][1] = bj;
        out[3][0] = bi + 1; out[3][1] = bj + 1;
        return 1;
    }
    if (2 * i + 1 >= h || 2 * j + 1 >= w) return 0;
    out[0][0] = 2 * i;     out[0][1] = 2 * j;
    out[1][0] = 2 * i;     out[1][1] = 2 * j + 1;
    out[2][0] = 2 * i + 1; out[2][1] = 2 * j;
    out[3][0] = 2 * i + 1; out[3][1] = 2 * j + 1;
    return 1;
}

/* (max as f32).log2() as u8 -- encoder_decoder.rs:166.  Saturating cast:
 * max == 0 -> log2 = -inf -> 0.  Uses the host libm log2f, which is what the
 * Rust intrinsic lowers to on Linux. */
int spiht_ref_max_n(int32_t max_abs)
{
    float f = log2f((float)max_abs);
    if (!(f > 0.0f)) return 0;
    if (f >= 255.0f) return 255;
    return (int)f;
}

/* ---------- array view with a "would panic" flag ---------- */
typedef struct {
    const int32_t *a;
    uint64_t c, h, w;
    int panic;
} view_t;

static inline int32_t at(view_t *v, uint64_t k, uint64_t i, uint64_t j)
{
    if (k >= v->c || i >= v->h || j >= v->w) { v->panic = 1; return 0; }
    return v->a[(k * v->h + i) * v->w + j];
}

/* encoder_decoder.rs:78-99 */
static int is_set_sig(view_t *v, uint64_t k, uint64_t i, uint64_t j, unsigned n,
                      uint64_t ll_h, uint64_t ll_w)
{
    if (spiht_ref_is_element_sig(at(v, k, i, j), n)) return 1;
    if (v->panic) return 0;
    uint64_t off[4][2];
    if (spiht_ref_get_offspring(i, j, v->h, v->w, ll_h, ll_w, off)) {
        for (int q = 0; q < 4; ++q)
            if (is_set_sig(v, k, off[q][0], off[q][1], n, ll_h, ll_w)) return 1;
    }
    return 0;
}

/* encoder_decoder.rs:101-121 */
static int is_l_sig(view_t *v, uint64_t k, uint64_t i, uint64_t j, unsigned n,
                    uint64_t ll_h, uint64_t ll_w)
{
    uint64_t off[4][2], off2[4][2];
    if (spiht_ref_get_offspring(i, j, v->h, v->w, ll_h, ll_w, off)) {
        for (int q = 0; q < 4; ++q) {
            if (spiht_ref_get_offspring(off[q][0], off[q][1], v->h, v->w, ll_h, ll_w, off2)) {
                for (int r = 0; r < 4; ++r)
                    if (is_set_sig(v, k, off2[r][0], off2[r][1], n, ll_h, ll_w)) return 1;
            }
        }
    }
    return 0;
}

/* ---------- FIFO lists ---------- */
typedef struct { uint32_t t, k, i, j; } ent_t;
typedef struct { ent_t *d; size_t head, len, cap; } fifo_t;

static int fifo_push(fifo_t *f, ent_t e)
{
    if (f->head + f->len == f->cap) {
        if (f->head > 0 && f->head >= f->len) { /* compact */
            memmove(f->d, f->d + f->head, f->len * sizeof(ent_t));
            f->head = 0;
        } else {
            size_t nc = f->cap ? f->cap * 2 : 1024;
            ent_t *nd = (ent_t *)realloc(f->d, nc * sizeof(ent_t));
            if (!nd) return -1;
            f->d = nd; f->cap = nc;
        }
    }
    f->d[f->head + f->len++] = e;
    return 0;
}
static inline int fifo_pop(fifo_t *f, ent_t *e)
{
    if (!f->len) return 0;
    *e = f->d[f->head++]; f->len--;
    return 1;
}
static void fifo_free(fifo_t *f) { free(f->d); memset(f, 0, sizeof(*f)); }

/* ---------- growable bit vector, LSB-first bytes (lib.rs:29) ---------- */
typedef struct { uint8_t *d; uint64_t nbits, capbytes; } bits_t;
static int bits_push(bits_t *b, int bit)
{
    if ((b->nbits >> 3) >= b->capbytes) {
        uint64_t nc = b->capbytes ? b->capbytes * 2 : 4096;
        uint8_t *nd = (uint8_t *)realloc(b->d, nc);
        if (!nd) return -1;
        memset(nd + b->capbytes, 0, nc - b->capbytes);
        b->d = nd; b->capbytes = nc;
    }
    if (bit) b->d[b->nbits >> 3] |= (uint8_t)(1u << (b->nbits & 7));
    b->nbits++;
    return 0;
}

static void init_lists(fifo_t *lip, fifo_t *lis, uint64_t c, uint64_t ll_h, uint64_t ll_w)
{
    /* encoder_decoder.rs:170-190 (and 329-348): i, then j, channel innermost */
    for (uint64_t i = 0; i < ll_h; ++i)
        for (uint64_t j = 0; j < ll_w; ++j)
            for (uint64_t k = 0; k < c; ++k) {
                ent_t e = {1, (uint32_t)k, (uint32_t)i, (uint32_t)j};
                fifo_push(lip, e);
            }
    for (uint64_t i = 0; i < ll_h; ++i)
        for (uint64_t j = 0; j < ll_w; ++j) {
            if (i % 2 == 0 && j % 2 == 0) continue;
            for (uint64_t k = 0; k < c; ++k) {
                ent_t e = {1, (uint32_t)k, (uint32_t)i, (uint32_t)j};
                fifo_push(lis, e);
            }
        }
}

void spiht_ref_free(void *p) { free(p); }

/* encoder_decoder.rs:155-303.  arr is C-contiguous int32 [c][h][w].
 * *out_bytes is malloc'ed (free with spiht_ref_free), ceil(nbits/8) bytes. */
int spiht_ref_encode(const int32_t *arr, uint64_t c, uint64_t h, uint64_t w,
                     uint64_t ll_h, uint64_t ll_w, uint64_t max_bits,
                     uint8_t **out_bytes, uint64_t *out_nbits, int *out_max_n)
{
    if (!(ll_h > 1) || !(ll_w > 1)) return SPIHT_REF_EBADARG;
    if (c == 0 || h == 0 || w == 0) return SPIHT_REF_EPANIC; /* .max().unwrap() on empty */
    view_t v = {arr, c, h, w, 0};
    bits_t data = {0, 0, 0};
    fifo_t lip = {0}, lis = {0}, lsp = {0}, lip_retain = {0}, lis_retain = {0};
    int rc = SPIHT_REF_OK;

    int32_t max = 0;
    for (uint64_t t = 0; t < c * h * w; ++t) {
        int32_t a = arr[t] < 0 ? -arr[t] : arr[t];
        if (a > max) max = a;
    }
    unsigned n = (unsigned)spiht_ref_max_n(max);
    int max_n = (int)n;

    init_lists(&lip, &lis, c, ll_h, ll_w);

#define PUSH_BIT(b)                                                     \
    do {                                                                \
        if (bits_push(&data, (b))) { rc = SPIHT_REF_ENOMEM; goto done; } \
        if (data.nbits == max_bits) goto done;                          \
    } while (0)
#define CHECK_PANIC() do { if (v.panic) { rc = SPIHT_REF_EPANIC; goto done; } } while (0)

    for (;;) {
        size_t lsp_len = lsp.len;
        ent_t e;

        /* LIP pass, 207-222 */
        lip_retain.head = lip_retain.len = 0;
        while (fifo_pop(&lip, &e)) {
            int32_t x = at(&v, e.k, e.i, e.j);
            CHECK_PANIC();
            int sig = spiht_ref_is_element_sig(x, n);
            PUSH_BIT(sig);
            if (sig) {
                fifo_push(&lsp, e);
                PUSH_BIT(x >= 0);
            } else {
                fifo_push(&lip_retain, e);
            }
        }
        { fifo_t t = lip; lip = lip_retain; lip_retain = t; }

        /* LIS pass, 224-284 */
        lis_retain.head = lis_retain.len = 0;
        while (fifo_pop(&lis, &e)) {
            uint64_t off[4][2];
            if (e.t) { /* type A */
                int desc_sig = 0;
                int has = spiht_ref_get_offspring(e.i, e.j, h, w, ll_h, ll_w, off);
                if (has) {
                    for (int q = 0; q < 4; ++q) {
                        if (is_set_sig(&v, e.k, off[q][0], off[q][1], n, ll_h, ll_w)) { desc_sig = 1; break; }
                        CHECK_PANIC();
                    }
                }
                PUSH_BIT(desc_sig);
                if (desc_sig) {
                    for (int q = 0; q < 4; ++q) {
                        int32_t x = at(&v, e.k, off[q][0], off[q][1]);
                        CHECK_PANIC();
                        int sig = spiht_ref_is_element_sig(x, n);
                        PUSH_BIT(sig);
                        ent_t ch = {1, e.k, (uint32_t)off[q][0], (uint32_t)off[q][1]};
                        if (sig) {
                            fifo_push(&lsp, ch);
                            PUSH_BIT(x >= 0);
                        } else {
                            fifo_push(&lip, ch);
                        }
                    }
                    if (spiht_ref_has_descendents_past_offspring(e.i, e.j, h, w)) {
                        ent_t b = {0, e.k, e.i, e.j};
                        fifo_push(&lis, b);
                    }
                } else {
                    fifo_push(&lis_retain, e);
                }
            } else { /* type B */
                int l_sig = is_l_sig(&v, e.k, e.i, e.j, n, ll_h, ll_w);
                CHECK_PANIC();
                PUSH_BIT(l_sig);
                if (l_sig) {
                    if (spiht_ref_get_offspring(e.i, e.j, h, w, ll_h, ll_w, off)) {
                        for (int q = 0; q < 4; ++q) {
                            ent_t a = {1, e.k, (uint32_t)off[q][0], (uint32_t)off[q][1]};
                            fifo_push(&lis, a);
                        }
                    }
                } else {
                    fifo_push(&lis_retain, e);
                }
            }
        }
        { fifo_t t = lis; lis = lis_retain; lis_retain = t; }

        /* refinement, 287-292 */
        for (size_t q = 0; q < lsp_len; ++q) {
            ent_t s = lsp.d[lsp.head + q];
            PUSH_BIT(spiht_ref_is_bit_set(at(&v, s.k, s.i, s.j), n));
        }

        if (n == 0) break;
        n -= 1;
    }
done:
    if (v.panic && rc == SPIHT_REF_OK) rc = SPIHT_REF_EPANIC;
    fifo_free(&lip); fifo_free(&lis); fifo_free(&lsp);
    fifo_free(&lip_retain); fifo_free(&lis_retain);
    if (rc != SPIHT_REF_OK) { free(data.d); return rc; }
    if (!data.d) data.d = (uint8_t *)calloc(1, 1);
    *out_bytes = data.d;
    *out_nbits = data.nbits;
    *out_max_n = max_n;
    return SPIHT_REF_OK;
#undef PUSH_BIT
#undef CHECK_PANIC
}

/* encoder_decoder.rs:307-454 with the byte->bit expansion of lib.rs:15-21:
 * the decoder sees 8*nbytes bits (pad bits are decoded as data).
 * out is int32 [c][h][w], zero-filled here. */
int spiht_ref_decode(const uint8_t *data, uint64_t nbytes, unsigned n,
                     uint64_t c, uint64_t h, uint64_t w, uint64_t ll_h, uint64_t ll_w,
                     int32_t *out)
{
    if (!(ll_h > 1) || !(ll_w > 1)) return SPIHT_REF_EBADARG;
    memset(out, 0, (size_t)(c * h * w) * sizeof(int32_t));
    uint64_t nbits = nbytes * 8, cur = 0;
    fifo_t lip = {0}, lis = {0}, lsp = {0}, lip_retain = {0}, lis_retain = {0};
    int rc = SPIHT_REF_OK;
    init_lists(&lip, &lis, c, ll_h, ll_w);

#define POP_BIT(dst)                                              \
    do {                                                          \
        if (cur >= nbits) goto done;                              \
        (dst) = (data[cur >> 3] >> (cur & 7)) & 1;                \
        cur++;                                                    \
    } while (0)
#define REC(k, i, j) out[((uint64_t)(k) * h + (i)) * w + (j)]
#define CHECK_IDX(k, i, j) do { if ((k) >= c || (i) >= h || (j) >= w) { rc = SPIHT_REF_EPANIC; goto done; } } while (0)

    for (;;) {
        size_t lsp_len = lsp.len;
        ent_t e;
        int bit;

        lip_retain.head = lip_retain.len = 0;
        while (fifo_pop(&lip, &e)) {
            POP_BIT(bit);
            if (bit) {
                fifo_push(&lsp, e);
                POP_BIT(bit);
                int32_t sign = bit * 2 - 1;
                int32_t base = n == 0 ? 1 : (int32_t)((1u << (n - 1)) + (1u << n));
                CHECK_IDX(e.k, e.i, e.j);
                REC(e.k, e.i, e.j) = base * sign;
            } else {
                fifo_push(&lip_retain, e);
            }
        }
        { fifo_t t = lip; lip = lip_retain; lip_retain = t; }

        lis_retain.head = lis_retain.len = 0;
        while (fifo_pop(&lis, &e)) {
            uint64_t off[4][2];
            if (e.t) {
                POP_BIT(bit);
                if (bit) {
                    if (spiht_ref_get_offspring(e.i, e.j, h, w, ll_h, ll_w, off)) {
                        for (int q = 0; q < 4; ++q) {
                            ent_t ch = {1, e.k, (uint32_t)off[q][0], (uint32_t)off[q][1]};
                            POP_BIT(bit);
                            if (bit) {
                                fifo_push(&lsp, ch);
                                POP_BIT(bit);
                                int32_t sign = bit * 2 - 1;
                                int32_t base = n == 0 ? 1 : (int32_t)((1u << (n - 1)) + (1u << n));
                                CHECK_IDX(ch.k, ch.i, ch.j);
                                REC(ch.k, ch.i, ch.j) = sign * base;
                            } else {
                                fifo_push(&lip, ch);
                            }
                        }
                    }
                    if (spiht_ref_has_descendents_past_offspring(e.i, e.j, h, w)) {
                        ent_t b = {0, e.k, e.i, e.j};
                        fifo_push(&lis, b);
                    }
                } else {
                    fifo_push(&lis_retain, e);
                }
            } else {
                POP_BIT(bit);
                if (bit) {
                    if (spiht_ref_get_offspring(e.i, e.j, h, w, ll_h, ll_w, off)) {
                        for (int q = 0; q < 4; ++q) {
                            ent_t a = {1, e.k, (uint32_t)off[q][0], (uint32_t)off[q][1]};
                            fifo_push(&lis, a);
                        }
                    }
                } else {
                    fifo_push(&lis_retain, e);
                }
            }
        }
        { fifo_t t = lis; lis = lis_retain; lis_retain = t; }

        for (size_t q = 0; q < lsp_len; ++q) {
            ent_t s = lsp.d[lsp.head + q];
            POP_BIT(bit);
            CHECK_IDX(s.k, s.i, s.j);
            REC(s.k, s.i, s.j) = spiht_ref_set_bit(REC(s.k, s.i, s.j), n, bit);
        }

        if (n == 0) break;
        n -= 1;
    }
done:
    fifo_free(&lip); fifo_free(&lis); fifo_free(&lsp);
    fifo_free(&lip_retain); fifo_free(&lis_retain);
    return rc;
#undef POP_BIT
#undef REC
#undef CHECK_IDX
}
