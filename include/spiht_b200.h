/*
 * spiht_b200.h -- C ABI of libspiht_b200.so, the B200 (sm_100a) SPIHT hot path.
 *
 * Drop-in boundary for theAdamColton/spiht.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repo).
 * Plain pointers and sizes only; no torch / numpy types.  All functions return
 * SPIHTB_OK (0) or an error code; spihtb_last_error() gives the message of the
 * last failure on the calling thread.  There is no CPU fallback: every compute
 * entry point fails with SPIHTB_ECUDA when no CUDA device is usable.
 *
 * Conventions
 *   - "host" pointers are ordinary CPU memory, "dev" pointers are CUDA device
 *     memory on the context's device (e.g. torch tensors' data_ptr()).
 *   - coefficient arrays are C-contiguous int32 [B][C][enc_h][enc_w];
 *     images are C-contiguous planar [B][C][H][W], float32 or float64.
 *   - a bitstream is packed LSB-first: stream bit t is bit (t % 8) of byte t/8
 *     (src/lib.rs:29 `chunks(8).map(load_le::<u8>)`).
 *   - calls on one context are ordered on its CUDA stream; device-pointer
 *     entry points are asynchronous, host-pointer entry points synchronise.
 */
#ifndef SPIHT_B200_H
#define SPIHT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPIHTB_VERSION 100 /* 0.1.0 */

/* error codes */
#define SPIHTB_OK 0
#define SPIHTB_EINVAL 1 /* bad argument (null pointer, non-positive size, unknown id)   -> ValueError */
#define SPIHTB_ELL 2    /* ll_h <= 1 || ll_w <= 1: reference `assert!(ll_h > 1)` encoder_decoder.rs:160-161,310-311 */
#define SPIHTB_EGEOM 3  /* an LL-root offspring (encoder_decoder.rs:50-62) falls outside the array: reference panics */
#define SPIHTB_ESHAPE 4 /* bits(c)+bits(h)+bits(w) > 31: shape too large for the packed list entry */
#define SPIHTB_ECAP 5   /* output capacity too small */
#define SPIHTB_ECUDA 6  /* CUDA runtime error / no device */
#define SPIHTB_ENOMEM 7
#define SPIHTB_ELEVEL 8 /* level < 0, or the decomposition leaves no detail band (level 0) */

/* wavelet ids (PyWavelets names, spiht_wrapper.py:55 `wavelet`) */
#define SPIHTB_WAVELET_BIOR22 0
#define SPIHTB_WAVELET_BIOR44 1
#define SPIHTB_WAVELET_BIOR68 2
/* the rest of PyWavelets' bior family (spline pairs derived on the host; separable kernels with the taps as launch parameters) */
#define SPIHTB_WAVELET_BIOR11 3
#define SPIHTB_WAVELET_BIOR13 4
#define SPIHTB_WAVELET_BIOR15 5
#define SPIHTB_WAVELET_BIOR24 6
#define SPIHTB_WAVELET_BIOR26 7
#define SPIHTB_WAVELET_BIOR28 8
#define SPIHTB_WAVELET_BIOR31 9
#define SPIHTB_WAVELET_BIOR33 10
#define SPIHTB_WAVELET_BIOR35 11
#define SPIHTB_WAVELET_BIOR37 12
#define SPIHTB_WAVELET_BIOR39 13
#define SPIHTB_WAVELET_BIOR55 14   /* not a spline pair: stored table */
/* boundary mode ids (PyWavelets names, spiht_wrapper.py:57 `mode`) */
#define SPIHTB_MODE_REFLECT 0
#define SPIHTB_MODE_SYMMETRIC 1
#define SPIHTB_MODE_PERIODIZATION 2
/* colour model ids (spiht_wrapper.py:58 `color_model`) */
#define SPIHTB_COLOR_NONE 0
#define SPIHTB_COLOR_IPT 1
/* pixel dtypes */
#define SPIHTB_F32 0
#define SPIHTB_F64 1
/* forward direction only: uint8 pixels as stored on disk, scaled like the reference's loader
 * (spiht/utils.py:12-20 imload: im / 255 in float64) before the transform */
#define SPIHTB_U8 2

#define SPIHTB_MAX_LEVELS 24

typedef struct spihtb_ctx spihtb_ctx;

/* Band geometry of an L-level 2-D DWT laid out as pywt.coeffs_to_array does.
 * Replaces pywt.wavedecn_shapes + get_slices_and_h_w (spiht_wrapper.py:92-139). */
typedef struct spihtb_geom {
    int32_t h, w;            /* image size */
    int32_t wavelet, mode;   /* ids above */
    int32_t levels;          /* resolved number of levels (>= 1) */
    int32_t enc_h, enc_w;    /* coefficient array size (Hc, Wc) */
    int32_t ll_h, ll_w;      /* coarsest approximation band */
    int32_t rec_h, rec_w;    /* size of the image waverec2 returns (>= h, w) */
    /* per level, index 0 = finest (level 1): input size, band size, and the
     * offset (sh, sw) of that level's detail blocks in the coefficient array:
     * 'ad' at rows [0,bh) cols [sw,sw+bw); 'da' at rows [sh,sh+bh) cols [0,bw);
     * 'dd' at rows [sh,sh+bh) cols [sw,sw+bw). */
    int32_t in_h[SPIHTB_MAX_LEVELS], in_w[SPIHTB_MAX_LEVELS];
    int32_t band_h[SPIHTB_MAX_LEVELS], band_w[SPIHTB_MAX_LEVELS];
    int32_t off_h[SPIHTB_MAX_LEVELS], off_w[SPIHTB_MAX_LEVELS];
} spihtb_geom;

/* ---- library / context ------------------------------------------------- */
int spihtb_version(void);
const char *spihtb_last_error(void);
int spihtb_create(int device, spihtb_ctx **ctx);
int spihtb_destroy(spihtb_ctx *ctx);
/* run subsequent calls on `cuda_stream` (a cudaStream_t; NULL = legacy default stream).  A context owns one
 * set of workspaces, so its calls are serialised: when the stream changes, the new stream first waits (event)
 * for everything this context queued on the old one.  For concurrent calls use one context per stream. */
int spihtb_set_stream(spihtb_ctx *ctx, void *cuda_stream);
int spihtb_sync(spihtb_ctx *ctx);
/* number of kernels this context has launched so far (bench.py `gpu_launches`) */
int64_t spihtb_launch_count(spihtb_ctx *ctx);

/* Context options.
 * SPIHTB_OPT_SCRATCH_COEFFS (default 0): when non-zero, the caller promises not to read the coefficient array
 * spihtb_decode_images fills (`dev_coeffs_scratch`): the library then zeroes the finest detail bands of an image --
 * three quarters of the array -- only if the image's stream reaches them, and leaves them undefined otherwise (the
 * inverse transform does not read them).  The pixels are identical either way. */
#define SPIHTB_OPT_SCRATCH_COEFFS 1
int spihtb_set_option(spihtb_ctx *ctx, int32_t option, int64_t value);

/* which forward-transform path the last spihtb_forward / spihtb_encode_images call on this context took:
 * 12 = levels 1 and 2 fused in one kernel (TMA-staged tiles, dwt_fwd2.cu), 1 = one kernel per level.  bench.py
 * uses it to credit the dominant kernel with the right algorithmic bytes. */
int spihtb_forward_path(spihtb_ctx *ctx);

/* ---- stage timers (CUDA events on the context's stream; used by bench.py for the roofline) ----
 * Stages: 0 forward DWT level 1 (levels 1+2 when spihtb_forward_path() == 12), 1 forward remaining levels (+ colour, gap fill), 2 pyramid base pass (with
 * spihtb_encode_images: zero fill + fix-up of the cells the fused epilogue leaves open), 3 pyramid upper rings + LL roots, 4 SPIHT encode kernel, 5 SPIHT decode (zero fill + kernel),
 * 6 inverse DWT coarse levels, 7 inverse DWT finest level (+ colour). */
#define SPIHTB_NSTAGES 8
int spihtb_profile_enable(spihtb_ctx *ctx, int enable);
/* total milliseconds and number of recorded intervals per stage since the last reset; synchronises */
int spihtb_profile_read(spihtb_ctx *ctx, double *ms_out, int64_t *count_out, int reset);

/* ---- geometry (host only; replaces spiht_wrapper.py:92-139, pywt.wavedecn_shapes,
 *      and pywt's level=None -> dwtn_max_level) ----------------------------- */
/* level < 0 means "max level" (Python level=None). */
int spihtb_plan(int32_t h, int32_t w, int32_t wavelet, int32_t mode, int32_t level, spihtb_geom *out);
/* The filter bank the kernels use for a wavelet id, in PyWavelets' layout (pywt.Wavelet(name).dec_lo / .rec_lo, zero
 * padding included; dec_hi[i] = (-1)^(F-1-i) rec_lo[i], rec_hi[i] = (-1)^i dec_lo[i]).  Host only.  *flen = F (even,
 * <= 20); dec_lo and rec_lo receive F values each (room for 20). */
int spihtb_wavelet_filters(int32_t wavelet, int32_t *flen, double *dec_lo, double *rec_lo);

/* ---- raw SPIHT coder, HOST buffers: replaces the pyo3 module ------------ */
/* src/lib.rs:24-32  encode(x: int32[c,h,w], ll_h, ll_w, max_bits) -> (bytes, max_n)
 * -> src/encoder_decoder.rs:155-303.  max_bits == 0 never truncates (as in the
 * reference, whose length check follows the push).  *out_bytes points at a
 * context-owned host buffer of ceil(*out_nbits / 8) bytes, valid until the next
 * call on this context. */
int spihtb_encode(spihtb_ctx *ctx, const int32_t *host_coeffs, int32_t c, int32_t h, int32_t w,
                  int32_t ll_h, int32_t ll_w, uint64_t max_bits,
                  const uint8_t **out_bytes, uint64_t *out_nbits, int32_t *out_max_n);
/* src/lib.rs:35-42  decode(data: bytes, n, c, h, w, ll_h, ll_w) -> int32[c,h,w]
 * -> src/encoder_decoder.rs:307-454.  All 8*nbytes bits are decoded (pad bits
 * included, lib.rs:15-21).  host_out holds c*h*w int32. */
int spihtb_decode(spihtb_ctx *ctx, const uint8_t *host_data, uint64_t nbytes, int32_t n,
                  int32_t c, int32_t h, int32_t w, int32_t ll_h, int32_t ll_w, int32_t *host_out);

/* src/lib.rs:47-56  decode_with_metadata(data, n, c, h, w, ll_h, ll_w, top_slice, other_slices)
 * -> (int32[c,h,w], int32[8*nbytes + 1, 8])  -> src/encoder_decoder.rs:616-841.
 * The second array holds, for every bit position (and one past the last), the decoder state just before that bit
 * is read: action id 0..6, position of the coefficient inside its band scaled to -100000..100000 (row, column),
 * channel, filter (0 LL, 1 DA, 2 AD, 3 DD), depth, n, current value of the coefficient (:616-630).
 * top_slice: {start_i, end_i, start_j, end_j} of the LL band; other_slices: int32 [levels][3][4], coarsest level
 * first, the three bands in the caller's order da, ad, dd (spiht_wrapper.py:232-250), each
 * {start_i, end_i, start_j, end_j}.  host_meta holds (8 * nbytes + 1) * 8 int32.  Even LL sizes only
 * (SPIHTB_EGEOM otherwise): with an odd LL band the reference's trees share cells. */
int spihtb_decode_with_metadata(spihtb_ctx *ctx, const uint8_t *host_data, uint64_t nbytes, int32_t n,
                                int32_t c, int32_t h, int32_t w, int32_t ll_h, int32_t ll_w,
                                const int32_t *top_slice, const int32_t *other_slices, int32_t levels,
                                int32_t *host_out, int32_t *host_meta);

/* ---- batched SPIHT coder, DEVICE buffers -------------------------------- */
/* Same coder over a batch of B coefficient arrays of one shape.
 * dev_max_bits: optional uint64[B] per-image budgets (NULL: use max_bits for all).
 * dev_out: B rows of out_stride bytes (out_stride % 8 == 0); a row is written
 * up to ceil(nbits/32) words.  A budget larger than the row is cut at the row
 * size and flagged in dev_status[b] |= 1 (dev_status may be NULL).
 * dev_nbits uint64[B], dev_max_n int32[B]. */
int spihtb_encode_coeffs(spihtb_ctx *ctx, const int32_t *dev_coeffs, int32_t B, int32_t c, int32_t h, int32_t w,
                         int32_t ll_h, int32_t ll_w, uint64_t max_bits, const uint64_t *dev_max_bits,
                         uint8_t *dev_out, uint64_t out_stride, uint64_t *dev_nbits, int32_t *dev_max_n,
                         int32_t *dev_status);
/* dev_in: B rows of in_stride bytes (in_stride % 8 == 0, rows zero-padded past
 * their length); dev_nbytes uint64[B]; dev_n int32[B] (max_n per stream);
 * dev_coeffs_out int32 [B][c][h][w] is overwritten (zero where nothing decoded). */
int spihtb_decode_coeffs(spihtb_ctx *ctx, const uint8_t *dev_in, uint64_t in_stride, const uint64_t *dev_nbytes,
                         const int32_t *dev_n, int32_t B, int32_t c, int32_t h, int32_t w,
                         int32_t ll_h, int32_t ll_w, int32_t *dev_coeffs_out);

/* Largest magnitude per image of a coefficient batch: dev_out uint32[B] = max |x| over the per_image int32 values
 * of image b.  (encoder_decoder.rs:165 takes the same maximum; the host uses it to size the stream rows of an
 * untruncated encode, spiht_wrapper.py:174-176 max_bits=None.)  B <= 65535. */
int spihtb_max_abs(spihtb_ctx *ctx, const int32_t *dev_coeffs, int32_t B, uint64_t per_image, uint32_t *dev_out);

/* ---- transform stages, DEVICE buffers ----------------------------------- */
/* spiht_wrapper.py:158-172: optional RGB->colour (color_models.py:6-13), pywt.wavedec2 (:163),
 * pywt.coeffs_to_array (:165), per-channel scale (:167-170), quantize (:9-11, truncation toward
 * zero of (m_c * x) * q).  ch_scales: host double[C] or NULL.  pixel_dtype: SPIHTB_F32/F64/U8.
 * dev_coeffs int32 [B][C][enc_h][enc_w]. */
int spihtb_forward(spihtb_ctx *ctx, const void *dev_pixels, int32_t pixel_dtype, int32_t B, int32_t C,
                   const spihtb_geom *geom, int32_t color_model, const double *ch_scales, double q,
                   int32_t *dev_coeffs);
/* spiht_wrapper.py:259-281: (x / m_c) / q, pywt.array_to_coeffs (:275), pywt.waverec2 (:276),
 * optional colour->RGB.  dev_pixels_out [B][C][rec_h][rec_w]. */
int spihtb_inverse(spihtb_ctx *ctx, const int32_t *dev_coeffs, int32_t B, int32_t C, const spihtb_geom *geom,
                   int32_t color_model, const double *ch_scales, double q,
                   void *dev_pixels_out, int32_t pixel_dtype);

/* spiht/color_models.py:6-13  convert(im, src, dest) for the one colour model on the accelerated path:
 * planar images [B][3][plane].  src_model / dst_model are SPIHTB_COLOR_NONE (= RGB) or SPIHTB_COLOR_IPT.
 * RGB -> IPT: dev_in is in_dtype (F32 / F64 / U8), dev_out float64.  IPT -> RGB: dev_in float64 (in_dtype
 * must be SPIHTB_F64), dev_out is out_dtype (F32 / F64).  Same models: a copy with conversion to out_dtype is
 * not offered -- the call fails with SPIHTB_EINVAL. */
int spihtb_convert_color(spihtb_ctx *ctx, const void *dev_in, int32_t in_dtype, int32_t B, uint64_t plane,
                         int32_t src_model, int32_t dst_model, void *dev_out, int32_t out_dtype);

/* ---- fused image path, DEVICE buffers ----------------------------------- */
/* encode_image (spiht_wrapper.py:142-189) over a batch: forward + encode_coeffs.
 * dev_coeffs_scratch: int32 [B][C][enc_h][enc_w] (kept as the parity artefact). */
int spihtb_encode_images(spihtb_ctx *ctx, const void *dev_pixels, int32_t pixel_dtype, int32_t B, int32_t C,
                         const spihtb_geom *geom, int32_t color_model, const double *ch_scales, double q,
                         uint64_t max_bits, const uint64_t *dev_max_bits, int32_t *dev_coeffs_scratch,
                         uint8_t *dev_out, uint64_t out_stride, uint64_t *dev_nbits, int32_t *dev_max_n,
                         int32_t *dev_status);
/* decode_image (spiht_wrapper.py:192-281) over a batch: decode_coeffs + inverse. */
int spihtb_decode_images(spihtb_ctx *ctx, const uint8_t *dev_in, uint64_t in_stride, const uint64_t *dev_nbytes,
                         const int32_t *dev_n, int32_t B, int32_t C, const spihtb_geom *geom,
                         int32_t color_model, const double *ch_scales, double q,
                         int32_t *dev_coeffs_scratch, void *dev_pixels_out, int32_t pixel_dtype);

/* bytes a bitstream row needs so that a full (untruncated) encode of one
 * [c][h][w] array with max_n <= 30 always fits; multiple of 8 */
uint64_t spihtb_stream_bound(int32_t c, int32_t h, int32_t w, int32_t ll_h, int32_t ll_w);

#ifdef __cplusplus
}
#endif
#endif /* SPIHT_B200_H */
