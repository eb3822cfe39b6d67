// Internal launch interfaces between the translation units of libspiht_b200.
#pragma once
#include "common.cuh"

namespace spihtb {

// ---- pyramid.cu
// base_done: dp already holds every node's 2x2-cell plane and maxabs the per-image maximum (the forward
// transform wrote them, see PyrFuse); only the rings and the LL roots remain.
int launch_pyramid(spihtb_ctx *ctx, const int32_t *coeffs, int B, int C, int H, int W, int ll_h, int ll_w,
                   uint8_t *dp, uint8_t *lp, uint8_t *dpll, uint8_t *lpll, uint32_t *maxabs, bool base_done = false);
// cells [a0,a1) x [b0,b1) of the node grid recomputed from the finished coefficient array
struct FixRect {
    int a0, a1, b0, b1;
};
int launch_pyr_fix(spihtb_ctx *ctx, const int32_t *coeffs, int nz, int H, int W, uint8_t *dp, const FixRect *dev_rects,
                   const uint32_t *dev_prefix, int nrect, uint32_t total);

// out[b] = max |x| over image b's per_image coefficients
int launch_max_abs(spihtb_ctx *ctx, const int32_t *coeffs, int B, size_t per_image, uint32_t *out);

// ---- color.cu: stand-alone RGB <-> IPT passes over planar [B][3][plane] images
int launch_rgb_to_ipt(spihtb_ctx *ctx, const void *src, int src_dtype, double *dst, size_t plane, int B);
int launch_ipt_to_rgb(spihtb_ctx *ctx, const double *src, void *dst, int dst_dtype, size_t plane, int B);
int ipt_upload_tables();  // power tables of the colour transform -> constant memory of the current device

// ---- spiht_enc.cu
struct EncArgs {
    const int32_t *coeffs;  // [B][C][H][W]
    int B, C, H, W, ll_h, ll_w;
    const uint8_t *dp, *lp, *dpll, *lpll;
    const uint32_t *maxabs;        // [B]
    uint64_t max_bits;             // 0 = unlimited
    const uint64_t *dev_max_bits;  // optional [B]
    uint8_t *out;
    uint64_t out_stride;  // bytes, multiple of 8
    uint64_t *nbits;      // [B]
    int32_t *max_n;       // [B]
    int32_t *status;      // optional [B]
};
struct EncPlan {
    uint64_t pix_cap, lis_cap;  // list capacities per resident CTA ("slot"), in entries
    size_t per_slot;            // bytes of list storage per slot
    int max_slots;              // CTAs of the coder the device can hold at once
};
int plan_encode(spihtb_ctx *ctx, const EncArgs &a, EncPlan *pl);
int launch_encode(spihtb_ctx *ctx, const EncArgs &a, const EncPlan &pl, void *lists, int slots, unsigned int *counter);
int launch_encode(spihtb_ctx *ctx, const EncArgs &a);  // whole batch, context workspace

// ---- spiht_dec.cu
struct DecArgs {
    const uint8_t *in;
    uint64_t in_stride;      // bytes, multiple of 8
    const uint64_t *nbytes;  // [B]
    const int32_t *n;        // [B]
    int B, C, H, W, ll_h, ll_w;
    int32_t *out;  // [B][C][H][W]
    // optional [B*C][ceil(H/64)][ceil(W/64)] bytes, zeroed by the caller: the decoder marks every 64x64 block
    // of the array it writes a coefficient into (the inverse transform skips the detail bands of unmarked blocks)
    uint8_t *blk = nullptr;
    // decode_with_metadata: rows [B][meta_rows][8] (zero-filled by the caller), band rectangles, error flag
    int32_t *meta = nullptr;
    uint64_t meta_rows = 0;
    int level = 0, top_ei = 1, top_ej = 1;
    const int32_t *slices = nullptr;
    int32_t *meta_err = nullptr;
    // optional second set of block marks, same shape as blk: only coefficients in the finest detail bands (row >=
    // fine_h0 or column >= fine_w0, the offsets of those bands in the array) mark it.  The level-1 inverse uses these:
    // a block on the edge of the finest bands is otherwise marked by coarser coefficients too.
    uint8_t *blk1 = nullptr;
    uint8_t *l1_any = nullptr;   // [B], zeroed by the caller: set to 1 for an image that marks anything in blk1
    int fine_h0 = 0, fine_w0 = 0;
    // lazy zero fill (needs blk1): only the corner above / left of the finest bands is zeroed before the launch; an
    // image zeroes its finest bands itself when its stream first reaches them (see DecK)
    bool lazy_zero = false;
};
int launch_decode(spihtb_ctx *ctx, const DecArgs &a);

// ---- dwt_fwd.cu / dwt_inv.cu
struct XformArgs {
    int B, C;
    spihtb_geom g;
    int color;
    double scale[8];  // per-channel multipliers m_c (1.0 when none)
    double q;
    int pixel_dtype;
    const uint8_t *blk = nullptr;  // inverse only: block marks of the coefficient array (see DecArgs), or null
    const uint8_t *blk1 = nullptr; // inverse only: the marks the finest level uses instead (see DecArgs), or null
    const uint8_t *l1_any = nullptr;  // inverse only: per image, 0 = no non-zero coefficient in the finest bands, or null
};
// Pyramid base pass fused into the forward transform: dp planes [B*C][enc_h/2][enc_w/2] and maxabs [B]
// are complete when launch_forward returns (stream order).
struct PyrFuse {
    uint8_t *dp;
    uint32_t *maxabs;
};
int launch_forward(spihtb_ctx *ctx, const void *pixels, const XformArgs &x, int32_t *coeffs,
                   const PyrFuse *pf = nullptr);
int launch_inverse(spihtb_ctx *ctx, const int32_t *coeffs, const XformArgs &x, void *pixels_out);
// ---- dwt_gen.cu: the rest of the bior family (filter length and taps as launch parameters)
#define SPIHTB_GEN_MAXF 20
bool wavelet_is_generic(int wid);
// PyWavelets' dec_lo / rec_lo of a generic wavelet (zero padding included), SPIHTB_GEN_MAXF entries each
bool generic_wavelet_taps(int wid, int *F, double *dec_lo, double *rec_lo);
struct GenFwdLevel {
    const void *src;       // [nz][src_h][src_w] of src_dtype
    int src_dtype;
    int src_h, src_w, bh, bw;
    double *dst_ll;        // [nz][bh][bw] float64 scratch (unused at the last level)
    int32_t *coeffs;       // [nz][Hc][Wc]
    int Hc, Wc, sh, sw, mode, C, last;
    double scale[8];
    double q;
};
int launch_gen_fwd_level(spihtb_ctx *ctx, int wid, const GenFwdLevel &a, int nz);
struct GenInvLevel {
    const double *src_a;   // [nz][a_h][a_w], or null at the coarsest level (LL corner of the array)
    int a_h, a_w;
    const int32_t *coeffs;
    int Hc, Wc, sh, sw, bh, bw, oh, ow, mode, C;
    double rscale[8];
    double rq;
    const uint8_t *blk;    // optional block marks (see DecArgs)
    int BH, BW;
    void *dst;             // [nz][oh][ow] float64, or float32 when out_f32
    bool out_f32;
};
int launch_gen_inv_level(spihtb_ctx *ctx, int wid, const GenInvLevel &a, int nz);

// dwt_fwd2.cu: levels 1 and 2 fused (TMA-staged tiles); *done = false when the geometry takes the other path
int launch_forward_fused12(spihtb_ctx *ctx, const void *src, int pixel_dtype, const XformArgs &x, int32_t *coeffs,
                           double *ll2, const PyrFuse *pf, const double *u8lut, bool *done);

}  // namespace spihtb
