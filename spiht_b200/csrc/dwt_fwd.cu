// Forward transform: (optional RGB->IPT) -> L-level separable 2-D DWT ->
// coeffs_to_array layout -> per-channel scale and truncating int32 quantiser.
// Replaces spiht_wrapper.py:158-172 (colour.convert, pywt.wavedec2,
// pywt.coeffs_to_array, channel_mults, quantize).
//
// One kernel per level.  A CTA produces a 32x32 tile of each of the four bands
// of one (image, channel) plane: it stages the (64+F-2)^2 input window in
// shared memory with the boundary rule applied on load, filters along axis -2
// then along axis -1 (PyWavelets' order) in float64, and writes
//   - the three detail bands, scaled and truncated to int32, straight to their
//     final place in the coefficient array, and
//   - the approximation band to a float64 scratch plane for the next level
//     (or, at the last level, quantised into the LL corner).
// analysis (non-periodization): out[k] = sum_j f[j] x_ext[2k + 1 - j]
// periodization:                out[k] = sum_j f[j] x_per[(2k + F/2 - j) mod Np]
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "wavelets.cuh"

namespace spihtb {

constexpr int FW_TH = 32, FW_TW = 32, FW_NT = 256;

struct FwdK {
    const void *src;        // [nz][src_h][src_w] planes of Tin
    int src_h, src_w;
    int bh, bw;             // band size of this level
    double *dst_ll;         // [nz][bh][bw] scratch (null at the last level)
    int32_t *coeffs;        // [nz][Hc][Wc]
    int Hc, Wc, sh, sw;     // detail block offsets of this level
    int mode, C, last;
    int tiles_x, tiles_y;
    double scale[8];
    double q;
};

// spiht_wrapper.py:9-11,167-172: ((m_c * x) * q).astype(int32), truncation toward zero
__device__ __forceinline__ int32_t quantise(double x, double m, double q) { return __double2int_rz((m * x) * q); }

template <typename Tin, int WID>
__global__ void __launch_bounds__(FW_NT) dwt_fwd_level_kernel(const FwdK p)
{
    constexpr int F = Wav<WID>::F;
    constexpr int IH = 2 * FW_TH + F - 2, IW = 2 * FW_TW + F - 2, IW2 = IW / 2;
    extern __shared__ __align__(16) unsigned char smem[];
    Tin *s_in = reinterpret_cast<Tin *>(smem);  // [IH][IW]
    double *s_v = reinterpret_cast<double *>(smem + ((IH * IW * sizeof(Tin) + 15) / 16) * 16);  // [2][TH][2][IW2]

    const int tid = threadIdx.x;
    uint32_t bid = blockIdx.x;
    const int tx = bid % p.tiles_x;
    bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    const int z = bid / p.tiles_y;

    const int r0 = ty * FW_TH, c0 = tx * FW_TW;
    const int shift = p.mode == SPIHTB_MODE_PERIODIZATION ? (F / 2 - 1) : 0;
    const int gr0 = 2 * r0 - (F - 2) + shift, gc0 = 2 * c0 - (F - 2) + shift;

    const Tin *src = static_cast<const Tin *>(p.src) + (size_t)z * p.src_h * p.src_w;
    for (int idx = tid; idx < IH * IW; idx += FW_NT) {
        const int li = idx / IW, lj = idx - li * IW;
        const int gi = ext_index(gr0 + li, p.src_h, p.mode);
        const int gj = ext_index(gc0 + lj, p.src_w, p.mode);
        s_in[idx] = src[(size_t)gi * p.src_w + gj];
    }
    __syncthreads();

    // axis -2: rows.  local input row of tap j for output row r: 2r + F-1 - j
    for (int idx = tid; idx < FW_TH * IW; idx += FW_NT) {
        const int r = idx / IW, x = idx - r * IW;
        double lo = 0.0, hi = 0.0;
#pragma unroll
        for (int j = 0; j < F; ++j) {
            const double v = (double)s_in[(2 * r + F - 1 - j) * IW + x];
            if (Wav<WID>::dec_lo(j) != 0.0) lo = fma(Wav<WID>::dec_lo(j), v, lo);
            if (wav_dec_hi<WID>(j) != 0.0) hi = fma(wav_dec_hi<WID>(j), v, hi);
        }
        const int o = (r * 2 + (x & 1)) * IW2 + (x >> 1);
        s_v[o] = lo;
        s_v[FW_TH * IW + o] = hi;
    }
    __syncthreads();

    // axis -1: columns, then write
    const int zc = z % p.C;
    const double m = p.scale[zc], q = p.q;
    int32_t *cz = p.coeffs + (size_t)z * p.Hc * p.Wc;
    for (int idx = tid; idx < FW_TH * FW_TW; idx += FW_NT) {
        const int r = idx / FW_TW, c = idx - r * FW_TW;
        double aa = 0.0, ad = 0.0, da = 0.0, dd = 0.0;
#pragma unroll
        for (int j = 0; j < F; ++j) {
            constexpr int dummy = 0;
            (void)dummy;
            const int col = F - 1 - j;  // + 2c
            const int o = (r * 2 + (col & 1)) * IW2 + c + (col >> 1);
            const double vlo = s_v[o], vhi = s_v[FW_TH * IW + o];
            if (Wav<WID>::dec_lo(j) != 0.0) {
                aa = fma(Wav<WID>::dec_lo(j), vlo, aa);
                da = fma(Wav<WID>::dec_lo(j), vhi, da);
            }
            if (wav_dec_hi<WID>(j) != 0.0) {
                ad = fma(wav_dec_hi<WID>(j), vlo, ad);
                dd = fma(wav_dec_hi<WID>(j), vhi, dd);
            }
        }
        const int gr = r0 + r, gc = c0 + c;
        if (gr < p.bh && gc < p.bw) {
            cz[(size_t)gr * p.Wc + p.sw + gc] = quantise(ad, m, q);
            cz[(size_t)(p.sh + gr) * p.Wc + gc] = quantise(da, m, q);
            cz[(size_t)(p.sh + gr) * p.Wc + p.sw + gc] = quantise(dd, m, q);
            if (p.last)
                cz[(size_t)gr * p.Wc + gc] = quantise(aa, m, q);
            else
                p.dst_ll[((size_t)z * p.bh + gr) * p.bw + gc] = aa;
        }
    }
}

// zero the gaps coeffs_to_array leaves between a level's off-diagonal blocks
// and the square of coarser levels: rows [bh,sh) x cols [sw,sw+bw) and
// rows [sh,sh+bh) x cols [bw,sw)
struct GapK {
    int32_t *coeffs;
    int Hc, Wc, nz, levels;
    int bh[SPIHTB_MAX_LEVELS], bw[SPIHTB_MAX_LEVELS], sh[SPIHTB_MAX_LEVELS], sw[SPIHTB_MAX_LEVELS];
};
__global__ void __launch_bounds__(256) gap_fill_kernel(const GapK p)
{
    for (int z = blockIdx.y; z < p.nz; z += gridDim.y) {
    int32_t *cz = p.coeffs + (size_t)z * p.Hc * p.Wc;
    for (int l = 0; l < p.levels; ++l) {
        const int gh = p.sh[l] - p.bh[l], gw = p.sw[l] - p.bw[l];
        const int n1 = gh > 0 ? gh * p.bw[l] : 0;
        const int n2 = gw > 0 ? p.bh[l] * gw : 0;
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n1 + n2; t += gridDim.x * blockDim.x) {
            if (t < n1) {
                const int r = t / p.bw[l], c = t - r * p.bw[l];
                cz[(size_t)(p.bh[l] + r) * p.Wc + p.sw[l] + c] = 0;
            } else {
                const int u = t - n1;
                const int r = u / gw, c = u - r * gw;
                cz[(size_t)(p.sh[l] + r) * p.Wc + p.bw[l] + c] = 0;
            }
        }
    }
    }
}

// ---- RGB -> IPT (color_models.py:6-13 -> colour.convert(.., 'RGB', 'IPT')) ----
// linear sRGB -> XYZ (4-digit IEC matrix) -> LMS -> sign(x)|x|^0.43 -> IPT, float64.
__device__ __forceinline__ double spow(double a, double e) { return a == 0.0 ? 0.0 : copysign(pow(fabs(a), e), a); }

template <typename Tin>
__global__ void __launch_bounds__(256) rgb_to_ipt_kernel(const Tin *__restrict__ src, double *__restrict__ dst,
                                                         size_t plane, size_t nimg)
{
    const size_t total = plane * nimg;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / plane, o = t - b * plane;
        const Tin *s = src + b * 3 * plane + o;
        const double R = (double)s[0], G = (double)s[plane], B = (double)s[2 * plane];
        const double X = 0.4124 * R + 0.3576 * G + 0.1805 * B;
        const double Y = 0.2126 * R + 0.7152 * G + 0.0722 * B;
        const double Z = 0.0193 * R + 0.1192 * G + 0.9505 * B;
        const double L = spow(0.4002 * X + 0.7075 * Y + -0.0807 * Z, 0.43);
        const double M = spow(-0.2280 * X + 1.1500 * Y + 0.0612 * Z, 0.43);
        const double S = spow(0.0 * X + 0.0 * Y + 0.9184 * Z, 0.43);
        double *d = dst + b * 3 * plane + o;
        d[0] = 0.4000 * L + 0.4000 * M + 0.2000 * S;
        d[plane] = 4.4550 * L + -4.8510 * M + 0.3960 * S;
        d[2 * plane] = 0.8056 * L + 0.3572 * M + -1.1628 * S;
    }
}

template <typename Tin, int WID>
static int launch_level(spihtb_ctx *ctx, const FwdK &k, int nz)
{
    constexpr int F = Wav<WID>::F;
    constexpr int IH = 2 * FW_TH + F - 2, IW = 2 * FW_TW + F - 2;
    const size_t smem = ((IH * IW * sizeof(Tin) + 15) / 16) * 16 + 2 * FW_TH * IW * sizeof(double);
    auto kern = dwt_fwd_level_kernel<Tin, WID>;
    static bool attr_set = false;
    if (!attr_set) {
        SPIHTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long nb = (long long)k.tiles_x * k.tiles_y * nz;
    if (nb > 0x7fffffffLL) {
        set_error("forward DWT grid too large");
        return SPIHTB_ESHAPE;
    }
    kern<<<(unsigned)nb, FW_NT, smem, ctx->stream>>>(k);
    ctx->launches++;
    return SPIHTB_OK;
}

template <typename Tin>
static int launch_level_w(spihtb_ctx *ctx, int wid, const FwdK &k, int nz)
{
    switch (wid) {
        case SPIHTB_WAVELET_BIOR22: return launch_level<Tin, SPIHTB_WAVELET_BIOR22>(ctx, k, nz);
        case SPIHTB_WAVELET_BIOR44: return launch_level<Tin, SPIHTB_WAVELET_BIOR44>(ctx, k, nz);
        case SPIHTB_WAVELET_BIOR68: return launch_level<Tin, SPIHTB_WAVELET_BIOR68>(ctx, k, nz);
    }
    set_error("unknown wavelet id %d", wid);
    return SPIHTB_EINVAL;
}

int launch_forward(spihtb_ctx *ctx, const void *pixels, const XformArgs &x, int32_t *coeffs)
{
    const spihtb_geom &g = x.g;
    const int nz = x.B * x.C;
    const int L = g.levels;
    // scratch: two float64 approximation planes (level 1 output is the largest)
    const size_t ll1 = (size_t)nz * g.band_h[0] * g.band_w[0] * sizeof(double);
    const size_t ll2 = L > 1 ? (size_t)nz * g.band_h[1] * g.band_w[1] * sizeof(double) : 0;
    int rc = ctx->ensure(ctx->tmpa, ll1 + 256);
    if (rc) return rc;
    rc = ctx->ensure(ctx->tmpb, ll2 + 256);
    if (rc) return rc;

    const void *src = pixels;
    bool src_is_f64 = x.pixel_dtype == SPIHTB_F64;
    if (x.color == SPIHTB_COLOR_IPT) {
        if (x.C != 3) {
            set_error("IPT colour model needs 3 channels, got %d", x.C);
            return SPIHTB_EINVAL;
        }
        const size_t plane = (size_t)g.h * g.w;
        rc = ctx->ensure(ctx->io2, (size_t)nz * plane * sizeof(double) + 256);
        if (rc) return rc;
        const unsigned nb = (unsigned)std::min<size_t>((plane * x.B + 255) / 256, (size_t)ctx->sm_count * 32);
        if (src_is_f64)
            rgb_to_ipt_kernel<double><<<nb, 256, 0, ctx->stream>>>(static_cast<const double *>(pixels),
                                                                   static_cast<double *>(ctx->io2.p), plane, x.B);
        else
            rgb_to_ipt_kernel<float><<<nb, 256, 0, ctx->stream>>>(static_cast<const float *>(pixels),
                                                                  static_cast<double *>(ctx->io2.p), plane, x.B);
        ctx->launches++;
        src = ctx->io2.p;
        src_is_f64 = true;
    }

    for (int l = 0; l < L; ++l) {
        FwdK k;
        k.src = src;
        k.src_h = g.in_h[l];
        k.src_w = g.in_w[l];
        k.bh = g.band_h[l];
        k.bw = g.band_w[l];
        k.last = (l == L - 1);
        k.dst_ll = k.last ? nullptr : static_cast<double *>((l & 1) ? ctx->tmpb.p : ctx->tmpa.p);
        k.coeffs = coeffs;
        k.Hc = g.enc_h;
        k.Wc = g.enc_w;
        k.sh = g.off_h[l];
        k.sw = g.off_w[l];
        k.mode = g.mode;
        k.C = x.C;
        k.tiles_x = (k.bw + FW_TW - 1) / FW_TW;
        k.tiles_y = (k.bh + FW_TH - 1) / FW_TH;
        for (int c = 0; c < 8; ++c) k.scale[c] = x.scale[c];
        k.q = x.q;
        if (l == 0) ctx->stage_begin(0);
        if (l == 1) ctx->stage_begin(1);
        rc = src_is_f64 ? launch_level_w<double>(ctx, g.wavelet, k, nz) : launch_level_w<float>(ctx, g.wavelet, k, nz);
        if (rc) return rc;
        if (l == 0) ctx->stage_end(0);
        src = k.dst_ll;
        src_is_f64 = true;
    }
    {
        GapK gk;
        gk.coeffs = coeffs;
        gk.Hc = g.enc_h;
        gk.Wc = g.enc_w;
        gk.nz = nz;
        gk.levels = L;
        bool any = false;
        for (int l = 0; l < L; ++l) {
            gk.bh[l] = g.band_h[l];
            gk.bw[l] = g.band_w[l];
            gk.sh[l] = g.off_h[l];
            gk.sw[l] = g.off_w[l];
            any |= (gk.sh[l] > gk.bh[l]) || (gk.sw[l] > gk.bw[l]);
        }
        if (L == 1) ctx->stage_begin(1);
        if (any) {
            gap_fill_kernel<<<dim3(8, std::min(nz, 65535)), 256, 0, ctx->stream>>>(gk);
            ctx->launches++;
        }
        ctx->stage_end(1);
    }
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
