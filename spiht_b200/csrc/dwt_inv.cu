// Inverse transform: dequantise -> array_to_coeffs -> L-level inverse DWT ->
// (optional IPT->RGB).  Replaces spiht_wrapper.py:259-281 ((x / m_c) / q,
// pywt.array_to_coeffs, pywt.waverec2, colour.convert).
//
// One kernel per level, coarsest first.  A CTA produces a 32x64 tile of the
// level's output for one (image, channel) plane: it stages the needed window
// of the four bands in shared memory (details are dequantised on load from
// their place in the coefficient array), synthesises along axis -1 then along
// axis -2 (PyWavelets' idwtn order) in float64.
// synthesis (non-periodization): rec[n] = sum_t g[t] c[(n + F-2 - t)/2]   for even n+F-2-t, 0 <= k < m
// periodization:                 rec[n] = sum_t g[t] c[((n + F/2-1 - t)/2) mod m]
// waverec2 drops the approximation's trailing row/column when it is one
// longer than the detail band (odd sizes); the window simply never reads it.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "wavelets.cuh"

namespace spihtb {

constexpr int IV_TH = 32, IV_TW = 64, IV_NT = 256;

struct InvK {
    const double *src_a;  // approximation planes [nz][a_h][a_w] (null at the coarsest level: LL corner of coeffs)
    int a_h, a_w;
    const int32_t *coeffs;  // [nz][Hc][Wc]
    int Hc, Wc, sh, sw;
    int bh, bw;   // band size
    int oh, ow;   // output size of this level
    void *dst;    // [nz][oh][ow] double scratch, or the final pixels
    int dst_f32;  // final level only: 1 = float32 output
    int mode, C;
    int tiles_x, tiles_y;
    double scale[8];
    double q;
};

__device__ __forceinline__ int floor_div2(int a) { return a >> 1; }  // arithmetic shift = floor for negatives

template <int WID>
__global__ void __launch_bounds__(IV_NT) dwt_inv_level_kernel(const InvK p)
{
    constexpr int F = Wav<WID>::F;
    constexpr int KH = IV_TH / 2 + F / 2 + 1, KW = IV_TW / 2 + F / 2 + 1;
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_b = reinterpret_cast<double *>(smem);  // [4][KH][KW]: aa, ad, da, dd
    double *s_h = s_b + 4 * KH * KW;                 // [2][KH][IV_TW]

    const int tid = threadIdx.x;
    uint32_t bid = blockIdx.x;
    const int tx = bid % p.tiles_x;
    bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    const int z = bid / p.tiles_y;
    const int n0r = ty * IV_TH, n0c = tx * IV_TW;
    const bool per = p.mode == SPIHTB_MODE_PERIODIZATION;
    const int S = per ? (F / 2 - 1) : (F - 2);
    const int klo_r = floor_div2(n0r + S - (F - 1)), klo_c = floor_div2(n0c + S - (F - 1));

    const int zc = z % p.C;
    const double m = p.scale[zc], q = p.q;
    const int32_t *cz = p.coeffs + (size_t)z * p.Hc * p.Wc;
    const double *az = p.src_a ? p.src_a + (size_t)z * p.a_h * p.a_w : nullptr;

    for (int idx = tid; idx < KH * KW; idx += IV_NT) {
        const int lr = idx / KW, lc = idx - lr * KW;
        int kr = klo_r + lr, kc = klo_c + lc;
        bool ok = true;
        if (per) {
            kr %= p.bh;
            if (kr < 0) kr += p.bh;
            kc %= p.bw;
            if (kc < 0) kc += p.bw;
        } else {
            ok = kr >= 0 && kr < p.bh && kc >= 0 && kc < p.bw;
        }
        double aa = 0.0, ad = 0.0, da = 0.0, dd = 0.0;
        if (ok) {
            // spiht_wrapper.py:270-274: (x / m_c) / q
            aa = az ? az[(size_t)kr * p.a_w + kc] : ((double)cz[(size_t)kr * p.Wc + kc] / m) / q;
            ad = ((double)cz[(size_t)kr * p.Wc + p.sw + kc] / m) / q;
            da = ((double)cz[(size_t)(p.sh + kr) * p.Wc + kc] / m) / q;
            dd = ((double)cz[(size_t)(p.sh + kr) * p.Wc + p.sw + kc] / m) / q;
        }
        s_b[idx] = aa;
        s_b[KH * KW + idx] = ad;
        s_b[2 * KH * KW + idx] = da;
        s_b[3 * KH * KW + idx] = dd;
    }
    __syncthreads();

    // axis -1: every staged coefficient row -> IV_TW output columns
    for (int idx = tid; idx < KH * IV_TW; idx += IV_NT) {
        const int lr = idx / IV_TW, c = idx - lr * IV_TW;
        const int n = n0c + c;
        const int par = (n + S) & 1;
        const int kbase = ((n + S - par) >> 1) - klo_c;  // local index of tap t = par
        double lo = 0.0, hi = 0.0;
        const double *ra = s_b + lr * KW;
        if (par == 0) {
#pragma unroll
            for (int u = 0; u < F / 2; ++u) {
                const int t = 2 * u;
                const int kk = kbase - u;
                if (Wav<WID>::rec_lo(t) != 0.0) {
                    lo = fma(Wav<WID>::rec_lo(t), ra[kk], lo);
                    hi = fma(Wav<WID>::rec_lo(t), ra[2 * KH * KW + kk], hi);
                }
                if (wav_rec_hi<WID>(t) != 0.0) {
                    lo = fma(wav_rec_hi<WID>(t), ra[KH * KW + kk], lo);
                    hi = fma(wav_rec_hi<WID>(t), ra[3 * KH * KW + kk], hi);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < F / 2; ++u) {
                const int t = 2 * u + 1;
                const int kk = kbase - u;
                if (Wav<WID>::rec_lo(t) != 0.0) {
                    lo = fma(Wav<WID>::rec_lo(t), ra[kk], lo);
                    hi = fma(Wav<WID>::rec_lo(t), ra[2 * KH * KW + kk], hi);
                }
                if (wav_rec_hi<WID>(t) != 0.0) {
                    lo = fma(wav_rec_hi<WID>(t), ra[KH * KW + kk], lo);
                    hi = fma(wav_rec_hi<WID>(t), ra[3 * KH * KW + kk], hi);
                }
            }
        }
        s_h[idx] = lo;
        s_h[KH * IV_TW + idx] = hi;
    }
    __syncthreads();

    // axis -2
    for (int idx = tid; idx < IV_TH * IV_TW; idx += IV_NT) {
        const int r = idx / IV_TW, c = idx - r * IV_TW;
        const int n = n0r + r;
        const int par = (n + S) & 1;
        const int kbase = ((n + S - par) >> 1) - klo_r;
        double v = 0.0;
        if (par == 0) {
#pragma unroll
            for (int u = 0; u < F / 2; ++u) {
                const int t = 2 * u;
                const int kk = (kbase - u) * IV_TW + c;
                if (Wav<WID>::rec_lo(t) != 0.0) v = fma(Wav<WID>::rec_lo(t), s_h[kk], v);
                if (wav_rec_hi<WID>(t) != 0.0) v = fma(wav_rec_hi<WID>(t), s_h[KH * IV_TW + kk], v);
            }
        } else {
#pragma unroll
            for (int u = 0; u < F / 2; ++u) {
                const int t = 2 * u + 1;
                const int kk = (kbase - u) * IV_TW + c;
                if (Wav<WID>::rec_lo(t) != 0.0) v = fma(Wav<WID>::rec_lo(t), s_h[kk], v);
                if (wav_rec_hi<WID>(t) != 0.0) v = fma(wav_rec_hi<WID>(t), s_h[KH * IV_TW + kk], v);
            }
        }
        const int gc = n0c + c;
        if (n < p.oh && gc < p.ow) {
            const size_t o = ((size_t)z * p.oh + n) * p.ow + gc;
            if (p.dst_f32)
                static_cast<float *>(p.dst)[o] = (float)v;
            else
                static_cast<double *>(p.dst)[o] = v;
        }
    }
}

// ---- IPT -> RGB (colour.convert(.., 'IPT', 'RGB')): numpy-inverted IPT matrices,
// exponent 1/0.43, then the hard-coded 4-digit XYZ -> sRGB matrix ----
struct IptInv {
    double ipt2lms[9];  // inv(M_LMS'->IPT)
    double lms2xyz[9];  // inv(M_XYZ->LMS)
};
__device__ __forceinline__ double spow_inv(double a, double e) { return a == 0.0 ? 0.0 : copysign(pow(fabs(a), e), a); }

template <typename Tout>
__global__ void __launch_bounds__(256) ipt_to_rgb_kernel(const double *__restrict__ src, Tout *__restrict__ dst,
                                                         size_t plane, size_t nimg, const IptInv mi)
{
    const size_t total = plane * nimg;
    const double e = 1.0 / 0.43;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / plane, o = t - b * plane;
        const double *s = src + b * 3 * plane + o;
        const double I = s[0], P = s[plane], T = s[2 * plane];
        const double L = spow_inv(mi.ipt2lms[0] * I + mi.ipt2lms[1] * P + mi.ipt2lms[2] * T, e);
        const double M = spow_inv(mi.ipt2lms[3] * I + mi.ipt2lms[4] * P + mi.ipt2lms[5] * T, e);
        const double Sv = spow_inv(mi.ipt2lms[6] * I + mi.ipt2lms[7] * P + mi.ipt2lms[8] * T, e);
        const double X = mi.lms2xyz[0] * L + mi.lms2xyz[1] * M + mi.lms2xyz[2] * Sv;
        const double Y = mi.lms2xyz[3] * L + mi.lms2xyz[4] * M + mi.lms2xyz[5] * Sv;
        const double Z = mi.lms2xyz[6] * L + mi.lms2xyz[7] * M + mi.lms2xyz[8] * Sv;
        Tout *d = dst + b * 3 * plane + o;
        d[0] = (Tout)(3.2406 * X + -1.5372 * Y + -0.4986 * Z);
        d[plane] = (Tout)(-0.9689 * X + 1.8758 * Y + 0.0415 * Z);
        d[2 * plane] = (Tout)(0.0557 * X + -0.2040 * Y + 1.0570 * Z);
    }
}

static void inv3(const double a[9], double r[9])
{
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                       a[2] * (a[3] * a[7] - a[4] * a[6]);
    r[0] = (a[4] * a[8] - a[5] * a[7]) / det;
    r[1] = (a[2] * a[7] - a[1] * a[8]) / det;
    r[2] = (a[1] * a[5] - a[2] * a[4]) / det;
    r[3] = (a[5] * a[6] - a[3] * a[8]) / det;
    r[4] = (a[0] * a[8] - a[2] * a[6]) / det;
    r[5] = (a[2] * a[3] - a[0] * a[5]) / det;
    r[6] = (a[3] * a[7] - a[4] * a[6]) / det;
    r[7] = (a[1] * a[6] - a[0] * a[7]) / det;
    r[8] = (a[0] * a[4] - a[1] * a[3]) / det;
}

template <int WID>
static int launch_inv_level(spihtb_ctx *ctx, const InvK &k, int nz)
{
    constexpr int F = Wav<WID>::F;
    constexpr int KH = IV_TH / 2 + F / 2 + 1, KW = IV_TW / 2 + F / 2 + 1;
    const size_t smem = (4 * KH * KW + 2 * KH * IV_TW) * sizeof(double);
    auto kern = dwt_inv_level_kernel<WID>;
    static bool attr_set = false;
    if (!attr_set) {
        SPIHTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long nb = (long long)k.tiles_x * k.tiles_y * nz;
    if (nb > 0x7fffffffLL) {
        set_error("inverse DWT grid too large");
        return SPIHTB_ESHAPE;
    }
    kern<<<(unsigned)nb, IV_NT, smem, ctx->stream>>>(k);
    ctx->launches++;
    return SPIHTB_OK;
}

int launch_inverse(spihtb_ctx *ctx, const int32_t *coeffs, const XformArgs &x, void *pixels_out)
{
    const spihtb_geom &g = x.g;
    const int nz = x.B * x.C;
    const int L = g.levels;
    const bool per = g.mode == SPIHTB_MODE_PERIODIZATION;
    const int F = wavelet_flen(g.wavelet);
    auto out_len = [&](int m) { return per ? 2 * m : 2 * m - F + 2; };

    // scratch planes: outputs of levels L-1 .. 1 (level 0 writes the image)
    size_t big = 0;
    for (int l = 1; l < L; ++l)
        big = std::max(big, (size_t)out_len(g.band_h[l]) * out_len(g.band_w[l]));
    const size_t img = (size_t)g.rec_h * g.rec_w;
    const bool color = x.color == SPIHTB_COLOR_IPT;
    if (color && x.C != 3) {
        set_error("IPT colour model needs 3 channels, got %d", x.C);
        return SPIHTB_EINVAL;
    }
    int rc = ctx->ensure(ctx->tmpa, (size_t)nz * big * sizeof(double) + 256);
    if (rc) return rc;
    rc = ctx->ensure(ctx->tmpb, (size_t)nz * big * sizeof(double) + 256);
    if (rc) return rc;
    if (color) {
        rc = ctx->ensure(ctx->io2, (size_t)nz * img * sizeof(double) + 256);
        if (rc) return rc;
    }

    const double *src = nullptr;
    int a_h = 0, a_w = 0;
    if (L > 1) ctx->stage_begin(6);
    for (int l = L - 1; l >= 0; --l) {
        InvK k;
        if (l == 0) {
            if (L > 1) ctx->stage_end(6);
            ctx->stage_begin(7);
        }
        k.src_a = src;
        k.a_h = a_h;
        k.a_w = a_w;
        k.coeffs = coeffs;
        k.Hc = g.enc_h;
        k.Wc = g.enc_w;
        k.sh = g.off_h[l];
        k.sw = g.off_w[l];
        k.bh = g.band_h[l];
        k.bw = g.band_w[l];
        k.oh = out_len(k.bh);
        k.ow = out_len(k.bw);
        if (k.oh <= 0 || k.ow <= 0) {
            set_error("band too short for this wavelet");
            return SPIHTB_EINVAL;
        }
        k.mode = g.mode;
        k.C = x.C;
        k.tiles_x = (k.ow + IV_TW - 1) / IV_TW;
        k.tiles_y = (k.oh + IV_TH - 1) / IV_TH;
        for (int c = 0; c < 8; ++c) k.scale[c] = x.scale[c];
        k.q = x.q;
        if (l == 0) {
            k.dst = color ? ctx->io2.p : pixels_out;
            k.dst_f32 = (!color && x.pixel_dtype == SPIHTB_F32) ? 1 : 0;
        } else {
            k.dst = ((L - 1 - l) & 1) ? ctx->tmpb.p : ctx->tmpa.p;
            k.dst_f32 = 0;
        }
        switch (g.wavelet) {
            case SPIHTB_WAVELET_BIOR22: rc = launch_inv_level<SPIHTB_WAVELET_BIOR22>(ctx, k, nz); break;
            case SPIHTB_WAVELET_BIOR44: rc = launch_inv_level<SPIHTB_WAVELET_BIOR44>(ctx, k, nz); break;
            case SPIHTB_WAVELET_BIOR68: rc = launch_inv_level<SPIHTB_WAVELET_BIOR68>(ctx, k, nz); break;
            default: set_error("unknown wavelet id %d", g.wavelet); rc = SPIHTB_EINVAL;
        }
        if (rc) return rc;
        src = static_cast<const double *>(k.dst);
        a_h = k.oh;
        a_w = k.ow;
    }
    if (color) {
        IptInv mi;
        const double lms2ipt[9] = {0.4000, 0.4000, 0.2000, 4.4550, -4.8510, 0.3960, 0.8056, 0.3572, -1.1628};
        const double xyz2lms[9] = {0.4002, 0.7075, -0.0807, -0.2280, 1.1500, 0.0612, 0.0, 0.0, 0.9184};
        inv3(lms2ipt, mi.ipt2lms);
        inv3(xyz2lms, mi.lms2xyz);
        const unsigned nb = (unsigned)std::min<size_t>((img * x.B + 255) / 256, (size_t)ctx->sm_count * 32);
        if (x.pixel_dtype == SPIHTB_F32)
            ipt_to_rgb_kernel<float><<<nb, 256, 0, ctx->stream>>>(static_cast<const double *>(ctx->io2.p),
                                                                  static_cast<float *>(pixels_out), img, x.B, mi);
        else
            ipt_to_rgb_kernel<double><<<nb, 256, 0, ctx->stream>>>(static_cast<const double *>(ctx->io2.p),
                                                                   static_cast<double *>(pixels_out), img, x.B, mi);
        ctx->launches++;
    }
    ctx->stage_end(7);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
