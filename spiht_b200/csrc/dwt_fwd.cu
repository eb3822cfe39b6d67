// Forward transform: (optional RGB->IPT) -> L-level separable 2-D DWT ->
// coeffs_to_array layout -> per-channel scale and truncating int32 quantiser.
// Replaces spiht_wrapper.py:158-172 (colour.convert, pywt.wavedec2,
// pywt.coeffs_to_array, channel_mults, quantize).
//
// One kernel per level.  A warp streams a strip of one (image, channel) plane
// (see dwt_fwd_task): filters along axis -2 in registers as it walks down the
// rows, then along axis -1 through warp shuffles (PyWavelets' order, float64),
// and writes
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

constexpr int FW_WARPS = 4;  // warps (= independent tasks) per CTA

struct FwdK {
    const void *src;        // [nz][src_h][src_w] planes of Tin
    int src_h, src_w;
    int bh, bw;             // band size of this level
    double *dst_ll;         // [nz][bh][bw] scratch (null at the last level)
    int32_t *coeffs;        // [nz][Hc][Wc]
    int Hc, Wc, sh, sw;     // detail block offsets of this level
    int mode, C, last;
    int tiles_x, tiles_y;   // strips across, row chunks down
    int RH;                 // output rows per chunk
    long long ntasks;       // nz * tiles_y * tiles_x
    double scale[8];
    double q;
};

// spiht_wrapper.py:9-11,167-172: ((m_c * x) * q).astype(int32), truncation toward zero
__device__ __forceinline__ int32_t quantise(double x, double m, double q) { return __double2int_rz((m * x) * q); }

template <typename T>
struct Pair;
template <>
struct Pair<float> {
    using type = float2;
};
template <>
struct Pair<double> {
    using type = double2;
};

// One warp = one task: a strip of NOUT = 32 - (F/2 - 1) output columns by RH output
// rows of one (image, channel) plane.  Lane l owns the input column pair
// (E, O) = (x[2m], x[2m+1]), m = k0 - (F/2-1) + l, and walks down the rows:
//   axis -2: a register window of F rows of its two columns gives the row-filtered
//            (lo, hi) pair of both columns -- no exchange needed;
//   axis -1: out[k] = sum_u f[2u] O[k-u] + f[2u+1] E[k-u] takes the neighbours'
//            values by warp shuffle (lanes l-1 .. l-(F/2-1)); lanes >= F/2-1 own an
//            output column.
// No shared memory and no barrier: warps are independent, every input sample is
// read from HBM once per strip (the F-2 halo columns hit L1/L2), loads are
// prefetched PD row pairs ahead, each lane writes its four band values
// (approximation as float64 scratch for the next level, details quantised).
template <typename Tin, int WID, bool INSIDE>
__device__ __forceinline__ void dwt_fwd_task(const FwdK &p, int tx, int ty, int z)
{
    constexpr int F = Wav<WID>::F;
    constexpr int HF = F / 2;
    constexpr int NOUT = 32 - (HF - 1);
    constexpr int PD = (HF % 3 == 0) ? 3 : HF;  // prefetch distance in row pairs; divides HF
    using Tin2 = typename Pair<Tin>::type;
    const int lane = threadIdx.x & 31;
    const int src_h = p.src_h, src_w = p.src_w, mode = p.mode;
    const int sft = mode == SPIHTB_MODE_PERIODIZATION ? (HF - 1) : 0;  // even for every supported wavelet
    const int k0 = tx * NOUT, r0 = ty * p.RH;
    const int nrows = min(p.RH, p.bh - r0);
    const int k = k0 - (HF - 1) + lane;  // this lane's pair index = its output column
    const int gcE = 2 * k + sft;
    const int gr0 = 2 * r0 - (F - 2) + sft;

    const Tin *plane = static_cast<const Tin *>(p.src) + (size_t)z * src_h * src_w;
    // whole strip inside the plane and rows pair-aligned: one vector load per row
    const int gc_first = 2 * (k0 - (HF - 1)) + sft;
    const bool vec = gc_first >= 0 && gc_first + 63 < src_w && (src_w & 1) == 0 &&
                     (reinterpret_cast<uintptr_t>(plane) & (2 * sizeof(Tin) - 1)) == 0;
    const int colE = vec ? gcE : ext_index(gcE, src_w, mode);
    const int colO = vec ? gcE + 1 : ext_index(gcE + 1, src_w, mode);

    auto load_row = [&](int li, Tin &e, Tin &o) {
        const int gi = INSIDE ? gr0 + li : ext_index(gr0 + li, src_h, mode);
        const Tin *row = plane + (size_t)gi * src_w;
        if (vec) {
            const Tin2 v = __ldg(reinterpret_cast<const Tin2 *>(row + colE));
            e = v.x;
            o = v.y;
        } else {
            e = __ldg(row + colE);
            o = __ldg(row + colO);
        }
    };

    // local row li of the chunk lives in window slot li % F; output row i needs rows 2i .. 2i+F-1
    double wE[F], wO[F];
#pragma unroll
    for (int t = 0; t < F - 2; ++t) {
        Tin e, o;
        load_row(t, e, o);
        wE[t] = (double)e;
        wO[t] = (double)o;
    }
    Tin qE[PD][2], qO[PD][2];
#pragma unroll
    for (int s = 0; s < PD; ++s) {
        qE[s][0] = qE[s][1] = qO[s][0] = qO[s][1] = (Tin)0;
        if (s < nrows) {
            load_row(F - 2 + 2 * s, qE[s][0], qO[s][0]);
            load_row(F - 1 + 2 * s, qE[s][1], qO[s][1]);
        }
    }

    const int zc = z % p.C;
    const double m = p.scale[zc], q = p.q;
    const bool out_active = lane >= HF - 1 && k < p.bw;
    const int Wc = p.Wc;
    int32_t *cz = p.coeffs + (size_t)z * p.Hc * Wc;
    int32_t *p_aa = cz + (size_t)r0 * Wc + k;           // LL corner (last level only)
    int32_t *p_ad = p_aa + p.sw;                        // rows lo, cols hi: top right
    int32_t *p_da = cz + (size_t)(p.sh + r0) * Wc + k;  // rows hi, cols lo: bottom left
    int32_t *p_dd = p_da + p.sw;
    const bool ll_scratch = !p.last;
    double *p_ll = ll_scratch ? p.dst_ll + ((size_t)z * p.bh + r0) * p.bw + k : nullptr;
    const int bw = p.bw;

    for (int rb = 0; rb < nrows; rb += HF) {
#pragma unroll
        for (int u = 0; u < HF; ++u) {
            const int i = rb + u;
            if (i < nrows) {
                const int slot = u % PD;
                wE[(2 * u + F - 2) % F] = (double)qE[slot][0];
                wO[(2 * u + F - 2) % F] = (double)qO[slot][0];
                wE[(2 * u + F - 1) % F] = (double)qE[slot][1];
                wO[(2 * u + F - 1) % F] = (double)qO[slot][1];
                if (i + PD < nrows) {
                    load_row(2 * (i + PD) + F - 2, qE[slot][0], qO[slot][0]);
                    load_row(2 * (i + PD) + F - 1, qE[slot][1], qO[slot][1]);
                }
                // axis -2: tap j multiplies local row 2i + F-1 - j
                double loE = 0.0, loO = 0.0, hiE = 0.0, hiO = 0.0;
#pragma unroll
                for (int j = 0; j < F; ++j) {
                    const double ve = wE[(2 * u + F - 1 - j) % F], vo = wO[(2 * u + F - 1 - j) % F];
                    if (Wav<WID>::dec_lo(j) != 0.0) {
                        loE = fma(Wav<WID>::dec_lo(j), ve, loE);
                        loO = fma(Wav<WID>::dec_lo(j), vo, loO);
                    }
                    if (wav_dec_hi<WID>(j) != 0.0) {
                        hiE = fma(wav_dec_hi<WID>(j), ve, hiE);
                        hiO = fma(wav_dec_hi<WID>(j), vo, hiO);
                    }
                }
                // axis -1: tap 2u' multiplies O[k-u'], tap 2u'+1 multiplies E[k-u']
                double aa = 0.0, ad = 0.0, da = 0.0, dd = 0.0;
#pragma unroll
                for (int v = 0; v < HF; ++v) {
                    constexpr unsigned FULL = 0xffffffffu;
                    const bool needO = Wav<WID>::dec_lo(2 * v) != 0.0 || wav_dec_hi<WID>(2 * v) != 0.0;
                    const bool needE = Wav<WID>::dec_lo(2 * v + 1) != 0.0 || wav_dec_hi<WID>(2 * v + 1) != 0.0;
                    if (needO) {
                        const double lo_o = v ? __shfl_up_sync(FULL, loO, v) : loO;
                        const double hi_o = v ? __shfl_up_sync(FULL, hiO, v) : hiO;
                        if (Wav<WID>::dec_lo(2 * v) != 0.0) {
                            aa = fma(Wav<WID>::dec_lo(2 * v), lo_o, aa);
                            da = fma(Wav<WID>::dec_lo(2 * v), hi_o, da);
                        }
                        if (wav_dec_hi<WID>(2 * v) != 0.0) {
                            ad = fma(wav_dec_hi<WID>(2 * v), lo_o, ad);
                            dd = fma(wav_dec_hi<WID>(2 * v), hi_o, dd);
                        }
                    }
                    if (needE) {
                        const double lo_e = v ? __shfl_up_sync(FULL, loE, v) : loE;
                        const double hi_e = v ? __shfl_up_sync(FULL, hiE, v) : hiE;
                        if (Wav<WID>::dec_lo(2 * v + 1) != 0.0) {
                            aa = fma(Wav<WID>::dec_lo(2 * v + 1), lo_e, aa);
                            da = fma(Wav<WID>::dec_lo(2 * v + 1), hi_e, da);
                        }
                        if (wav_dec_hi<WID>(2 * v + 1) != 0.0) {
                            ad = fma(wav_dec_hi<WID>(2 * v + 1), lo_e, ad);
                            dd = fma(wav_dec_hi<WID>(2 * v + 1), hi_e, dd);
                        }
                    }
                }
                if (out_active) {
                    *p_ad = quantise(ad, m, q);
                    *p_da = quantise(da, m, q);
                    *p_dd = quantise(dd, m, q);
                    if (ll_scratch)
                        *p_ll = aa;
                    else
                        *p_aa = quantise(aa, m, q);
                }
                p_aa += Wc;
                p_ad += Wc;
                p_da += Wc;
                p_dd += Wc;
                if (ll_scratch) p_ll += bw;
            }
        }
    }
}

template <typename Tin, int WID>
__global__ void __launch_bounds__(FW_WARPS * 32) dwt_fwd_level_kernel(const FwdK p)
{
    constexpr int F = Wav<WID>::F;
    long long task = (long long)blockIdx.x * FW_WARPS + (threadIdx.x >> 5);
    if (task >= p.ntasks) return;
    const int tx = (int)(task % p.tiles_x);
    task /= p.tiles_x;
    const int ty = (int)(task % p.tiles_y);
    const int z = (int)(task / p.tiles_y);
    const int r0 = ty * p.RH;
    const int nrows = min(p.RH, p.bh - r0);
    const int sft = p.mode == SPIHTB_MODE_PERIODIZATION ? (F / 2 - 1) : 0;
    const int gr0 = 2 * r0 - (F - 2) + sft;
    // rows of the chunk's window all inside the plane (warp-uniform)
    if (gr0 >= 0 && gr0 + 2 * nrows + F - 2 <= p.src_h)
        dwt_fwd_task<Tin, WID, true>(p, tx, ty, z);
    else
        dwt_fwd_task<Tin, WID, false>(p, tx, ty, z);
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
static int launch_level(spihtb_ctx *ctx, FwdK k, int nz)
{
    constexpr int F = Wav<WID>::F;
    constexpr int NOUT = 32 - (F / 2 - 1);
    constexpr int RHMAX = 64;
    k.tiles_x = (k.bw + NOUT - 1) / NOUT;
    // balanced row chunks (no nearly empty tail chunk)
    k.tiles_y = (k.bh + RHMAX - 1) / RHMAX;
    k.RH = (k.bh + k.tiles_y - 1) / k.tiles_y;
    k.ntasks = (long long)k.tiles_x * k.tiles_y * nz;
    const long long nb = (k.ntasks + FW_WARPS - 1) / FW_WARPS;
    if (nb > 0x7fffffffLL) {
        set_error("forward DWT grid too large");
        return SPIHTB_ESHAPE;
    }
    dwt_fwd_level_kernel<Tin, WID><<<(unsigned)nb, FW_WARPS * 32, 0, ctx->stream>>>(k);
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
