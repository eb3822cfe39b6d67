// The rest of PyWavelets' bior family (bior1.1 1.3 1.5 2.4 2.6 2.8 3.1 3.3 3.5 3.7 3.9 derived, 5.5 stored): filter banks on the
// host, and one pair of separable kernels per transform level with the filter length and the taps as launch
// parameters ("remaining bior family"; the reference passes SpihtSettings.wavelet straight
// to pywt.wavedec2 / waverec2, spiht_wrapper.py:163,276).
//
// The three wavelets BASELINE.json names (bior2.2, bior4.4, bior6.8) keep their register / shuffle kernels with the
// taps as immediates (dwt_fwd.cu, dwt_inv.cu); those kernels rely on F/2 being odd (vector alignment of the
// periodization shift, whole-lane halos), which half of this family is not.  Here every thread owns one output
// column and the passes go through float64 scratch planes:
//   forward level:  rows  (axis -2): lo/hi[i][x]   = sum_j f[j] X[ext(2i + s - j)][x]
//                   cols  (axis -1): aa ad da dd[i][k] = sum_j f[j] lo/hi[i][ext(2k + s - j)]   -> quantise, place
//   inverse level:  cols  (axis -1): Xlo/Xhi[i][n] from (aa, ad) / (da, dd)  (dequantised on load, block marks obeyed)
//                   rows  (axis -2): out[r][n]     from Xlo, Xhi
// with s = 1 (reflect, symmetric) or F/2 (periodization) and PyWavelets' order of the axes and of the taps, so the
// float64 results agree with the float64 restatement in the tests to rounding.  Accesses are coalesced along x / k / n; the F row or column
// reads of neighbouring outputs overlap in L1 / L2.  Algorithmic traffic per level is that of the specialised
// kernels plus one write and one read of the two scratch planes.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "wavelets.cuh"

namespace spihtb {

// ---------------------------------------------------------------- filter banks
// Cohen-Daubechies-Feauveau spline pair biorNr.Nd in exact integer arithmetic:
//   rec_lo = sqrt2 (1 + z)^nr / 2^nr
//   dec_lo = sqrt2 (1 + z)^nd  sum_{m < K} C(K-1+m, m) (-1/z + 2 - z)^m 4^(K-1-m)  /  (2^nd 4^(K-1)),  K = (nr + nd) / 2
// and PyWavelets' layout: even common length F; odd-length filters (even nr) get a leading zero, dec_lo centred on
// F/2 and rec_lo on F/2 - 1; even-length ones (odd nr) are both centred on (F-1)/2.
static bool spline_pair(int wid, int *nr, int *nd)
{
    switch (wid) {
        case SPIHTB_WAVELET_BIOR11: *nr = 1; *nd = 1; return true;
        case SPIHTB_WAVELET_BIOR13: *nr = 1; *nd = 3; return true;
        case SPIHTB_WAVELET_BIOR15: *nr = 1; *nd = 5; return true;
        case SPIHTB_WAVELET_BIOR24: *nr = 2; *nd = 4; return true;
        case SPIHTB_WAVELET_BIOR26: *nr = 2; *nd = 6; return true;
        case SPIHTB_WAVELET_BIOR28: *nr = 2; *nd = 8; return true;
        case SPIHTB_WAVELET_BIOR31: *nr = 3; *nd = 1; return true;
        case SPIHTB_WAVELET_BIOR33: *nr = 3; *nd = 3; return true;
        case SPIHTB_WAVELET_BIOR35: *nr = 3; *nd = 5; return true;
        case SPIHTB_WAVELET_BIOR37: *nr = 3; *nd = 7; return true;
        case SPIHTB_WAVELET_BIOR39: *nr = 3; *nd = 9; return true;
    }
    return false;
}

bool wavelet_is_generic(int wid)
{
    int a, b;
    return wid == SPIHTB_WAVELET_BIOR55 || spline_pair(wid, &a, &b);
}

static void poly_mul(std::vector<long long> &a, const std::vector<long long> &b)
{
    std::vector<long long> out(a.size() + b.size() - 1, 0);
    for (size_t i = 0; i < a.size(); ++i)
        for (size_t j = 0; j < b.size(); ++j) out[i + j] += a[i] * b[j];
    a.swap(out);
}

static long long binom(int n, int k)
{
    long long r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

bool generic_wavelet_taps(int wid, int *F_out, double *dec_lo, double *rec_lo)
{
    int nr, nd;
    if (wid == SPIHTB_WAVELET_BIOR55) {   // not a spline pair: PyWavelets' table (9 / 11 taps in a length of 12)
        static const double dl[12] = {0.0, 0.0, 0.03968708834740544, 0.007948108637240322, -0.05446378846823691,
                                      0.34560528195603346, 0.7366601814282105, 0.34560528195603346,
                                      -0.05446378846823691, 0.007948108637240322, 0.03968708834740544, 0.0};
        static const double rl[12] = {0.013456709459118716, -0.002694966880111507, -0.13670658466432914,
                                      -0.09350469740093886, 0.47680326579848425, 0.8995061097486484,
                                      0.47680326579848425, -0.09350469740093886, -0.13670658466432914,
                                      -0.002694966880111507, 0.013456709459118716, 0.0};
        for (int i = 0; i < SPIHTB_GEN_MAXF; ++i) {
            dec_lo[i] = i < 12 ? dl[i] : 0.0;
            rec_lo[i] = i < 12 ? rl[i] : 0.0;
        }
        *F_out = 12;
        return true;
    }
    if (!spline_pair(wid, &nr, &nd)) return false;
    const int K = (nr + nd) / 2;
    const std::vector<long long> one_z = {1, 1}, s4 = {-1, 2, -1};
    std::vector<long long> q(2 * K - 1, 0), pw = {1};
    for (int m = 0; m < K; ++m) {
        long long w4 = 1;
        for (int t = 0; t < K - 1 - m; ++t) w4 *= 4;
        const int off = K - 1 - m;
        for (size_t i = 0; i < pw.size(); ++i) q[off + i] += binom(K - 1 + m, m) * w4 * pw[i];
        poly_mul(pw, s4);
    }
    std::vector<long long> dec = q, rec = {1};
    for (int t = 0; t < nd; ++t) poly_mul(dec, one_z);
    for (int t = 0; t < nr; ++t) poly_mul(rec, one_z);
    const double dden = ldexp(1.0, nd + 2 * (K - 1)), rden = ldexp(1.0, nr);
    const int taps = (int)dec.size();
    const int F = (nr & 1) ? taps : taps + 1;
    const int d0 = (nr & 1) ? 0 : 1;
    const int r0 = (nr & 1) ? (F - (nr + 1)) / 2 : F / 2 - 1 - nr / 2;
    const double s2 = 1.4142135623730951;  // sqrt(2) rounded to float64
    for (int i = 0; i < SPIHTB_GEN_MAXF; ++i) dec_lo[i] = rec_lo[i] = 0.0;
    for (int i = 0; i < taps; ++i) dec_lo[d0 + i] = s2 * (double)dec[i] / dden;   // integer < 2^53, power of two: one rounding
    for (int i = 0; i <= nr; ++i) rec_lo[r0 + i] = s2 * (double)rec[i] / rden;
    *F_out = F;
    return true;
}

// ---------------------------------------------------------------- forward
struct GenTaps {
    int F;
    double lo[SPIHTB_GEN_MAXF], hi[SPIHTB_GEN_MAXF];
};

template <typename Tin>
__device__ __forceinline__ double gen_px(Tin v) { return (double)v; }
template <>
__device__ __forceinline__ double gen_px<uint8_t>(uint8_t v) { return (double)v / 255.0; }  // utils.py:19  im / 255

// rows: lo / hi [z][i][x], i < bh, x < src_w.  A thread takes one column and GEN_R consecutive output rows: the
// 2 GEN_R + F - 2 input rows they need are loaded once into a register window (one output row per thread re-read every
// input row F / 2 times through L2: 12 -> 6.6 ms for level 1 of 256 x 3 x 1024^2 bior3.5 came from dropping the
// extension map, the rest from this).  F is a template parameter so that the window indices and the taps' constant-bank
// offsets are immediates; the taps are still launch parameters.  Accumulation order as before (tap 0 first).
constexpr int GEN_R = 8;
template <typename Tin, int F>
__global__ void __launch_bounds__(256) gen_fwd_rows_kernel(const __grid_constant__ GenTaps t, const Tin *__restrict__ src,
                                                           int src_h, int src_w, int bh, int mode,
                                                           double *__restrict__ lo, double *__restrict__ hi)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * GEN_R;
    if (x >= src_w) return;
    const int s = mode == SPIHTB_MODE_PERIODIZATION ? F / 2 : 1;
    const int z = blockIdx.z;
    const Tin *plane = src + (size_t)z * src_h * src_w + x;
    constexpr int NW = 2 * GEN_R + F - 2;
    const int rbot = 2 * i0 + s - (F - 1);   // input row in window slot 0; output row i0 + u, tap j reads slot F-1 + 2u - j
    double w[NW];
    if (rbot >= 0 && rbot + NW <= src_h) {   // interior (block-uniform): no extension map
        const Tin *q = plane + (size_t)rbot * src_w;
#pragma unroll
        for (int m = 0; m < NW; ++m) w[m] = gen_px<Tin>(q[(size_t)m * src_w]);
    } else {
#pragma unroll
        for (int m = 0; m < NW; ++m) w[m] = gen_px<Tin>(plane[(size_t)ext_index(rbot + m, src_h, mode) * src_w]);
    }
#pragma unroll
    for (int u = 0; u < GEN_R; ++u) {
        if (i0 + u >= bh) break;
        double a = 0.0, d = 0.0;
#pragma unroll
        for (int j = 0; j < F; ++j) {
            a = fma(t.lo[j], w[F - 1 + 2 * u - j], a);
            d = fma(t.hi[j], w[F - 1 + 2 * u - j], d);
        }
        const size_t o = ((size_t)z * bh + i0 + u) * src_w + x;
        lo[o] = a;
        hi[o] = d;
    }
}

struct GenFwdCols {
    const double *lo, *hi;   // [nz][bh][src_w]
    int src_w, bh, bw, mode, C, last;
    double *dst_ll;          // [nz][bh][bw] or null (last level: the LL corner of the array)
    int32_t *coeffs;         // [nz][Hc][Wc]
    int Hc, Wc, sh, sw;
    double scale[8];
    double q;
};

// F is a template parameter so that the tap loop unrolls and the taps become constant-bank operands of the DFMAs
// (with a run-time F the loop spent 4.5 x the instructions of its arithmetic on indexing: ncu, profiles/)
template <int F>
__global__ void __launch_bounds__(256) gen_fwd_cols_kernel(const __grid_constant__ GenTaps t, const __grid_constant__ GenFwdCols p)
{
    // blockDim (64, 4): 64 output columns of four band rows (a band of 517 columns fills 256-wide blocks to 67 %)
    const int k = blockIdx.x * 64 + threadIdx.x;
    const int i = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    const int s = p.mode == SPIHTB_MODE_PERIODIZATION ? F / 2 : 1;
    // the block's segment of the low and the high row goes through shared memory: coalesced loads (the extension map
    // applied once per sample), then the taps read it at a two-sample lane stride
    constexpr int NSEG = 2 * 64 + F - 2;
    __shared__ double s_l[4][NSEG], s_h[4][NSEG];
    const int cb = 2 * (blockIdx.x * 64) + s - (F - 1);   // column of segment slot 0; output 64 bx + x, tap j reads slot F-1 + 2x - j
    if (i < p.bh) {
        const double *rl = p.lo + ((size_t)z * p.bh + i) * p.src_w;
        const double *rh = p.hi + ((size_t)z * p.bh + i) * p.src_w;
        for (int m = threadIdx.x; m < NSEG; m += 64) {
            int c = cb + m;
            if (c < 0 || c >= p.src_w) c = ext_index(c, p.src_w, p.mode);
            s_l[threadIdx.y][m] = rl[c];
            s_h[threadIdx.y][m] = rh[c];
        }
    }
    __syncthreads();
    if (k >= p.bw || i >= p.bh) return;
    double aa = 0.0, ad = 0.0, da = 0.0, dd = 0.0;
    {
        const double *ql = &s_l[threadIdx.y][F - 1 + 2 * threadIdx.x], *qh = &s_h[threadIdx.y][F - 1 + 2 * threadIdx.x];
#pragma unroll
        for (int j = 0; j < F; ++j) {
            const double vl = ql[-j], vh = qh[-j];
            aa = fma(t.lo[j], vl, aa);
            ad = fma(t.hi[j], vl, ad);
            da = fma(t.lo[j], vh, da);
            dd = fma(t.hi[j], vh, dd);
        }
    }
    // spiht_wrapper.py:9-11,167-172: ((m_c * x) * q).astype(int32); coeffs_to_array: 'ad' top right, 'da' bottom left
    const double m = p.scale[z % p.C], q = p.q;
    int32_t *arr = p.coeffs + (size_t)z * p.Hc * p.Wc;
    arr[(size_t)i * p.Wc + p.sw + k] = __double2int_rz((m * ad) * q);
    arr[(size_t)(p.sh + i) * p.Wc + k] = __double2int_rz((m * da) * q);
    arr[(size_t)(p.sh + i) * p.Wc + p.sw + k] = __double2int_rz((m * dd) * q);
    if (p.last)
        arr[(size_t)i * p.Wc + k] = __double2int_rz((m * aa) * q);
    else
        p.dst_ll[((size_t)z * p.bh + i) * p.bw + k] = aa;
}

static int fill_taps_dec(int wid, GenTaps *t)
{
    double dl[SPIHTB_GEN_MAXF], rl[SPIHTB_GEN_MAXF];
    if (!generic_wavelet_taps(wid, &t->F, dl, rl)) {
        set_error("unknown wavelet id %d", wid);
        return SPIHTB_EINVAL;
    }
    for (int i = 0; i < SPIHTB_GEN_MAXF; ++i) {
        t->lo[i] = dl[i];
        t->hi[i] = i < t->F ? (((t->F - 1 - i) & 1) ? -1.0 : 1.0) * rl[i] : 0.0;   // dec_hi[i] = (-1)^(F-1-i) rec_lo[i]
    }
    return SPIHTB_OK;
}

static int fill_taps_rec(int wid, GenTaps *t)
{
    double dl[SPIHTB_GEN_MAXF], rl[SPIHTB_GEN_MAXF];
    if (!generic_wavelet_taps(wid, &t->F, dl, rl)) {
        set_error("unknown wavelet id %d", wid);
        return SPIHTB_EINVAL;
    }
    for (int i = 0; i < SPIHTB_GEN_MAXF; ++i) {
        t->lo[i] = rl[i];
        t->hi[i] = i < t->F ? ((i & 1) ? -1.0 : 1.0) * dl[i] : 0.0;                  // rec_hi[i] = (-1)^i dec_lo[i]
    }
    return SPIHTB_OK;
}

// the row kernels are instantiated per (even) filter length: 2 .. SPIHTB_GEN_MAXF
template <typename Fn>
static bool for_flen(int F, Fn &&fn)
{
    switch (F) {
        case 2: return fn(std::integral_constant<int, 2>{});
        case 4: return fn(std::integral_constant<int, 4>{});
        case 6: return fn(std::integral_constant<int, 6>{});
        case 8: return fn(std::integral_constant<int, 8>{});
        case 10: return fn(std::integral_constant<int, 10>{});
        case 12: return fn(std::integral_constant<int, 12>{});
        case 14: return fn(std::integral_constant<int, 14>{});
        case 16: return fn(std::integral_constant<int, 16>{});
        case 18: return fn(std::integral_constant<int, 18>{});
        case 20: return fn(std::integral_constant<int, 20>{});
        default: return false;
    }
}

int launch_gen_fwd_level(spihtb_ctx *ctx, int wid, const GenFwdLevel &a, int nz)
{
    GenTaps t;
    int rc = fill_taps_dec(wid, &t);
    if (rc) return rc;
    if (a.bh > 65535 || nz > 65535) {
        set_error("generic-wavelet transform: more than 65535 band rows or planes");
        return SPIHTB_ESHAPE;
    }
    const size_t plane = (size_t)nz * a.bh * a.src_w;
    rc = ctx->ensure(ctx->tail, 2 * plane * sizeof(double) + 256);
    if (rc) return rc;
    double *lo = static_cast<double *>(ctx->tail.p), *hi = lo + plane;
    const dim3 g1((a.src_w + 255) / 256, (a.bh + GEN_R - 1) / GEN_R, nz), g2((a.bw + 63) / 64, (a.bh + 3) / 4, nz);
    auto rows = [&](auto fc) {
        constexpr int F = decltype(fc)::value;
        switch (a.src_dtype) {
            case SPIHTB_F64:
                gen_fwd_rows_kernel<double, F><<<g1, 256, 0, ctx->stream>>>(t, static_cast<const double *>(a.src), a.src_h,
                                                                           a.src_w, a.bh, a.mode, lo, hi);
                return true;
            case SPIHTB_F32:
                gen_fwd_rows_kernel<float, F><<<g1, 256, 0, ctx->stream>>>(t, static_cast<const float *>(a.src), a.src_h,
                                                                          a.src_w, a.bh, a.mode, lo, hi);
                return true;
            case SPIHTB_U8:
                gen_fwd_rows_kernel<uint8_t, F><<<g1, 256, 0, ctx->stream>>>(t, static_cast<const uint8_t *>(a.src), a.src_h,
                                                                            a.src_w, a.bh, a.mode, lo, hi);
                return true;
            default:
                return false;
        }
    };
    if (!for_flen(t.F, rows)) {
        set_error("generic-wavelet transform: filter length %d or pixel dtype %d not supported", t.F, a.src_dtype);
        return SPIHTB_EINVAL;
    }
    GenFwdCols p;
    p.lo = lo; p.hi = hi;
    p.src_w = a.src_w; p.bh = a.bh; p.bw = a.bw; p.mode = a.mode; p.C = a.C; p.last = a.last;
    p.dst_ll = a.dst_ll;
    p.coeffs = a.coeffs;
    p.Hc = a.Hc; p.Wc = a.Wc; p.sh = a.sh; p.sw = a.sw;
    for (int c = 0; c < 8; ++c) p.scale[c] = a.scale[c];
    p.q = a.q;
    for_flen(t.F, [&](auto fc) {   // (t.F was accepted by the row pass above)
        gen_fwd_cols_kernel<decltype(fc)::value><<<g2, dim3(64, 4), 0, ctx->stream>>>(t, p);
        return true;
    });
    ctx->launches += 2;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

// ---------------------------------------------------------------- inverse
struct GenInvCols {
    const double *src_a;     // [nz][a_h][a_w] or null (coarsest level: LL corner of the array, dequantised)
    int a_h, a_w;
    const int32_t *coeffs;
    int Hc, Wc, sh, sw, bh, bw, ow, mode, C;
    double rscale[8];
    double rq;
    const uint8_t *blk;      // optional marks of the 64x64 blocks of the array that were written (others read as zero)
    int BH, BW;
    double *xlo, *xhi;       // [nz][bh][ow]
};

// synthesis (non-periodization): rec[n] = sum_t g[t] c[(n + F-2 - t)/2]   for even n+F-2-t, 0 <= k < m
// periodization:                 rec[n] = sum_t g[t] c[((n + F/2-1 - t)/2) mod m]
// A thread takes GEN_C consecutive outputs of one band row: the GEN_C / 2 + F / 2 (or so) coefficients per band they
// need are loaded and dequantised once (one output per thread dequantised every coefficient F times: the kernel was
// bound by the int -> float64 conversions and multiplies, not by memory).  blockDim (64, 4): 64 x GEN_C outputs of four
// band rows, so that the short rows of the coarse levels still fill a block.  Indexing as in gen_inv_rows_kernel.
constexpr int GEN_C = 4;
template <int F, bool PER>
__global__ void __launch_bounds__(256) gen_inv_cols_kernel(const __grid_constant__ GenTaps t, const __grid_constant__ GenInvCols p)
{
    static_assert(GEN_C % 2 == 0, "n0 must be even");
    const int n0 = (blockIdx.x * 64 + threadIdx.x) * GEN_C;
    const int i = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    // no early return: the whole block meets at the barrier before the write-out
    const bool active = n0 < p.ow && i < p.bh;
    // results cross the block through shared memory so that the global stores are whole lines: a thread's four
    // consecutive float64 outputs stored directly are 8-byte pieces at a 32-byte lane stride, four L2 transactions per
    // sector -- the pass was bound by exactly that (6.4 GB of level-1 intermediates).  [row][u][thread], rows padded to 68
    // doubles: conflict-free both ways (writes: consecutive threads; reads: banks 8 u + 2 q per half-warp).
    __shared__ double s_lo[4][GEN_C][68], s_hi[4][GEN_C][68];
    constexpr int OFF = PER ? F / 2 - 1 : F - 2;
    constexpr int DMIN = OFF - (F - 1), DMAX = OFF + GEN_C - 1;
    constexpr int KLO = DMIN >= 0 ? DMIN / 2 : -((-DMIN) / 2);
    constexpr int KHI = DMAX / 2;
    constexpr int NK = KHI - KLO + 1;
    if (active) {
        const double rm = p.rscale[z % p.C], rq = p.rq;
        const int32_t *arr = p.coeffs + (size_t)z * p.Hc * p.Wc;
        const uint8_t *bm = p.blk ? p.blk + (size_t)z * p.BH * p.BW : nullptr;
        // the two array rows this output row reads (details above / beside / below the approximation) and their mark rows
        const int32_t *row_t = arr + (size_t)i * p.Wc, *row_b = arr + (size_t)(p.sh + i) * p.Wc;
        const uint8_t *bm_t = bm ? bm + (size_t)(i >> 6) * p.BW : nullptr;
        const uint8_t *bm_b = bm ? bm + (size_t)((p.sh + i) >> 6) * p.BW : nullptr;
        auto coef = [&](const int32_t *row, const uint8_t *marks, int c) -> double {
            if (marks && !marks[c >> 6]) return 0.0;
            return ((double)row[c] * rm) * rq;   // spiht_wrapper.py:270-274, as in dwt_inv.cu
        };
        const double *row_a = p.src_a ? p.src_a + ((size_t)z * p.a_h + i) * p.a_w : nullptr;
        double aa[NK], ad[NK], da[NK], dd[NK];
    #pragma unroll
        for (int m = 0; m < NK; ++m) {
            int k = n0 / 2 + KLO + m;
            bool in = true;
            if (PER) {
                k %= p.bw;
                if (k < 0) k += p.bw;
            } else {
                in = k >= 0 && k < p.bw;   // taps outside the band contribute nothing
            }
            aa[m] = ad[m] = da[m] = dd[m] = 0.0;
            if (in) {
                aa[m] = row_a ? row_a[k] : ((double)row_t[k] * rm) * rq;
                ad[m] = coef(row_t, bm_t, p.sw + k);
                da[m] = coef(row_b, bm_b, k);
                dd[m] = coef(row_b, bm_b, p.sw + k);
            }
        }
        // outputs n0 + u
#pragma unroll
        for (int u = 0; u < GEN_C; ++u) {
            double lo = 0.0, hi = 0.0;
#pragma unroll
            for (int tt = 0; tt < F; ++tt) {
                if (((OFF + u - tt) & 1) == 0) {
                    const int m = (OFF + u - tt) / 2 - KLO;   // exact: the numerator is even
                    lo = fma(t.lo[tt], aa[m], lo);
                    lo = fma(t.hi[tt], ad[m], lo);
                    hi = fma(t.lo[tt], da[m], hi);
                    hi = fma(t.hi[tt], dd[m], hi);
                }
            }
            s_lo[threadIdx.y][u][threadIdx.x] = lo;
            s_hi[threadIdx.y][u][threadIdx.x] = hi;
        }
    }
    __syncthreads();
    if (i < p.bh) {   // write-out: thread x of a row stores outputs x, x + 64, x + 128, x + 192 of the row's 256
        const int nb = blockIdx.x * 64 * GEN_C;
        const size_t o = ((size_t)z * p.bh + i) * p.ow + nb;
#pragma unroll
        for (int v = 0; v < GEN_C; ++v) {
            const int g = threadIdx.x + 64 * v;
            if (nb + g < p.ow) {
                p.xlo[o + g] = s_lo[threadIdx.y][g % GEN_C][g / GEN_C];
                p.xhi[o + g] = s_hi[threadIdx.y][g % GEN_C][g / GEN_C];
            }
        }
    }
}

// GEN_R consecutive output rows of one column per thread: the GEN_R / 2 + F / 2 (or so) rows of xlo / xhi they need are
// loaded once into registers (one output row per thread re-read every intermediate row F / 2 times through L2).
// Output row r0 + u, tap tt reads band row r0 / 2 + (OFF + u - tt) / 2 for even OFF + u - tt; r0 is even.
template <typename Tout, int F, bool PER>
__global__ void __launch_bounds__(256) gen_inv_rows_kernel(const __grid_constant__ GenTaps t, const double *__restrict__ xlo,
                                                           const double *__restrict__ xhi, int bh, int oh, int ow,
                                                           Tout *__restrict__ dst)
{
    static_assert(GEN_R % 2 == 0, "r0 must be even");
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * GEN_R, z = blockIdx.z;
    if (n >= ow) return;
    constexpr int OFF = PER ? F / 2 - 1 : F - 2;
    constexpr int DMIN = OFF - (F - 1), DMAX = OFF + GEN_R - 1;
    constexpr int KLO = DMIN >= 0 ? DMIN / 2 : -((-DMIN) / 2);   // the smallest even OFF + u - tt, halved
    constexpr int KHI = DMAX / 2;
    constexpr int NK = KHI - KLO + 1;
    double xl[NK], xh[NK];
    const size_t col = (size_t)z * bh * ow + n;
#pragma unroll
    for (int m = 0; m < NK; ++m) {
        int k = r0 / 2 + KLO + m;
        bool in = true;
        if (PER) {
            k %= bh;
            if (k < 0) k += bh;
        } else {
            in = k >= 0 && k < bh;   // taps outside the band contribute nothing
        }
        xl[m] = in ? xlo[col + (size_t)k * ow] : 0.0;
        xh[m] = in ? xhi[col + (size_t)k * ow] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < GEN_R; ++u) {
        if (r0 + u >= oh) break;
        double v = 0.0;
#pragma unroll
        for (int tt = 0; tt < F; ++tt) {
            if (((OFF + u - tt) & 1) == 0) {
                const int m = (OFF + u - tt) / 2 - KLO;   // exact: the numerator is even
                v = fma(t.lo[tt], xl[m], v);
                v = fma(t.hi[tt], xh[m], v);
            }
        }
        dst[((size_t)z * oh + r0 + u) * ow + n] = (Tout)v;
    }
}

int launch_gen_inv_level(spihtb_ctx *ctx, int wid, const GenInvLevel &a, int nz)
{
    GenTaps t;
    int rc = fill_taps_rec(wid, &t);
    if (rc) return rc;
    if (a.bh > 65535 || a.oh > 65535 || nz > 65535) {
        set_error("generic-wavelet transform: more than 65535 rows or planes");
        return SPIHTB_ESHAPE;
    }
    const size_t plane = (size_t)nz * a.bh * a.ow;
    rc = ctx->ensure(ctx->tail, 2 * plane * sizeof(double) + 256);
    if (rc) return rc;
    GenInvCols p;
    p.src_a = a.src_a; p.a_h = a.a_h; p.a_w = a.a_w;
    p.coeffs = a.coeffs;
    p.Hc = a.Hc; p.Wc = a.Wc; p.sh = a.sh; p.sw = a.sw; p.bh = a.bh; p.bw = a.bw; p.ow = a.ow; p.mode = a.mode; p.C = a.C;
    for (int c = 0; c < 8; ++c) p.rscale[c] = a.rscale[c];
    p.rq = a.rq;
    p.blk = a.blk; p.BH = a.BH; p.BW = a.BW;
    p.xlo = static_cast<double *>(ctx->tail.p);
    p.xhi = p.xlo + plane;
    const bool per = a.mode == SPIHTB_MODE_PERIODIZATION;
    {
        const dim3 g1((a.ow + 64 * GEN_C - 1) / (64 * GEN_C), (a.bh + 3) / 4, nz);
        auto cols = [&](auto fc) {
            constexpr int F = decltype(fc)::value;
            if (per)
                gen_inv_cols_kernel<F, true><<<g1, dim3(64, 4), 0, ctx->stream>>>(t, p);
            else
                gen_inv_cols_kernel<F, false><<<g1, dim3(64, 4), 0, ctx->stream>>>(t, p);
            return true;
        };
        if (!for_flen(t.F, cols)) {
            set_error("generic-wavelet transform: filter length %d not supported", t.F);
            return SPIHTB_EINVAL;
        }
    }
    const dim3 g2((a.ow + 255) / 256, (a.oh + GEN_R - 1) / GEN_R, nz);
    auto rows = [&](auto fc) {
        constexpr int F = decltype(fc)::value;
        auto go = [&](auto *out, auto per_c) {
            using Tout = std::remove_pointer_t<decltype(out)>;
            gen_inv_rows_kernel<Tout, F, decltype(per_c)::value><<<g2, 256, 0, ctx->stream>>>(t, p.xlo, p.xhi, a.bh, a.oh,
                                                                                              a.ow, out);
        };
        if (a.out_f32) {
            if (per) go(static_cast<float *>(a.dst), std::true_type{});
            else go(static_cast<float *>(a.dst), std::false_type{});
        } else {
            if (per) go(static_cast<double *>(a.dst), std::true_type{});
            else go(static_cast<double *>(a.dst), std::false_type{});
        }
        return true;
    };
    if (!for_flen(t.F, rows)) {
        set_error("generic-wavelet transform: filter length %d not supported", t.F);
        return SPIHTB_EINVAL;
    }
    ctx->launches += 2;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
