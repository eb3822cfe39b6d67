// Inverse transform: dequantise -> array_to_coeffs -> L-level inverse DWT ->
// (optional IPT->RGB).  Replaces spiht_wrapper.py:259-281 ((x / m_c) / q,
// pywt.array_to_coeffs, pywt.waverec2, colour.convert).
//
// One kernel per level, coarsest first.  The mirror of the forward kernel: one
// warp = one task, a strip of NOUT = 32 - (F/2 - 1) coefficient columns (2 NOUT
// output columns) by RH output row pairs of one (image, channel) plane.  Lane j
// owns coefficient column kc = q0 + S/2 - (F/2-1) + j of the four bands and
// streams down the coefficient rows:
//   axis -1: the output column pair (2q, 2q+1), q = q0 + j, needs the columns
//            of lanes j .. j+F/2-1 (warp shuffle down):
//              X[2q]   = sum_u g[2u]   c[q + S/2 - u],
//              X[2q+1] = sum_u g[2u+1] c[q + S/2 - u]     (g = rec_lo on the
//            approximation / 'da' band, rec_hi on the 'ad' / 'dd' band);
//   axis -2: a register window of F/2 rows of (X_lo, X_hi) gives the output row
//            pair (2p, 2p+1) the same way.
// PyWavelets' idwtn order (axis -1 first), float64.  Details are dequantised
// on load from their place in the coefficient array.  No shared memory, no
// barrier; rows are prefetched PD steps ahead in registers; every lane stores
// two adjacent outputs per row (whole sectors per warp).
// synthesis (non-periodization): rec[n] = sum_t g[t] c[(n + F-2 - t)/2]   for even n+F-2-t, 0 <= k < m
// periodization:                 rec[n] = sum_t g[t] c[((n + F/2-1 - t)/2) mod m]
// waverec2 drops the approximation's trailing row/column when it is one
// longer than the detail band (odd sizes); the strips simply never read it.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "wavelets.cuh"

namespace spihtb {

constexpr int IV_WARPS = 4;

struct InvK {
    const double *src_a;  // approximation planes [nz][a_h][a_w] (null at the coarsest level: LL corner of coeffs)
    int a_h, a_w;
    const int32_t *coeffs;  // [nz][Hc][Wc]
    int Hc, Wc, sh, sw;
    int bh, bw;   // band size
    int oh, ow;   // output size of this level
    void *dst;    // [nz][oh][ow] double scratch, or the final pixels
    int mode, C;
    int tiles_x, tiles_y, RH;  // strips across, chunks of RH output row pairs down
    long long ntasks;
    double rscale[8];  // 1 / m_c
    double rq;         // 1 / q
    // optional marks of the 64x64 blocks of the coefficient array that hold a non-zero coefficient
    // ([nz][BH][BW] bytes, written by the SPIHT decoder): a task none of whose detail blocks is marked reads no
    // detail band and skips their half of the arithmetic (at low rates the finest bands are empty)
    const uint8_t *blk;
    int BH, BW;
    // Fused finest level (FUSE launches: the level-2 launch of spihtb_decode_images, bior2.2, not periodization).
    // l1_any [images]: 0 = the image has no non-zero coefficient in the finest detail bands.  For such an image the
    // task does not store its output (the level-1 approximation) but synthesises the pixels from it on the spot --
    // the finest level is then only an interpolation of the approximation -- and the level-1 launch skips the image
    // (skip_img).  pix: [nz][ph][pw] of TPix.
    const uint8_t *l1_any;
    const uint8_t *skip_img;
    void *pix;
    int ph, pw;
};

template <typename Tout>
struct Out2;
template <>
struct Out2<float> {
    using type = float2;
    static __device__ __forceinline__ float2 make(double a, double b) { return make_float2((float)a, (float)b); }
};
template <>
struct Out2<double> {
    using type = double2;
    static __device__ __forceinline__ double2 make(double a, double b) { return make_double2(a, b); }
};

// spiht_wrapper.py:270-274: (x / m_c) / q.  Evaluated as (x * (1/m_c)) * (1/q): at most 2 ulp from the
// reference's two divisions, far inside the float tolerance of the inverse path.
template <typename Tout, int WID, bool LLQ, bool FUSE = false, typename TPix = float>
__global__ void __launch_bounds__(IV_WARPS * 32) dwt_inv_level_kernel(const InvK p)
{
    constexpr int F = Wav<WID>::F;
    constexpr int HF = F / 2;
    constexpr int NOUT = 32 - (HF - 1);
    // prefetch distance in rows and unroll of the streaming loop (a multiple of HF: static window and queue slots)
    // (6 rows ahead pay off on the latency-bound coarse levels; the float32 output level is issue-bound: 3)
    constexpr int PD_FULL = HF == 3 ? (FUSE ? 3 : (sizeof(Tout) == 8 ? 9 : 3)) : HF;
    constexpr unsigned FULL = 0xffffffffu;
    long long task = (long long)blockIdx.x * IV_WARPS + (threadIdx.x >> 5);
    if (task >= p.ntasks) return;
    const int tx = (int)(task % p.tiles_x);
    task /= p.tiles_x;
    const int ty = (int)(task % p.tiles_y);
    const int z = (int)(task / p.tiles_y);
    if (!FUSE && p.skip_img && p.skip_img[z / p.C] == 0) return;   // the fused level-2 launch wrote this image's pixels
    const int lane = threadIdx.x & 31;
    const bool per = p.mode == SPIHTB_MODE_PERIODIZATION;
    const int S2 = per ? (HF - 1) / 2 : HF - 1;
    const int bh = p.bh, bw = p.bw, Wc = p.Wc;
    // (fused launches: strips and chunks overlap by one output pair, the halo of the second synthesis)
    const int q0 = tx * (FUSE ? NOUT - 1 : NOUT), p0 = ty * (FUSE ? p.RH - 1 : p.RH);
    const int npair = min(p.RH, p.oh / 2 - p0);  // output row pairs of this chunk

    // this lane's coefficient column
    int kc = q0 + S2 - (HF - 1) + lane;
    if (per) {
        kc %= bw;
        if (kc < 0) kc += bw;
    } else {
        kc = min(kc, bw - 1);  // lanes past the band only feed dropped outputs
    }
    const int zimg = z / p.C, zc = z - zimg * p.C;
    const double rm = p.rscale[zc], rq = p.rq;
    bool fuse_img = false;
    if constexpr (FUSE) fuse_img = p.l1_any[zimg] == 0;
    const int32_t *cz = p.coeffs + (size_t)z * p.Hc * Wc;
    const int32_t *c_ad = cz + p.sw + kc;
    const int32_t *c_da = cz + (size_t)p.sh * Wc + kc;
    const int32_t *c_dd = c_da + p.sw;
    const int32_t *c_aa = cz + kc;
    const double *a_aa = LLQ ? nullptr : p.src_a + (size_t)z * p.a_h * p.a_w + kc;
    const int a_w = p.a_w;

    // ---- are the three detail bands zero over everything this task reads?
    bool dz = false;
    if (p.blk) {
        // band rows / columns of the task (clamped like the loads; a wrapped range counts as the whole band)
        int ra = p0 + S2 - (HF - 1), rb = ra + npair + HF - 1;
        int ca = q0 + S2 - (HF - 1), cb = ca + 31;
        if (per && (ra < 0 || rb >= bh)) { ra = 0; rb = bh - 1; }
        if (per && (ca < 0 || cb >= bw)) { ca = 0; cb = bw - 1; }
        ra = max(ra, 0); rb = min(rb, bh - 1);
        ca = max(ca, 0); cb = min(cb, bw - 1);
        const uint8_t *bz = p.blk + (size_t)z * p.BH * p.BW;
        bool any = false;
#pragma unroll
        for (int band = 0; band < 3; ++band) {  // ad (rows 0.., cols sw..), da (rows sh.., cols 0..), dd
            const int ro = band == 0 ? 0 : p.sh, co = band == 1 ? 0 : p.sw;
            const int br0 = (ro + ra) >> 6, br1 = (ro + rb) >> 6, bc0 = (co + ca) >> 6, bc1 = (co + cb) >> 6;
            const int nbc = bc1 - bc0 + 1, nb = (br1 - br0 + 1) * nbc;
            for (int t = lane; t < nb; t += 32) any = any || bz[(size_t)(br0 + t / nbc) * p.BW + bc0 + t % nbc] != 0;
        }
        dz = !__any_sync(FULL, any);
    }

    // the streaming part, compiled twice: with and without the detail bands (warp-uniform choice)
    auto stream = [&](auto dzc) {
    constexpr bool DZ = decltype(dzc)::value;
    // (without the detail bands a row costs one load and two registers: 12 rows ahead on the float32 output level, 6
    // on the float64 ones -- measured)
    constexpr int PD = (DZ && HF == 3) ? (sizeof(Tout) == 4 ? 12 : 6) : PD_FULL;
    constexpr int UN = PD;
    static_assert(UN % HF == 0 && UN % PD == 0, "static slots");
    struct Raw {
        int32_t ad, da, dd, aaq;
        double aa;
    };
    // coefficient row of stream index j (any j >= 0; rows past the chunk's need are clamped / wrapped)
    const int rbase = p0 + S2 - (HF - 1);
    auto load_row = [&](int j, Raw &r) {
        int row = rbase + j;
        if (per) {
            row %= bh;
            if (row < 0) row += bh;
        } else {
            row = min(row, bh - 1);
        }
        const size_t o = (size_t)row * Wc;
        if constexpr (!DZ) {
            r.ad = __ldg(c_ad + o);
            r.da = __ldg(c_da + o);
            r.dd = __ldg(c_dd + o);
        } else {
            r.ad = r.da = r.dd = 0;
        }
        if (LLQ)
            r.aaq = __ldg(c_aa + o);
        else
            r.aa = __ldg(a_aa + (size_t)row * a_w);
    };

    // X[slot][0/1] = (even, odd) output column of the row-synthesised lo / hi signal of stream row j, slot = j % HF
    double xlo[HF][2], xhi[HF][2];
    auto h_synth = [&](const Raw &r, int slot) {
        const double aa = LLQ ? ((double)r.aaq * rm) * rq : r.aa;
        const double ad = ((double)r.ad * rm) * rq;
        const double da = ((double)r.da * rm) * rq;
        const double dd = ((double)r.dd * rm) * rq;
        double le = 0.0, lo = 0.0, he = 0.0, ho = 0.0;
#pragma unroll
        for (int u = 0; u < HF; ++u) {
            const int d = HF - 1 - u;  // lane j + d holds column q + S/2 - u
            const double vaa = d ? __shfl_down_sync(FULL, aa, d) : aa;
            if (Wav<WID>::rec_lo(2 * u) != 0.0) le = fma(Wav<WID>::rec_lo(2 * u), vaa, le);
            if (Wav<WID>::rec_lo(2 * u + 1) != 0.0) lo = fma(Wav<WID>::rec_lo(2 * u + 1), vaa, lo);
            if constexpr (DZ) continue;  // all three detail bands are zero over this task
            const double vad = d ? __shfl_down_sync(FULL, ad, d) : ad;
            const double vda = d ? __shfl_down_sync(FULL, da, d) : da;
            const double vdd = d ? __shfl_down_sync(FULL, dd, d) : dd;
            if (Wav<WID>::rec_lo(2 * u) != 0.0) he = fma(Wav<WID>::rec_lo(2 * u), vda, he);
            if (wav_rec_hi<WID>(2 * u) != 0.0) {
                le = fma(wav_rec_hi<WID>(2 * u), vad, le);
                he = fma(wav_rec_hi<WID>(2 * u), vdd, he);
            }
            if (Wav<WID>::rec_lo(2 * u + 1) != 0.0) ho = fma(Wav<WID>::rec_lo(2 * u + 1), vda, ho);
            if (wav_rec_hi<WID>(2 * u + 1) != 0.0) {
                lo = fma(wav_rec_hi<WID>(2 * u + 1), vad, lo);
                ho = fma(wav_rec_hi<WID>(2 * u + 1), vdd, ho);
            }
        }
        xlo[slot][0] = le;
        xlo[slot][1] = lo;
        xhi[slot][0] = he;
        xhi[slot][1] = ho;
    };

    Raw q[PD];
#pragma unroll
    for (int s = 0; s < PD; ++s) load_row(s, q[s]);
    // rows 0 .. HF-2 of the stream fill the window
    static_assert(PD >= 1, "prefetch");
    int jn = PD;  // next stream row to load
#pragma unroll
    for (int j = 0; j < HF - 1; ++j) {
        h_synth(q[j % PD], j % HF);
        load_row(jn++, q[j % PD]);
    }

    using O2 = Out2<Tout>;
    const int q_out = q0 + lane;
    const bool col_ok = lane < NOUT && q_out < p.ow / 2;
    Tout *dst = static_cast<Tout *>(p.dst) + (size_t)z * p.oh * p.ow + (size_t)(2 * p0) * p.ow + 2 * q_out;
    const int ow = p.ow;

    [[maybe_unused]] double hprev[4] = {0.0, 0.0, 0.0, 0.0};   // fused finest level: column-synthesised row 2P-1
    const int niter = (npair + UN - 1) / UN;
    for (int it = 0; it < niter; ++it) {
#pragma unroll
        for (int u0 = 0; u0 < UN; ++u0) {
            const int i = it * UN + u0;  // output row pair of the chunk (pairs past npair are computed and dropped)
            const int j = u0 + HF - 1;   // stream row completing it (up to a multiple of UN)
            h_synth(q[j % PD], j % HF);
            load_row(jn++, q[j % PD]);
            // axis -2: stream row (j - u) holds coefficient row p + S/2 - u
            double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;  // rows 2p (e) and 2p+1 (o), columns 2q, 2q+1
#pragma unroll
            for (int u = 0; u < HF; ++u) {
                const int slot = ((j - u) % HF + HF) % HF;
                if (Wav<WID>::rec_lo(2 * u) != 0.0) {
                    e0 = fma(Wav<WID>::rec_lo(2 * u), xlo[slot][0], e0);
                    e1 = fma(Wav<WID>::rec_lo(2 * u), xlo[slot][1], e1);
                }
                if (wav_rec_hi<WID>(2 * u) != 0.0 && !DZ) {
                    e0 = fma(wav_rec_hi<WID>(2 * u), xhi[slot][0], e0);
                    e1 = fma(wav_rec_hi<WID>(2 * u), xhi[slot][1], e1);
                }
                if (Wav<WID>::rec_lo(2 * u + 1) != 0.0) {
                    o0 = fma(Wav<WID>::rec_lo(2 * u + 1), xlo[slot][0], o0);
                    o1 = fma(Wav<WID>::rec_lo(2 * u + 1), xlo[slot][1], o1);
                }
                if (wav_rec_hi<WID>(2 * u + 1) != 0.0 && !DZ) {
                    o0 = fma(wav_rec_hi<WID>(2 * u + 1), xhi[slot][0], o0);
                    o1 = fma(wav_rec_hi<WID>(2 * u + 1), xhi[slot][1], o1);
                }
            }
            if constexpr (FUSE) {
                if (fuse_img) {
                    // e0 e1 / o0 o1 are rows 2P, 2P+1 (P = p0 + i), columns 2q', 2q'+1 (q' = q_out) of the level-1
                    // approximation A.  Without detail bands the finest synthesis is, per axis,
                    //   X[2m] = sum_u g[2u] A[m + 2 - u],  X[2m+1] = sum_u g[2u+1] A[m + 2 - u]     (g = rec_lo, F = 6),
                    // evaluated in the order of the level-by-level kernel above (bit-identical pixels).
                    // Columns: pairs m = 2q'-1 (from A[2q'-1 .. 2q'+1]) and m = 2q' (from A[2q' .. 2q'+2]):
                    // pixel columns 4q'-2 .. 4q'+1; the neighbours' columns come by shuffle.
                    auto hsyn = [&](double c0, double c1, double (&h)[4]) {
                        constexpr double g0 = Wav<WID>::rec_lo(0), g1 = Wav<WID>::rec_lo(1), g2 = Wav<WID>::rec_lo(2),
                                         g3 = Wav<WID>::rec_lo(3), g4 = Wav<WID>::rec_lo(4), g5 = Wav<WID>::rec_lo(5);
                        const double nx = __shfl_down_sync(FULL, c0, 1);                        // A[2q'+2]
                        const double pv = (g4 != 0.0 || g5 != 0.0) ? __shfl_up_sync(FULL, c1, 1) : 0.0;   // A[2q'-1]
                        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
                        // m = 2q'-1: A[m+2-u] = c1, c0, pv for u = 0, 1, 2
                        if (g0 != 0.0) a = fma(g0, c1, a);
                        if (g1 != 0.0) b = fma(g1, c1, b);
                        if (g2 != 0.0) a = fma(g2, c0, a);
                        if (g3 != 0.0) b = fma(g3, c0, b);
                        if (g4 != 0.0) a = fma(g4, pv, a);
                        if (g5 != 0.0) b = fma(g5, pv, b);
                        // m = 2q': A[m+2-u] = nx, c1, c0
                        if (g0 != 0.0) c = fma(g0, nx, c);
                        if (g1 != 0.0) d = fma(g1, nx, d);
                        if (g2 != 0.0) c = fma(g2, c1, c);
                        if (g3 != 0.0) d = fma(g3, c1, d);
                        if (g4 != 0.0) c = fma(g4, c0, c);
                        if (g5 != 0.0) d = fma(g5, c0, d);
                        h[0] = a; h[1] = b; h[2] = c; h[3] = d;
                    };
                    double h0[4], h1[4];
                    hsyn(e0, e1, h0);
                    hsyn(o0, o1, h1);
                    // Rows: pair r needs rows r+2, r+1 (, r) of the column-synthesised signal: r = 2P-2 from (h0, hprev),
                    // r = 2P-1 from (h1, h0); rows 2r, 2r+1 of the image.
                    static_assert(!FUSE || (Wav<WID>::rec_lo(4) == 0.0 && Wav<WID>::rec_lo(5) == 0.0 && Wav<WID>::F == 6),
                                  "fused finest level: two-row window (bior2.2)");
                    using P2 = Out2<TPix>;
                    const int P = p0 + i;
                    const int qp = q_out;   // q'
                    const bool lane_ok = lane < NOUT - 1 && qp < p.ow / 2 && i < npair;
                    auto emit = [&](int r, const double (&hb)[4], const double (&ha)[4]) {   // hb = row r+2, ha = row r+1
                        if (!lane_ok || r < 0 || r >= p.ph / 2) return;
                        constexpr double g0 = Wav<WID>::rec_lo(0), g1 = Wav<WID>::rec_lo(1), g2 = Wav<WID>::rec_lo(2),
                                         g3 = Wav<WID>::rec_lo(3);
                        TPix *row = static_cast<TPix *>(p.pix) + ((size_t)z * p.ph + 2 * (size_t)r) * p.pw;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int m = 2 * qp - 1 + half;   // output column pair
                            if (m < 0 || m >= p.pw / 2) continue;
                            double ev0 = 0.0, ev1 = 0.0, od0 = 0.0, od1 = 0.0;
                            if (g0 != 0.0) { ev0 = fma(g0, hb[2 * half], ev0); ev1 = fma(g0, hb[2 * half + 1], ev1); }
                            if (g1 != 0.0) { od0 = fma(g1, hb[2 * half], od0); od1 = fma(g1, hb[2 * half + 1], od1); }
                            if (g2 != 0.0) { ev0 = fma(g2, ha[2 * half], ev0); ev1 = fma(g2, ha[2 * half + 1], ev1); }
                            if (g3 != 0.0) { od0 = fma(g3, ha[2 * half], od0); od1 = fma(g3, ha[2 * half + 1], od1); }
                            *reinterpret_cast<typename P2::type *>(row + 2 * m) = P2::make(ev0, ev1);
                            *reinterpret_cast<typename P2::type *>(row + p.pw + 2 * m) = P2::make(od0, od1);
                        }
                    };
                    if (i > 0) emit(2 * P - 2, h0, hprev);
                    emit(2 * P - 1, h1, h0);
#pragma unroll
                    for (int t = 0; t < 4; ++t) hprev[t] = h1[t];
                } else if (col_ok && i < npair) {
                    *reinterpret_cast<typename O2::type *>(dst) = O2::make(e0, e1);
                    *reinterpret_cast<typename O2::type *>(dst + ow) = O2::make(o0, o1);
                }
            } else if (col_ok && i < npair) {
                *reinterpret_cast<typename O2::type *>(dst) = O2::make(e0, e1);
                *reinterpret_cast<typename O2::type *>(dst + ow) = O2::make(o0, o1);
            }
            dst += 2 * (size_t)ow;
        }
    }
    };
    if (dz)
        stream(std::true_type{});
    else
        stream(std::false_type{});
}

template <int WID>
static int launch_inv_level(spihtb_ctx *ctx, InvK k, int nz, bool out_f32, bool fuse = false, bool pix_f32 = true)
{
    constexpr int F = Wav<WID>::F;
    constexpr int NOUT = 32 - (F / 2 - 1);
    const int opw = k.ow / 2, oph = k.oh / 2;
    // output row pairs per chunk (multiples of the bior2.2 loop unroll): long chunks amortise the window fill on
    // the coarse levels, shorter ones balance the big finest level better (measured)
    const int RHMAX = oph >= 384 ? 60 : 96;
    if (fuse) {
        // strips / chunks advance by one pair less than they compute (see the kernel)
        k.tiles_x = (opw + NOUT - 2) / (NOUT - 1);
        k.tiles_y = std::max(1, (oph - 1 + RHMAX - 2) / (RHMAX - 1));
        k.RH = std::max(2, (oph - 1 + k.tiles_y - 1) / k.tiles_y + 1);
    } else {
        k.tiles_x = (opw + NOUT - 1) / NOUT;
        k.tiles_y = (oph + RHMAX - 1) / RHMAX;
        k.RH = (oph + k.tiles_y - 1) / k.tiles_y;
    }
    k.ntasks = (long long)k.tiles_x * k.tiles_y * nz;
    const long long nb = (k.ntasks + IV_WARPS - 1) / IV_WARPS;
    if (nb > 0x7fffffffLL) {
        set_error("inverse DWT grid too large");
        return SPIHTB_ESHAPE;
    }
    const dim3 grid((unsigned)nb), block(IV_WARPS * 32);
    const bool llq = k.src_a == nullptr;
    if constexpr (WID == SPIHTB_WAVELET_BIOR22) {
        if (fuse) {
            if (llq && pix_f32)
                dwt_inv_level_kernel<double, WID, true, true, float><<<grid, block, 0, ctx->stream>>>(k);
            else if (llq)
                dwt_inv_level_kernel<double, WID, true, true, double><<<grid, block, 0, ctx->stream>>>(k);
            else if (pix_f32)
                dwt_inv_level_kernel<double, WID, false, true, float><<<grid, block, 0, ctx->stream>>>(k);
            else
                dwt_inv_level_kernel<double, WID, false, true, double><<<grid, block, 0, ctx->stream>>>(k);
            ctx->launches++;
            return SPIHTB_OK;
        }
    }
    if (out_f32) {
        if (llq)
            dwt_inv_level_kernel<float, WID, true><<<grid, block, 0, ctx->stream>>>(k);
        else
            dwt_inv_level_kernel<float, WID, false><<<grid, block, 0, ctx->stream>>>(k);
    } else {
        if (llq)
            dwt_inv_level_kernel<double, WID, true><<<grid, block, 0, ctx->stream>>>(k);
        else
            dwt_inv_level_kernel<double, WID, false><<<grid, block, 0, ctx->stream>>>(k);
    }
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
        for (int c = 0; c < 8; ++c) k.rscale[c] = 1.0 / x.scale[c];
        k.rq = 1.0 / x.q;
        k.blk = (l == 0 && x.blk1) ? x.blk1 : x.blk;
        k.BH = (g.enc_h + 63) / 64;
        k.BW = (g.enc_w + 63) / 64;
        bool out_f32 = false;
        if (l == 0) {
            k.dst = color ? ctx->io2.p : pixels_out;
            out_f32 = !color && x.pixel_dtype == SPIHTB_F32;
        } else {
            k.dst = ((L - 1 - l) & 1) ? ctx->tmpb.p : ctx->tmpa.p;
        }
        // images without a coefficient in the finest bands (spihtb_decode_images knows which): their pixels come out of
        // the level-2 launch, and the level-1 launch skips them
        // (float32 pixels only: with float64 output -- and the colour transform works on float64 -- the pixels dominate
        // the traffic either way and the fused launch is the slower one, 42.4 against 39.4 ms at 512 x 3 x 2048^2;
        // SPIHTB_FUSED_INV_F64=1 takes it anyway: tests)
        const bool fusable = x.l1_any && L >= 2 && g.wavelet == SPIHTB_WAVELET_BIOR22 && !per &&
                             getenv("SPIHTB_NO_FUSED_INV") == nullptr &&
                             ((!color && x.pixel_dtype == SPIHTB_F32) || getenv("SPIHTB_FUSED_INV_F64") != nullptr);
        k.l1_any = nullptr;
        k.skip_img = nullptr;
        k.pix = nullptr;
        k.ph = k.pw = 0;
        bool fuse = false, pix_f32 = true;
        if (fusable && l == 1) {
            fuse = true;
            k.l1_any = x.l1_any;
            k.pix = color ? ctx->io2.p : pixels_out;
            pix_f32 = !color && x.pixel_dtype == SPIHTB_F32;
            k.ph = out_len(g.band_h[0]);
            k.pw = out_len(g.band_w[0]);
        }
        if (fusable && l == 0) k.skip_img = x.l1_any;
        switch (g.wavelet) {
            case SPIHTB_WAVELET_BIOR22: rc = launch_inv_level<SPIHTB_WAVELET_BIOR22>(ctx, k, nz, out_f32, fuse, pix_f32); break;
            case SPIHTB_WAVELET_BIOR44: rc = launch_inv_level<SPIHTB_WAVELET_BIOR44>(ctx, k, nz, out_f32); break;
            case SPIHTB_WAVELET_BIOR68: rc = launch_inv_level<SPIHTB_WAVELET_BIOR68>(ctx, k, nz, out_f32); break;
            default:
                if (wavelet_is_generic(g.wavelet)) {   // the rest of the bior family (dwt_gen.cu)
                    GenInvLevel a;
                    a.src_a = k.src_a; a.a_h = k.a_h; a.a_w = k.a_w;
                    a.coeffs = k.coeffs;
                    a.Hc = k.Hc; a.Wc = k.Wc; a.sh = k.sh; a.sw = k.sw; a.bh = k.bh; a.bw = k.bw; a.oh = k.oh; a.ow = k.ow;
                    a.mode = k.mode; a.C = k.C;
                    for (int c = 0; c < 8; ++c) a.rscale[c] = k.rscale[c];
                    a.rq = k.rq;
                    a.blk = k.blk; a.BH = k.BH; a.BW = k.BW;
                    a.dst = k.dst;
                    a.out_f32 = out_f32;
                    rc = launch_gen_inv_level(ctx, g.wavelet, a, nz);
                } else {
                    set_error("unknown wavelet id %d", g.wavelet);
                    rc = SPIHTB_EINVAL;
                }
        }
        if (rc) return rc;
        src = static_cast<const double *>(k.dst);
        a_h = k.oh;
        a_w = k.ow;
    }
    if (color) {
        rc = launch_ipt_to_rgb(ctx, static_cast<const double *>(ctx->io2.p), pixels_out, x.pixel_dtype, img, x.B);
        if (rc) return rc;
    }
    ctx->stage_end(7);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
