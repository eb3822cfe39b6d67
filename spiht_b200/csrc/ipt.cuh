// Per-pixel RGB <-> IPT arithmetic (float64) of the colour kernels (color.cu), which the image path runs as a pass of
// its own in front of / behind the transform.
//   colour.convert(x, 'RGB', 'IPT') (colour-science 0.4.4, called at spiht/color_models.py:12):
//   linear sRGB -> CIE XYZ (4-digit IEC 61966-2-1 matrix; no CCTF decoding, D65 -> D65 adaptation is the
//   identity) -> LMS -> sign(x) |x|^0.43 -> IPT (Ebner & Fairchild 1998).
//   The way back inverts every matrix in float64: colour-science 0.4.4 defines MATRIX_XYZ_TO_sRGB as
//   np.linalg.inv(MATRIX_sRGB_TO_XYZ), not the rounded 4-digit inverse of releases before 0.4.
#pragma once
#include <math.h>

#include "common.cuh"

namespace spihtb {

struct IptInv {
    double ipt2lms[9];  // inv(M_LMS'->IPT)
    double lms2xyz[9];  // inv(M_XYZ->LMS)
    double xyz2rgb[9];  // inv(M_sRGB->XYZ)
};

static inline void inv3(const double a[9], double r[9])
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

static inline IptInv make_ipt_inv()
{
    IptInv mi;
    const double lms2ipt[9] = {0.4000, 0.4000, 0.2000, 4.4550, -4.8510, 0.3960, 0.8056, 0.3572, -1.1628};
    const double xyz2lms[9] = {0.4002, 0.7075, -0.0807, -0.2280, 1.1500, 0.0612, 0.0, 0.0, 0.9184};
    const double rgb2xyz[9] = {0.4124, 0.3576, 0.1805, 0.2126, 0.7152, 0.0722, 0.0193, 0.1192, 0.9505};
    inv3(lms2ipt, mi.ipt2lms);
    inv3(xyz2lms, mi.lms2xyz);
    inv3(rgb2xyz, mi.xyz2rgb);
    return mi;
}

// ---- |x|^0.43 and |x|^(1/0.43) without the general pow() ------------------------------------------------
// The colour transform is three powers per pixel; CUDA's double-precision pow() (~200 instructions) made the
// stand-alone RGB -> IPT pass of 512 x 3 x 2048^2 take 56 ms against 22 ms for the whole DWT.  For a fixed exponent
// p:  x = m 2^e, m in [1, 2);  k = top four mantissa bits, c_k = 1 + (k + 1/2) / 16, r = m / c_k - 1, |r| <= 1/33;
//     x^p = 2^(p e) * c_k^p * (1 + r)^p,   (1 + r)^p = sum_n C(p, n) r^n  (binomial series, 10 terms: < 1e-17).
// 2^(p e) (e in [-64, 16)) and c_k^p, 1 / c_k are tables computed on the host in long double and rounded once;
// the result is within a few ulp of the correctly rounded power (tests/test_gpu_transform.py checks 4e-15 absolute
// on [0, 1] against numpy).  Exponents outside the table (|x| < 2^-64 or >= 2^16) take pow().
struct PowTab {
    double p;
    double expo[80];     // 2^(p e), e = -64 .. 15
    double inv_c[16];    // 1 / c_k
    double c_p[16];      // c_k^p
    double coef[10];     // C(p, n), n = 1 .. 10
};
struct IptPowTabs {
    PowTab fwd, inv;     // p = 0.43, p = 1 / 0.43
};

static inline void fill_pow_tab(PowTab &t, double p)
{
    t.p = p;
    const long double pl = (long double)p;
    for (int e = -64; e < 16; ++e) t.expo[e + 64] = (double)powl(2.0L, pl * (long double)e);
    for (int k = 0; k < 16; ++k) {
        const long double c = 1.0L + ((long double)k + 0.5L) / 16.0L;
        t.inv_c[k] = (double)(1.0L / c);
        t.c_p[k] = (double)powl(c, pl);
    }
    long double b = 1.0L;
    for (int n = 1; n <= 10; ++n) {
        b = b * (pl - (long double)(n - 1)) / (long double)n;
        t.coef[n - 1] = (double)b;
    }
}

#ifdef __CUDACC__
// one copy per translation unit that includes this header (the library is built without relocatable device
// code); each such unit uploads its copy with ipt_upload_tables_tu() when a context is created
static __constant__ IptPowTabs c_ipt_pow;
static inline int ipt_upload_tables_tu()
{
    IptPowTabs t;
    fill_pow_tab(t.fwd, 0.43);
    fill_pow_tab(t.inv, 1.0 / 0.43);
    SPIHTB_CUDA_CHECK(cudaMemcpyToSymbol(c_ipt_pow, &t, sizeof(t)));
    return SPIHTB_OK;
}

__device__ __forceinline__ double ipt_fast_pow(double a, const PowTab &t)   // a > 0
{
    const long long bits = __double_as_longlong(a);
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    if (e < -64 || e > 15) return pow(a, t.p);
    const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    const int k = (int)((bits >> 48) & 15);
    const double r = fma(m, t.inv_c[k], -1.0);
    // 1 + c1 r + ... + c10 r^10 by Estrin's scheme (dependency depth 4 instead of Horner's 10)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = fma(t.coef[0], r, 1.0), p23 = fma(t.coef[2], r, t.coef[1]), p45 = fma(t.coef[4], r, t.coef[3]),
                 p67 = fma(t.coef[6], r, t.coef[5]), p89 = fma(t.coef[8], r, t.coef[7]);
    const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(t.coef[9], r2, p89);
    const double s = fma(q2, r8, fma(q1, r4, q0));
    return (t.expo[e + 64] * t.c_p[k]) * s;
}
// sign(a) |a|^p.  `t`: the table in SHARED memory (ipt_stage_table): the interval and exponent indices differ from
// lane to lane, and constant memory serialises divergent indices.
__device__ __forceinline__ double ipt_spow(double a, const PowTab &t)
{
    if (a == 0.0) return 0.0;
    return copysign(ipt_fast_pow(fabs(a), t), a);
}
// copy the forward (p = 0.43) or inverse (p = 1 / 0.43) table from constant to shared memory (whole CTA; barrier)
__device__ __forceinline__ void ipt_stage_table(PowTab *dst, bool inv)
{
    const double *src = reinterpret_cast<const double *>(inv ? &c_ipt_pow.inv : &c_ipt_pow.fwd);
    double *d = reinterpret_cast<double *>(dst);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    for (int i = tid; i < (int)(sizeof(PowTab) / sizeof(double)); i += nthreads) d[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ void rgb_to_ipt_px(const PowTab &pt, double R, double G, double B, double &I, double &P,
                                              double &T)
{
    const double X = 0.4124 * R + 0.3576 * G + 0.1805 * B;
    const double Y = 0.2126 * R + 0.7152 * G + 0.0722 * B;
    const double Z = 0.0193 * R + 0.1192 * G + 0.9505 * B;
    const double L = ipt_spow(0.4002 * X + 0.7075 * Y + -0.0807 * Z, pt);
    const double M = ipt_spow(-0.2280 * X + 1.1500 * Y + 0.0612 * Z, pt);
    const double S = ipt_spow(0.0 * X + 0.0 * Y + 0.9184 * Z, pt);
    I = 0.4000 * L + 0.4000 * M + 0.2000 * S;
    P = 4.4550 * L + -4.8510 * M + 0.3960 * S;
    T = 0.8056 * L + 0.3572 * M + -1.1628 * S;
}

__device__ __forceinline__ void ipt_to_rgb_px(const PowTab &pt, const IptInv &mi, double I, double P, double T,
                                              double &R, double &G, double &B)
{
    const double L = ipt_spow(mi.ipt2lms[0] * I + mi.ipt2lms[1] * P + mi.ipt2lms[2] * T, pt);
    const double M = ipt_spow(mi.ipt2lms[3] * I + mi.ipt2lms[4] * P + mi.ipt2lms[5] * T, pt);
    const double S = ipt_spow(mi.ipt2lms[6] * I + mi.ipt2lms[7] * P + mi.ipt2lms[8] * T, pt);
    const double X = mi.lms2xyz[0] * L + mi.lms2xyz[1] * M + mi.lms2xyz[2] * S;
    const double Y = mi.lms2xyz[3] * L + mi.lms2xyz[4] * M + mi.lms2xyz[5] * S;
    const double Z = mi.lms2xyz[6] * L + mi.lms2xyz[7] * M + mi.lms2xyz[8] * S;
    R = mi.xyz2rgb[0] * X + mi.xyz2rgb[1] * Y + mi.xyz2rgb[2] * Z;
    G = mi.xyz2rgb[3] * X + mi.xyz2rgb[4] * Y + mi.xyz2rgb[5] * Z;
    B = mi.xyz2rgb[6] * X + mi.xyz2rgb[7] * Y + mi.xyz2rgb[8] * Z;
}
#endif

}  // namespace spihtb
