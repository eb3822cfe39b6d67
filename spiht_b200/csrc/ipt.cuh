// Per-pixel RGB <-> IPT arithmetic (float64), shared by the stand-alone colour kernels (color.cu) and the
// transform kernels that fuse the conversion into their level-1 loads / stores.
//   colour.convert(x, 'RGB', 'IPT') (colour-science 0.4.4, called at spiht/color_models.py:12):
//   linear sRGB -> CIE XYZ (4-digit IEC 61966-2-1 matrix; no CCTF decoding, D65 -> D65 adaptation is the
//   identity) -> LMS -> sign(x) |x|^0.43 -> IPT (Ebner & Fairchild 1998).
//   The way back inverts every matrix in float64: colour-science 0.4.4 defines MATRIX_XYZ_TO_sRGB as
//   np.linalg.inv(MATRIX_sRGB_TO_XYZ), not the rounded 4-digit inverse of releases before 0.4.
#pragma once
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

#ifdef __CUDACC__
__device__ __forceinline__ double ipt_spow(double a, double e) { return a == 0.0 ? 0.0 : copysign(pow(fabs(a), e), a); }

__device__ __forceinline__ void rgb_to_ipt_px(double R, double G, double B, double &I, double &P, double &T)
{
    const double X = 0.4124 * R + 0.3576 * G + 0.1805 * B;
    const double Y = 0.2126 * R + 0.7152 * G + 0.0722 * B;
    const double Z = 0.0193 * R + 0.1192 * G + 0.9505 * B;
    const double L = ipt_spow(0.4002 * X + 0.7075 * Y + -0.0807 * Z, 0.43);
    const double M = ipt_spow(-0.2280 * X + 1.1500 * Y + 0.0612 * Z, 0.43);
    const double S = ipt_spow(0.0 * X + 0.0 * Y + 0.9184 * Z, 0.43);
    I = 0.4000 * L + 0.4000 * M + 0.2000 * S;
    P = 4.4550 * L + -4.8510 * M + 0.3960 * S;
    T = 0.8056 * L + 0.3572 * M + -1.1628 * S;
}

__device__ __forceinline__ void ipt_to_rgb_px(const IptInv &mi, double I, double P, double T, double &R, double &G,
                                              double &B)
{
    const double e = 1.0 / 0.43;
    const double L = ipt_spow(mi.ipt2lms[0] * I + mi.ipt2lms[1] * P + mi.ipt2lms[2] * T, e);
    const double M = ipt_spow(mi.ipt2lms[3] * I + mi.ipt2lms[4] * P + mi.ipt2lms[5] * T, e);
    const double S = ipt_spow(mi.ipt2lms[6] * I + mi.ipt2lms[7] * P + mi.ipt2lms[8] * T, e);
    const double X = mi.lms2xyz[0] * L + mi.lms2xyz[1] * M + mi.lms2xyz[2] * S;
    const double Y = mi.lms2xyz[3] * L + mi.lms2xyz[4] * M + mi.lms2xyz[5] * S;
    const double Z = mi.lms2xyz[6] * L + mi.lms2xyz[7] * M + mi.lms2xyz[8] * S;
    R = mi.xyz2rgb[0] * X + mi.xyz2rgb[1] * Y + mi.xyz2rgb[2] * Z;
    G = mi.xyz2rgb[3] * X + mi.xyz2rgb[4] * Y + mi.xyz2rgb[5] * Z;
    B = mi.xyz2rgb[6] * X + mi.xyz2rgb[7] * Y + mi.xyz2rgb[8] * Z;
}
#endif

}  // namespace spihtb
