// Filter banks and boundary rules of the DWT, as PyWavelets 1.5.0 defines them
// (the reference calls pywt.wavedec2 / waverec2 at spiht_wrapper.py:163,276).
// dec_lo / rec_lo are stored exactly as pywt stores them, zero padding
// included; dec_hi[i] = (-1)^(F-1-i) rec_lo[i], rec_hi[i] = (-1)^i dec_lo[i].
#pragma once
#include "common.cuh"

namespace spihtb {

template <int WID>
struct Wav;

template <>
struct Wav<SPIHTB_WAVELET_BIOR22> {
    static constexpr int F = 6;
    __host__ __device__ static constexpr double dec_lo(int i)
    {
        constexpr double t[F] = {0.0, -0.1767766952966369, 0.3535533905932738, 1.0606601717798214,
                                 0.3535533905932738, -0.1767766952966369};
        return t[i];
    }
    __host__ __device__ static constexpr double rec_lo(int i)
    {
        constexpr double t[F] = {0.0, 0.3535533905932738, 0.7071067811865476, 0.3535533905932738, 0.0, 0.0};
        return t[i];
    }
};

template <>
struct Wav<SPIHTB_WAVELET_BIOR44> {
    static constexpr int F = 10;
    __host__ __device__ static constexpr double dec_lo(int i)
    {
        constexpr double t[F] = {0.0,
                                 0.03782845550726404,
                                 -0.023849465019556843,
                                 -0.11062440441843718,
                                 0.37740285561283066,
                                 0.8526986790088938,
                                 0.37740285561283066,
                                 -0.11062440441843718,
                                 -0.023849465019556843,
                                 0.03782845550726404};
        return t[i];
    }
    __host__ __device__ static constexpr double rec_lo(int i)
    {
        constexpr double t[F] = {0.0,
                                 -0.06453888262869706,
                                 -0.04068941760916406,
                                 0.41809227322161724,
                                 0.7884856164055829,
                                 0.41809227322161724,
                                 -0.04068941760916406,
                                 -0.06453888262869706,
                                 0.0,
                                 0.0};
        return t[i];
    }
};

template <>
struct Wav<SPIHTB_WAVELET_BIOR68> {
    static constexpr int F = 18;
    __host__ __device__ static constexpr double dec_lo(int i)
    {
        constexpr double t[F] = {0.0,
                                 0.0019088317364812906,
                                 -0.0019142861290887667,
                                 -0.016990639867602342,
                                 0.01193456527972926,
                                 0.04973290349094079,
                                 -0.07726317316720414,
                                 -0.09405920349573646,
                                 0.4207962846098268,
                                 0.8259229974584023,
                                 0.4207962846098268,
                                 -0.09405920349573646,
                                 -0.07726317316720414,
                                 0.04973290349094079,
                                 0.01193456527972926,
                                 -0.016990639867602342,
                                 -0.0019142861290887667,
                                 0.0019088317364812906};
        return t[i];
    }
    __host__ __device__ static constexpr double rec_lo(int i)
    {
        constexpr double t[F] = {0.0,
                                 0.0,
                                 0.0,
                                 0.014426282505624435,
                                 0.014467504896790148,
                                 -0.07872200106262882,
                                 -0.04036797903033992,
                                 0.41784910915027457,
                                 0.7589077294536541,
                                 0.41784910915027457,
                                 -0.04036797903033992,
                                 -0.07872200106262882,
                                 0.014467504896790148,
                                 0.014426282505624435,
                                 0.0,
                                 0.0,
                                 0.0,
                                 0.0};
        return t[i];
    }
};

template <int WID>
__host__ __device__ constexpr double wav_dec_hi(int i)
{
    return (((Wav<WID>::F - 1 - i) & 1) ? -1.0 : 1.0) * Wav<WID>::rec_lo(i);
}
template <int WID>
__host__ __device__ constexpr double wav_rec_hi(int i)
{
    return ((i & 1) ? -1.0 : 1.0) * Wav<WID>::dec_lo(i);
}

static inline int wavelet_flen(int wid)
{
    switch (wid) {
        case SPIHTB_WAVELET_BIOR22: return 6;
        case SPIHTB_WAVELET_BIOR44: return 10;
        case SPIHTB_WAVELET_BIOR68: return 18;
        // biorNr.Nd, the rest of the family (dwt_gen.cu): Nr + 2 Nd - 1 taps, padded to an even length
        case SPIHTB_WAVELET_BIOR11: return 2;
        case SPIHTB_WAVELET_BIOR13: return 6;
        case SPIHTB_WAVELET_BIOR15: return 10;
        case SPIHTB_WAVELET_BIOR24: return 10;
        case SPIHTB_WAVELET_BIOR26: return 14;
        case SPIHTB_WAVELET_BIOR28: return 18;
        case SPIHTB_WAVELET_BIOR31: return 4;
        case SPIHTB_WAVELET_BIOR33: return 8;
        case SPIHTB_WAVELET_BIOR35: return 12;
        case SPIHTB_WAVELET_BIOR37: return 16;
        case SPIHTB_WAVELET_BIOR39: return 20;
        case SPIHTB_WAVELET_BIOR55: return 12;
        default: return 0;
    }
}

// Map any integer sample index onto [0, n) for the boundary mode.
// reflect: whole-sample symmetry (period 2n-2); symmetric: half-sample
// (period 2n); periodization: period n + (n & 1), the pad sample repeats the last.
__host__ __device__ __forceinline__ int ext_index(int g, int n, int mode)
{
    if (mode == SPIHTB_MODE_REFLECT) {
        if (n == 1) return 0;
        const int p = 2 * n - 2;
        int m = g % p;
        if (m < 0) m += p;
        return m >= n ? p - m : m;
    } else if (mode == SPIHTB_MODE_SYMMETRIC) {
        const int p = 2 * n;
        int m = g % p;
        if (m < 0) m += p;
        return m >= n ? p - 1 - m : m;
    } else {
        const int p = n + (n & 1);
        int m = g % p;
        if (m < 0) m += p;
        return m >= n ? n - 1 : m;
    }
}

}  // namespace spihtb
