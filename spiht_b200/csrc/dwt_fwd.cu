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
#include <stdlib.h>
#include <string.h>

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

// out-of-line boundary rule: keeps the division sequences out of the streaming loop's code
__device__ __noinline__ int ext_index_slow(int g, int n, int mode) { return ext_index(g, n, mode); }

// spiht_wrapper.py:9-11,167-172: ((m_c * x) * q).astype(int32), truncation toward zero
__device__ __forceinline__ int32_t quantise(double x, double m, double q) { return __double2int_rz((m * x) * q); }
// m == 1.0: (1.0 * x) * q == x * q exactly
__device__ __forceinline__ int32_t quantise1(double x, double q) { return __double2int_rz(x * q); }

// ---- asynchronous global -> shared copies (LDGSTS): the input prefetch queue lives in shared memory,
// costs no registers and no scoreboard slot, and can run many row pairs ahead of the arithmetic
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t saddr, const void *gptr)
{
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async copies 4, 8 or 16 bytes");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(saddr), "l"(gptr), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// NC elements of Tin from this lane's shared-memory slot
template <typename Tin, int NC>
__device__ __forceinline__ void lds_frag(uint32_t saddr, Tin (&v)[NC])
{
    constexpr int BYTES = NC * (int)sizeof(Tin);
    if constexpr (BYTES == 8) {
        uint2 r;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr));
        memcpy(v, &r, 8);
    } else {
        static_assert(BYTES % 16 == 0, "row fragment must be 8 bytes or a multiple of 16");
#pragma unroll
        for (int q = 0; q < BYTES / 16; ++q) {
            uint4 r;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                         : "r"(saddr + 16 * q));
            memcpy(reinterpret_cast<char *>(v) + 16 * q, &r, 16);
        }
    }
}

// A lane holds two adjacent int32 outputs (v0 at p[0], v1 at p[1]; ok0/ok1: inside the band).  When p is
// 8-byte aligned (warp-uniform: lanes are 8 bytes apart) the pair goes out as one store; otherwise every
// lane takes its left neighbour's v1 and stores (v1', v0) at p - 1, and the row's two end columns go out
// alone.  `row_ok` is warp-uniform.
__device__ __forceinline__ void store_pair(int32_t *p, int32_t v0, int32_t v1, bool row_ok, bool ok0, bool ok1,
                                           bool ok_prev, bool ok_next)
{
    if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
        if (row_ok && ok0 && ok1)
            *reinterpret_cast<int2 *>(p) = make_int2(v0, v1);
        else if (row_ok && ok0)
            p[0] = v0;
    } else {
        const int32_t up = __shfl_up_sync(0xffffffffu, v1, 1);
        if (row_ok && ok0) {
            if (ok_prev)
                *reinterpret_cast<int2 *>(p - 1) = make_int2(up, v0);
            else
                p[0] = v0;
        }
        if (row_ok && ok1 && !ok_next) p[1] = v1;
    }
}

template <int WID>
struct FwdCfg {
    static constexpr int F = Wav<WID>::F;
    static constexpr int HF = F / 2;
    static constexpr int NP = F == 6 ? 2 : 1;  // column pairs per lane (register budget: F rows x 2 NP columns)
    static constexpr int NC = 2 * NP;
    static constexpr int NOUT = 32 * NP - (HF - 1);
    // prefetch ring depth in row pairs: a multiple of HF (static slot offsets in the unrolled loop)
    static constexpr int DEPTH = HF == 3 ? 6 : HF;
};

// One warp = one task: a strip of NOUT = 32 NP - (F/2 - 1) output columns by RH output
// rows of one (image, channel) plane.  Lane l owns the NP input column pairs
// (E, O) = (x[2m], x[2m+1]), m = k0 - (F/2-1) + NP l + t, and walks down the rows:
//   axis -2: a register window of F rows of its 2 NP columns gives the row-filtered
//            (lo, hi) values of every column -- no exchange needed;
//   axis -1: out[k] = sum_v f[2v] O[k-v] + f[2v+1] E[k-v] takes the pairs to its left
//            from its own registers or, by warp shuffle, from the lanes before it;
//            the first (F/2-1)/NP lanes only feed their neighbours.
// No shared memory and no barrier: warps are independent, every input sample is
// read from HBM once per strip (the F-2 halo columns hit L1/L2).  Input rows are
// prefetched DEPTH row pairs ahead with cp.async into a per-warp shared-memory ring
// (every lane copies and later reads back only its own fragment, so no barrier is
// needed); each lane writes its band values (approximation as float64 scratch for
// the next level, details quantised).  A lane whose columns lie inside the plane
// (and whose rows are vector-aligned) copies its fragment as one vector; the halo
// lanes of the edge strips go through the boundary map one element at a time.
template <typename Tin, int WID, int NP, bool UNIT_M, bool LAST>
__device__ __forceinline__ void dwt_fwd_task(const FwdK &p, int tx, int ty, int z, uint32_t ring, bool aligned)
{
    constexpr int F = Wav<WID>::F;
    constexpr int HF = F / 2;
    constexpr int NC = 2 * NP;
    static_assert((HF - 1) % NP == 0, "halo must be a whole number of lanes");
    constexpr int HL = (HF - 1) / NP;           // lanes that only feed their neighbours
    constexpr int NOUT = 32 * NP - (HF - 1);
    constexpr int DEPTH = FwdCfg<WID>::DEPTH;
    constexpr int FRAG = NC * (int)sizeof(Tin);  // bytes of one lane's row fragment
    constexpr int SLOT = 2 * 32 * FRAG;          // one row pair of the warp
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int src_h = p.src_h, src_w = p.src_w, mode = p.mode;
    const int sft = mode == SPIHTB_MODE_PERIODIZATION ? (HF - 1) : 0;  // even for every supported wavelet
    const int k0 = tx * NOUT, r0 = ty * p.RH;
    const int nrows = min(p.RH, p.bh - r0);
    const int kf = k0 - (HF - 1) + NP * lane;  // this lane's first pair index = its first output column
    const int gc = 2 * kf + sft;               // its first input column
    const int gr0 = 2 * r0 - (F - 2) + sft;    // first input row of the chunk's window

    const Tin *plane = static_cast<const Tin *>(p.src) + (size_t)z * src_h * src_w;
    const bool lane_vec = aligned && gc >= 0 && gc + NC <= src_w;
    int col[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) col[c] = lane_vec ? gc + c : ext_index(gc + c, src_w, mode);

    // rows gr, gr+1 (warp-uniform; any integer) -> ring slot `slot` (row 0 at +0, row 1 at +32 FRAG)
    const Tin *runp = plane + (ptrdiff_t)gr0 * src_w;  // row gr of the plane while gr is inside it
    int gr = gr0;
    uint32_t my = ring + lane * FRAG;
    asm volatile("" : "+r"(my));  // keep the shared address in a register (not rebuilt from special registers)
    auto fetch_pair = [&](uint32_t slot_off) {
        const Tin *pa = runp, *pb = runp + src_w;
        if (gr < 0 || gr + 1 >= src_h) {  // rare (warp-uniform): boundary rows go through the extension map
            pa = plane + (size_t)ext_index_slow(gr, src_h, mode) * src_w;
            pb = plane + (size_t)ext_index_slow(gr + 1, src_h, mode) * src_w;
        }
        const uint32_t sa = my + slot_off, sb = sa + 32 * FRAG;
        if (lane_vec) {
            constexpr int VB = FRAG >= 16 ? 16 : 8;
#pragma unroll
            for (int o = 0; o < FRAG; o += VB) {
                cp_async<VB>(sa + o, reinterpret_cast<const char *>(pa + gc) + o);
                cp_async<VB>(sb + o, reinterpret_cast<const char *>(pb + gc) + o);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                cp_async<(int)sizeof(Tin)>(sa + c * (int)sizeof(Tin), pa + col[c]);
                cp_async<(int)sizeof(Tin)>(sb + c * (int)sizeof(Tin), pb + col[c]);
            }
        }
        cp_async_commit();
        gr += 2;
        runp += 2 * (ptrdiff_t)src_w;
    };
    // the oldest outstanding row pair (DEPTH are in flight) has landed: read it back
    auto take_pair = [&](uint32_t slot_off, Tin (&a)[NC], Tin (&b)[NC]) {
        cp_async_wait<DEPTH - 1>();
        const uint32_t sa = my + slot_off;
        lds_frag<Tin, NC>(sa, a);
        lds_frag<Tin, NC>(sa + 32 * FRAG, b);
    };

    // row pair j of the chunk (local rows 2j, 2j+1) goes through ring slot j % DEPTH; pairs 0 .. HF-2
    // fill the window, pair HF-1+i completes output row i.  DEPTH pairs are always in flight.
    static_assert(DEPTH >= HF - 1, "ring depth");
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) fetch_pair(j * SLOT);
    // local row li of the chunk lives in window slot li % F; output row i needs rows 2i .. 2i+F-1
    double w[F][NC];
#pragma unroll
    for (int j = 0; j < HF - 1; ++j) {
        Tin a[NC], b[NC];
        take_pair(j * SLOT, a, b);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            w[2 * j][c] = (double)a[c];
            w[2 * j + 1][c] = (double)b[c];
        }
        fetch_pair(j * SLOT);  // pair j + DEPTH
    }
    uint32_t slot_off = ((HF - 1) % DEPTH) * SLOT;  // ring slot of the next pair to take

    const int zc = z % p.C;
    const double m = UNIT_M ? 1.0 : p.scale[zc], qs = p.q;
    bool col_ok[NP];
#pragma unroll
    for (int t = 0; t < NP; ++t) col_ok[t] = lane >= HL && kf + t < p.bw;
    // does the lane before / after this one store its last / first column?  (NP == 2 pair stores)
    const bool ok_prev = lane >= 1 && lane - 1 >= HL && kf - 1 < p.bw;
    const bool ok_next = lane < 31 && lane + 1 >= HL && kf + NP < p.bw;
    const int Wc = p.Wc;
    int32_t *cz = p.coeffs + (size_t)z * p.Hc * Wc;
    int32_t *p_aa = cz + (ptrdiff_t)r0 * Wc + kf;           // LL corner (last level only)
    int32_t *p_ad = p_aa + p.sw;                            // rows lo, cols hi: top right
    int32_t *p_da = cz + (ptrdiff_t)(p.sh + r0) * Wc + kf;  // rows hi, cols lo: bottom left
    int32_t *p_dd = p_da + p.sw;
    constexpr bool ll_scratch = !LAST;
    double *p_ll = ll_scratch ? p.dst_ll + ((size_t)z * p.bh + r0) * p.bw + kf : nullptr;
    const int bw = p.bw;

    // the loop body is unrolled over HF rows so that the window slots are static
    const int niter = (nrows + HF - 1) / HF;
    for (int it = 0; it < niter; ++it) {
#pragma unroll
        for (int u = 0; u < HF; ++u) {
            const int i = it * HF + u;  // rows past nrows (fewer than HF) are computed and dropped
            {
                Tin a[NC], b[NC];
                take_pair(slot_off, a, b);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    w[(2 * u + F - 2) % F][c] = (double)a[c];
                    w[(2 * u + F - 1) % F][c] = (double)b[c];
                }
            }
            fetch_pair(slot_off);
            slot_off = slot_off + SLOT == DEPTH * SLOT ? 0u : slot_off + SLOT;
            // axis -2: tap j multiplies local row 2i + F-1 - j
            double lo[NC], hi[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) lo[c] = hi[c] = 0.0;
#pragma unroll
            for (int j = 0; j < F; ++j) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const double v = w[(2 * u + F - 1 - j) % F][c];
                    if (Wav<WID>::dec_lo(j) != 0.0) lo[c] = fma(Wav<WID>::dec_lo(j), v, lo[c]);
                    if (wav_dec_hi<WID>(j) != 0.0) hi[c] = fma(wav_dec_hi<WID>(j), v, hi[c]);
                }
            }
            // pairs at offsets -(HF-1) .. NP-1 from this lane's first pair: own registers or the lanes before
            double xlo[HF - 1 + NP][2], xhi[HF - 1 + NP][2];
#pragma unroll
            for (int j = 0; j < HF - 1 + NP; ++j) {
                const int s = j - (HF - 1);
                if (s >= 0) {
                    xlo[j][0] = lo[2 * s];
                    xlo[j][1] = lo[2 * s + 1];
                    xhi[j][0] = hi[2 * s];
                    xhi[j][1] = hi[2 * s + 1];
                } else {
                    const int d = (-s + NP - 1) / NP, idx = s + d * NP;
#pragma unroll
                    for (int eo = 0; eo < 2; ++eo) {
                        // tap 2v multiplies O (eo = 1), tap 2v+1 multiplies E (eo = 0); skip all-zero taps
                        bool used = false;
#pragma unroll
                        for (int t = 0; t < NP; ++t) {
                            const int v = t - s;
                            if (v >= 0 && v < HF) {
                                const int tap = eo ? 2 * v : 2 * v + 1;
                                used = used || Wav<WID>::dec_lo(tap) != 0.0 || wav_dec_hi<WID>(tap) != 0.0;
                            }
                        }
                        xlo[j][eo] = used ? __shfl_up_sync(FULL, lo[2 * idx + eo], d) : 0.0;
                        xhi[j][eo] = used ? __shfl_up_sync(FULL, hi[2 * idx + eo], d) : 0.0;
                    }
                }
            }
            const bool row_ok = i < nrows;
            int32_t q_ad[NP], q_da[NP], q_dd[NP], q_aa[NP];
            double f_aa[NP];
#pragma unroll
            for (int t = 0; t < NP; ++t) {
                // axis -1: tap 2v multiplies O[k-v], tap 2v+1 multiplies E[k-v]
                double aa = 0.0, ad = 0.0, da = 0.0, dd = 0.0;  // this lane's output column kf + t
#pragma unroll
                for (int v = 0; v < HF; ++v) {
                    const int j = t - v + (HF - 1);
                    if (Wav<WID>::dec_lo(2 * v) != 0.0) {
                        aa = fma(Wav<WID>::dec_lo(2 * v), xlo[j][1], aa);
                        da = fma(Wav<WID>::dec_lo(2 * v), xhi[j][1], da);
                    }
                    if (wav_dec_hi<WID>(2 * v) != 0.0) {
                        ad = fma(wav_dec_hi<WID>(2 * v), xlo[j][1], ad);
                        dd = fma(wav_dec_hi<WID>(2 * v), xhi[j][1], dd);
                    }
                    if (Wav<WID>::dec_lo(2 * v + 1) != 0.0) {
                        aa = fma(Wav<WID>::dec_lo(2 * v + 1), xlo[j][0], aa);
                        da = fma(Wav<WID>::dec_lo(2 * v + 1), xhi[j][0], da);
                    }
                    if (wav_dec_hi<WID>(2 * v + 1) != 0.0) {
                        ad = fma(wav_dec_hi<WID>(2 * v + 1), xlo[j][0], ad);
                        dd = fma(wav_dec_hi<WID>(2 * v + 1), xhi[j][0], dd);
                    }
                }
                if (!UNIT_M) {  // (1.0 * x) * q == x * q exactly
                    ad *= m;
                    da *= m;
                    dd *= m;
                    if (!ll_scratch) aa *= m;
                }
                q_ad[t] = quantise1(ad, qs);
                q_da[t] = quantise1(da, qs);
                q_dd[t] = quantise1(dd, qs);
                q_aa[t] = ll_scratch ? 0 : quantise1(aa, qs);
                f_aa[t] = aa;
            }
            if constexpr (NP == 2) {
                // both columns of a lane go out as one 8-byte store (whole 32-byte sectors per warp)
                store_pair(p_ad, q_ad[0], q_ad[1], row_ok, col_ok[0], col_ok[1], ok_prev, ok_next);
                store_pair(p_da, q_da[0], q_da[1], row_ok, col_ok[0], col_ok[1], ok_prev, ok_next);
                store_pair(p_dd, q_dd[0], q_dd[1], row_ok, col_ok[0], col_ok[1], ok_prev, ok_next);
                if (ll_scratch) {
                    if (row_ok && col_ok[0] && col_ok[1] && (reinterpret_cast<uintptr_t>(p_ll) & 15) == 0) {
                        *reinterpret_cast<double2 *>(p_ll) = make_double2(f_aa[0], f_aa[1]);
                    } else {
                        if (row_ok && col_ok[0]) p_ll[0] = f_aa[0];
                        if (row_ok && col_ok[1]) p_ll[1] = f_aa[1];
                    }
                } else {
                    store_pair(p_aa, q_aa[0], q_aa[1], row_ok, col_ok[0], col_ok[1], ok_prev, ok_next);
                }
            } else {
#pragma unroll
                for (int t = 0; t < NP; ++t) {
                    if (row_ok && col_ok[t]) {
                        p_ad[t] = q_ad[t];
                        p_da[t] = q_da[t];
                        p_dd[t] = q_dd[t];
                        if (ll_scratch)
                            p_ll[t] = f_aa[t];
                        else
                            p_aa[t] = q_aa[t];
                    }
                }
            }
            p_aa += Wc;
            p_ad += Wc;
            p_da += Wc;
            p_dd += Wc;
            if (ll_scratch) p_ll += bw;
        }
    }
}

template <typename Tin, int WID, int NP, bool UNIT_M, bool LAST>
__global__ void __launch_bounds__(FW_WARPS * 32) dwt_fwd_level_kernel(const FwdK p)
{
    constexpr int F = Wav<WID>::F;
    constexpr int NOUT = 32 * NP - (F / 2 - 1);
    constexpr int DEPTH = FwdCfg<WID>::DEPTH;
    constexpr int WARP_RING = DEPTH * 2 * 32 * 2 * NP * (int)sizeof(Tin);
    __shared__ __align__(16) unsigned char s_ring[FW_WARPS * WARP_RING];
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_ring) + (threadIdx.x >> 5) * WARP_RING;
    long long task = (long long)blockIdx.x * FW_WARPS + (threadIdx.x >> 5);
    if (task >= p.ntasks) return;
    const int tx = (int)(task % p.tiles_x);
    task /= p.tiles_x;
    const int ty = (int)(task % p.tiles_y);
    const int z = (int)(task / p.tiles_y);
    // every row fragment of the strip vector-aligned (warp-uniform)
    const int sft = p.mode == SPIHTB_MODE_PERIODIZATION ? (F / 2 - 1) : 0;
    const int gc_first = 2 * (tx * NOUT - (F / 2 - 1)) + sft;
    constexpr int VB = 2 * NP * (int)sizeof(Tin) >= 16 ? 16 : 8;  // vector bytes
    const long long base = (long long)reinterpret_cast<uintptr_t>(p.src) +
                           (long long)z * p.src_h * p.src_w * (long long)sizeof(Tin);
    const bool aligned = ((size_t)p.src_w * sizeof(Tin)) % VB == 0 &&
                         (base + (long long)gc_first * (long long)sizeof(Tin)) % VB == 0;
    dwt_fwd_task<Tin, WID, NP, UNIT_M, LAST>(p, tx, ty, z, ring, aligned);
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
    constexpr int NP = FwdCfg<WID>::NP;
    constexpr int NOUT = FwdCfg<WID>::NOUT;
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
    bool unit = true;
    for (int c = 0; c < k.C && c < 8; ++c) unit = unit && k.scale[c] == 1.0;
    const dim3 grid((unsigned)nb), block(FW_WARPS * 32);
    if (unit && k.last)
        dwt_fwd_level_kernel<Tin, WID, NP, true, true><<<grid, block, 0, ctx->stream>>>(k);
    else if (unit)
        dwt_fwd_level_kernel<Tin, WID, NP, true, false><<<grid, block, 0, ctx->stream>>>(k);
    else if (k.last)
        dwt_fwd_level_kernel<Tin, WID, NP, false, true><<<grid, block, 0, ctx->stream>>>(k);
    else
        dwt_fwd_level_kernel<Tin, WID, NP, false, false><<<grid, block, 0, ctx->stream>>>(k);
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

    // One launch per level over the whole batch.  (Running the first levels image group by image group,
    // so that a group's float64 approximation planes stay in L2 for the next level, was measured slower
    // on B200: the short launches leave the SMs idle at every kernel boundary.)
    auto run_level = [&](int l, int z0, int nzg) -> int {
        const bool in_f64 = l > 0 || x.pixel_dtype == SPIHTB_F64 || x.color == SPIHTB_COLOR_IPT;
        const size_t esz = in_f64 ? sizeof(double) : sizeof(float);
        FwdK k;
        const void *lsrc = l == 0 ? src : (((l - 1) & 1) ? ctx->tmpb.p : ctx->tmpa.p);
        k.src_h = g.in_h[l];
        k.src_w = g.in_w[l];
        k.src = static_cast<const char *>(lsrc) + (size_t)z0 * k.src_h * k.src_w * esz;
        k.bh = g.band_h[l];
        k.bw = g.band_w[l];
        k.last = (l == L - 1);
        k.dst_ll = k.last ? nullptr
                          : static_cast<double *>((l & 1) ? ctx->tmpb.p : ctx->tmpa.p) + (size_t)z0 * k.bh * k.bw;
        k.Hc = g.enc_h;
        k.Wc = g.enc_w;
        k.coeffs = coeffs + (size_t)z0 * k.Hc * k.Wc;
        k.sh = g.off_h[l];
        k.sw = g.off_w[l];
        k.mode = g.mode;
        k.C = x.C;
        for (int c = 0; c < 8; ++c) k.scale[c] = x.scale[c];
        k.q = x.q;
        const int st = l == 0 ? 0 : 1;
        ctx->stage_begin(st);
        const int r = in_f64 ? launch_level_w<double>(ctx, g.wavelet, k, nzg)
                             : launch_level_w<float>(ctx, g.wavelet, k, nzg);
        ctx->stage_end(st);
        return r;
    };
    for (int l = 0; l < L; ++l) {
        rc = run_level(l, 0, nz);
        if (rc) return rc;
    }
    (void)src_is_f64;
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
        if (any) {
            ctx->stage_begin(1);
            gap_fill_kernel<<<dim3(8, std::min(nz, 65535)), 256, 0, ctx->stream>>>(gk);
            ctx->launches++;
            ctx->stage_end(1);
        }
    }
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
