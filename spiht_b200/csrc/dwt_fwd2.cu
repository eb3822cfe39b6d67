// Forward transform, levels 1 and 2 fused in one kernel with TMA-staged input tiles.
//
// The level-by-level kernel (dwt_fwd.cu) writes the float64 level-1 approximation to HBM and reads it back
// for level 2: 13.2 MB per 1024^2 RGB image of traffic that carries no algorithmic byte.  Here a CTA owns a
// column strip of one (image, channel) plane and streams down its rows; both levels are computed from one
// pass over the pixels and the level-1 approximation never leaves shared memory.
//
//   input     TMA (cp.async.bulk.tensor, one 1-row box per plane row, mbarrier completion) into a ring of
//             stages of 8 rows; boundary rows are ordinary row coordinates (the extension map is applied to
//             the coordinate), boundary columns are read through a per-thread column map
//   V1        a thread owns one input column: register window of F rows, row filter -> (lo, hi) rows in smem
//   H1        a thread owns one level-1 output column of the lo or the hi row: column filter from smem pairs;
//             aa -> ring of level-1 approximation rows (smem, float64); ad / da / dd quantised -> int32 tile
//   OUT1      tile rows -> coefficient array as aligned 16-byte stores; 2x2 cells of the tile -> pyramid bytes
//   V2 / H2 / OUT2   the same one level down, from the ring; aa2 -> float64 scratch for level 3
//
// Arithmetic (tap order, fma accumulation, (m x) q truncation) is identical to dwt_fwd.cu, so both paths give
// bit-identical coefficient arrays (tests/test_gpu_fused12.py).
//
// Index conventions, one axis (F taps, sft = F/2 - 1 in periodization mode, else 0):
//   out[k] = sum_j f[j] x_ext[2k + 1 + sft - j]            window [2k + sft - (F-2), 2k + sft + 1]
// Columns of a strip: level-2 outputs m in [M0 - 1, M1) (column M0 - 1 only completes the first pyramid cell);
// level-1 "virtual" columns kv = KV0 + c1, KV0 = 2 (M0-1) + sft - (F-2), c1 in [0, nk), nk = 2 (M1-M0+1) + F-2;
// input virtual columns xv = XV0 + t, XV0 = 2 KV0 + sft - (F-2), t in [0, cw), cw = 2 nk + F-2.  Level-1 output
// c1 reads input pairs c1 .. c1+F/2-1.  Rows likewise (level-2 rows [R0 - 1, R1) of a row chunk).
// In periodization mode virtual level-1 columns / rows ARE the band (periodic); in the other modes level 2
// extends the level-1 band by its own boundary rule, so virtual index v is read at ext_index(v, band length).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "wavelets.cuh"

namespace spihtb {

constexpr int F2_NT = 256;   // threads per CTA
constexpr int F2_CW = 256;   // columns of TMA box A (= most input columns a strip can use)
constexpr int F2_CWB = 64;   // columns of TMA box B (columns the boundary map takes from the far side of the plane)
constexpr int F2_NST = 3;    // input stages in flight
constexpr int F2_SR = 8;     // input rows per stage
constexpr int F2_SB = 4;     // level-1 rows per step
constexpr int F2_TW1 = 128;  // ints per level-1 tile row
constexpr int F2_TW2 = 72;   // ints per level-2 tile row
constexpr int F2_NKP = 128;  // padded level-1 columns

template <int WID>
struct F2Cfg {
    static constexpr int F = Wav<WID>::F;
    static constexpr int HF = F / 2;
    static constexpr int FILL1 = (F - 2 + F2_SR - 1) / F2_SR;  // input stages before the first level-1 row
    static constexpr int FILL2 = (F - 2) / F2_SB;              // V2 steps before the first level-2 row
    static constexpr int LAGS = (F - 5 + 3) / 4;               // steps V2 runs behind the level-1 rows it reads
    // level-1 approximation rows kept in smem: at the top of a plane V2 reads rows 1 .. F (mirrored) while H1 is
    // already up to 4 (LAGS + ceil(F / 4)) + 3 rows further
    static constexpr int RING = F == 6 ? 16 : (F == 10 ? 32 : 64);
    static constexpr int NK = (F2_CW - (F - 2)) / 2;           // level-1 columns per strip at most
    // level-2 columns per strip: nk = 2 (NM + 1) + F - 2 (one extra column on the left completes the strip's
    // first pyramid cell), and a level-1 tile row holds 2 NM + 8 ints
    static constexpr int NM = ((NK - (F - 2)) / 2 - 1) < 60 ? ((NK - (F - 2)) / 2 - 1) : 60;
    static_assert((F - 2) % F2_SB == 0, "window fill is a whole number of steps");
    static_assert(RING >= 4 * (LAGS + (F + 3) / 4) + 4 && (RING & (RING - 1)) == 0, "ring");
    static_assert(4 + 3 + 2 * NM + 1 <= F2_TW1 && 4 + 3 + NM + 1 <= F2_TW2, "tile rows");
    static_assert(2 * (NM + 1) + F - 2 <= NK && NK <= F2_NKP && NM + 1 + HF <= F2_NKP / 2, "strip width");
};

struct F2K {
    int src_h, src_w;
    int bh1, bw1, bh2, bw2;
    int Hc, Wc;
    int sh1, sw1, sh2, sw2;
    int mode, C;
    int nstrips, nchunks, NMs, NRc;
    int use_b;                 // some strip needs box B
    int32_t *coeffs;
    double *ll2;               // [nz][bh2][bw2]
    uint8_t *dp;               // optional
    uint32_t *maxabs;
    int NH, NW;
    double scale[8];
    double q;
    const double *u8lut;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// one box of the 3-D tensor {columns, rows, planes} -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int x, int y, int z, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

// out-of-line boundary rule (keeps the division sequences out of the streaming loop)
__device__ __noinline__ int ext_index_f2(int g, int n, int mode) { return ext_index(g, n, mode); }

template <typename Tin>
__device__ __forceinline__ double f2_px(uint32_t saddr, const double *lut);
template <>
__device__ __forceinline__ double f2_px<float>(uint32_t saddr, const double *)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return (double)v;
}
template <>
__device__ __forceinline__ double f2_px<double>(uint32_t saddr, const double *)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
    return v;
}
template <>
__device__ __forceinline__ double f2_px<uint8_t>(uint32_t saddr, const double *lut)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return lut[v];  // k / 255.0 (utils.py:19), table in shared memory
}

// One output of the column filters from F/2 (E, O) pairs: ascending tap order, fma accumulation, zero taps
// skipped -- tap 2v multiplies O[k-v], tap 2v+1 multiplies E[k-v], as in dwt_fwd.cu.
template <int WID>
__device__ __forceinline__ void f2_hfilter(const double2 *pairs, double &o_lo, double &o_hi)
{
    constexpr int F = Wav<WID>::F, HF = F / 2;
    double2 pr[HF];
#pragma unroll
    for (int p = 0; p < HF; ++p) pr[p] = pairs[p];
    double lo = 0.0, hi = 0.0;
#pragma unroll
    for (int v = 0; v < HF; ++v) {
        const double2 e = pr[HF - 1 - v];  // pair k - v
        if (Wav<WID>::dec_lo(2 * v) != 0.0) lo = fma(Wav<WID>::dec_lo(2 * v), e.y, lo);
        if (wav_dec_hi<WID>(2 * v) != 0.0) hi = fma(wav_dec_hi<WID>(2 * v), e.y, hi);
        if (Wav<WID>::dec_lo(2 * v + 1) != 0.0) lo = fma(Wav<WID>::dec_lo(2 * v + 1), e.x, lo);
        if (wav_dec_hi<WID>(2 * v + 1) != 0.0) hi = fma(wav_dec_hi<WID>(2 * v + 1), e.x, hi);
    }
    o_lo = lo;
    o_hi = hi;
}

// ---- tiles: quantised detail coefficients of the rows in flight, staged in shared memory ----
// A tile row holds one band row of one detail band; index i is array column TB + i, TB even (so that the two
// columns of a pyramid cell sit at an even / odd index pair in every row).  OUT copies the stored part of a tile
// row to the coefficient array and forms the cells whose second row and second column are stored by this task.
struct BandTile {
    int ro, co;   // band origin in the coefficient array
    int tb;       // array column of tile index 0 (even)
    int i0, n;    // stored columns: tile indices [i0, i0 + n)
    int cmin;     // first array column of the band (cells need their first column inside the band)
};
// What a lane does for one band in OUT, fixed for the whole task (nothing here depends on the row):
//   copy : tile indices i0 + lane + 32 u, u < 4, those below i0 + n (bit u of cmask)
//   cells: tile indices 4 lane .. 4 lane + 3 = two cells; bit 0 / 1 of cellmask: the cell is this task's (its
//          second column is stored here and its first column lies in the band)
struct LaneOut {
    uint32_t cmask, cellmask;
    int goff;   // array column of the lane's first copied element
    int bc;     // cell column of the lane's first cell
};
__device__ __forceinline__ LaneOut make_lane_out(const BandTile &b, int lane, int NW)
{
    LaneOut o;
    o.cmask = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (lane + 32 * u < b.n) o.cmask |= 1u << u;
    o.goff = b.tb + b.i0 + lane;
    const int i = 4 * lane;
    o.bc = (b.tb + i) >> 1;
    o.cellmask = 0;
    if (i + 1 >= b.i0 && i + 1 < b.i0 + b.n && b.tb + i >= b.cmin && o.bc < NW) o.cellmask |= 1u;
    if (i + 3 >= b.i0 && i + 3 < b.i0 + b.n && b.tb + i + 2 >= b.cmin && o.bc + 1 < NW) o.cellmask |= 2u;
    return o;
}
// one tile row -> coefficient array row `grow` (warp-wide, 4-byte coalesced stores); returns the largest magnitude
__device__ __forceinline__ uint32_t f2_copy_row(const int32_t *trow, int32_t *grow, const LaneOut &o, int i0, int lane)
{
    const int32_t *t = trow + i0 + lane;
    int32_t *g = grow + o.goff;
    uint32_t mx = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (o.cmask & (1u << u)) {
            const int32_t v = t[32 * u];
            g[32 * u] = v;
            mx = max(mx, absu(v));
        }
    }
    return mx;
}
// the lane's two cells of the cell row whose second row is `trow` (first row `trow_prev`); drow: dp row of the cells
__device__ __forceinline__ void f2_cells_row(const int32_t *trow_prev, const int32_t *trow, uint8_t *drow, const LaneOut &o,
                                             int lane)
{
    if (o.cellmask == 0) return;
    const int4 q1 = *reinterpret_cast<const int4 *>(trow + 4 * lane);
    const int4 q0 = *reinterpret_cast<const int4 *>(trow_prev + 4 * lane);
    const uint32_t m0 = max(max(absu(q0.x), absu(q0.y)), max(absu(q1.x), absu(q1.y)));
    const uint32_t m1 = max(max(absu(q0.z), absu(q0.w)), max(absu(q1.z), absu(q1.w)));
    if (o.cellmask & 1u) drow[o.bc] = (uint8_t)plane1(m0);
    if (o.cellmask & 2u) drow[o.bc + 1] = (uint8_t)plane1(m1);
}

template <typename Tin, int WID>
__global__ void __launch_bounds__(F2_NT, (Wav<WID>::F == 18 ? 1 : 2))
dwt_fwd12_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const F2K p)
{
    using Cfg = F2Cfg<WID>;
    constexpr int F = Cfg::F, HF = Cfg::HF, FILL1 = Cfg::FILL1, FILL2 = Cfg::FILL2, LAGS = Cfg::LAGS, RING = Cfg::RING;
    constexpr int ES = (int)sizeof(Tin);
    constexpr int ROWB = ((F2_CW + F2_CWB) * ES + 127) / 128 * 128;   // bytes of one staged row (box A, then box B); TMA destinations are 128-byte aligned
    constexpr int STAGEB = F2_SR * ROWB;
    constexpr bool U8 = sizeof(Tin) == 1;

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sp = smem;
    unsigned char *s_in = sp;                 sp += F2_NST * STAGEB;
    double *s_v1 = reinterpret_cast<double *>(sp);   sp += 2 * F2_SB * F2_CW * sizeof(double);       // [lohi][row][col]
    double *s_ring = reinterpret_cast<double *>(sp); sp += RING * F2_NKP * sizeof(double);          // [slot][c1]
    double *s_v2 = reinterpret_cast<double *>(sp);   sp += 2 * 2 * F2_NKP * sizeof(double);         // [lohi][row][c1]
    int32_t *s_t1 = reinterpret_cast<int32_t *>(sp); sp += 3 * 8 * F2_TW1 * sizeof(int32_t);        // [band][row & 7][i]
    int32_t *s_t2 = reinterpret_cast<int32_t *>(sp); sp += 3 * 4 * F2_TW2 * sizeof(int32_t);        // [band][row & 3][i]
    double *s_lut = reinterpret_cast<double *>(sp);  sp += (U8 ? 256 : 0) * sizeof(double);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(sp); sp += F2_NST * sizeof(uint64_t);
    int *s_misc = reinterpret_cast<int *>(sp);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mode = p.mode;
    const bool per = mode == SPIHTB_MODE_PERIODIZATION;
    const int sft = per ? HF - 1 : 0;

    // ---- task
    int task = blockIdx.x;
    const int strip = task % p.nstrips;
    task /= p.nstrips;
    const int chunk = task % p.nchunks;
    const int z = task / p.nchunks;

    // level-2 columns [M0, M1) and rows [R0, R1) are stored; column M0 - 1 and row R0 - 1 are computed as well
    // (details only) so that every pyramid cell whose second row / column is stored here is complete
    const int M0 = strip * p.NMs, M1 = min(M0 + p.NMs, p.bw2), nm = M1 - M0;
    const int KV0 = 2 * (M0 - 1) + sft - (F - 2), nk = 2 * (nm + 1) + F - 2;
    const int XV0 = 2 * KV0 + sft - (F - 2), cw = 2 * nk + F - 2;
    const int R0 = chunk * p.NRc, R1 = min(R0 + p.NRc, p.bh2);
    const int VR0 = 2 * (R0 - 1) + sft - (F - 2), VRl = 2 * R1 + sft - 1;
    const int PR0 = per ? VR0 : max(VR0, 0);
    const int Pl = per ? VRl : min(VRl, p.bh1 - 1);
    const int nps = (Pl - PR0 + 1 + F2_SB - 1) / F2_SB;
    const int nvs = (VRl - VR0 + 1 + F2_SB - 1) / F2_SB;
    const int nsteps = max(nps, nvs + LAGS);
    const int IR0 = 2 * PR0 + sft - F2_SR * FILL1;
    const int nstages = nps + FILL1;

    // ---- column maps
    if (tid == 0) s_misc[0] = 0x7fffffff;
    if (U8)
        for (int k = tid; k < 256; k += F2_NT) s_lut[k] = p.u8lut[k];
    __syncthreads();
    // TMA boxes start at a 16-byte boundary of the row (the unit's addressing granularity)
    constexpr int AL = 16 / ES;
    const int AX0 = XV0 - (((XV0 % AL) + AL) % AL);
    const int rc = ext_index(XV0 + min(tid, cw - 1), p.src_w, mode);  // real input column of this thread
    const bool inA = rc >= AX0 && rc < AX0 + F2_CW;
    if (!inA) atomicMin(&s_misc[0], rc);
    __syncthreads();
    const bool need_b = s_misc[0] != 0x7fffffff;   // CTA-uniform
    const int BXB = need_b ? s_misc[0] - (((s_misc[0] % AL) + AL) % AL) : 0;
    const uint32_t in_off = (uint32_t)((inA ? rc - AX0 : F2_CW + (rc - BXB)) * ES);
    // level-1 column this thread's V2 window reads (threads < nk)
    const int c1 = min(tid, nk - 1);
    const int c1map = per ? c1 : min(max(ext_index(KV0 + c1, p.bw1, mode) - KV0, 0), nk - 1);

    // ---- barriers + first loads
    const uint32_t bar0 = smem_u32(s_bar);
    const uint32_t in0 = smem_u32(s_in);
    if (tid == 0) {
        for (int s = 0; s < F2_NST; ++s) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // one warp issues a stage: lane i loads row i into box A, lane 8 + i the same row into box B.  (A row
    // coordinate outside the plane goes through the boundary map; the division sequence of the map stays out of
    // the common path.)  The expect-tx arrive may land after the first complete-tx: the phase cannot complete
    // before the arrive, and the transaction count may be transiently negative.
    auto issue_stage = [&](int q) {
        const int s = q % F2_NST;
        const uint32_t bar = bar0 + 8 * s;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)(F2_SR * (F2_CW + (need_b ? F2_CWB : 0)) * ES));
        __syncwarp();
        if (lane < 2 * F2_SR) {
            const int i = lane & (F2_SR - 1);
            const int vrow = IR0 + F2_SR * q + i;
            const int row = (unsigned)vrow < (unsigned)p.src_h ? vrow : ext_index_f2(vrow, p.src_h, mode);
            const uint32_t dst = in0 + s * STAGEB + i * ROWB;
            if (lane < F2_SR)
                tma_load_3d(dst, &tmA, AX0, row, z, bar);
            else if (need_b)
                tma_load_3d(dst + F2_CW * ES, &tmB, BXB, row, z, bar);
        }
    };
    if (warp == 0)
        for (int q = 0; q < F2_NST && q < nstages; ++q) issue_stage(q);

    // ---- per-plane constants
    const int zc = z % p.C;
    const double mq = p.scale[zc], qs = p.q;
    const bool unit = mq == 1.0;   // (1.0 x) q == x q exactly
    int32_t *cplane = p.coeffs + (size_t)z * p.Hc * p.Wc;
    uint8_t *dpz = p.dp ? p.dp + (size_t)z * p.NH * p.NW : nullptr;
    // stored level-1 columns / rows of this task; level-1 column k sits at H1 thread c = k - KV0
    const int k1first = 2 * M0, n1core = max(0, min(2 * M1, p.bw1) - k1first);
    const int r1lo = 2 * R0, r1hi = min(2 * R1, p.bh1);
    // tiles: band b of level 1 (ad, da, dd), then of level 2
    auto make_tile = [&](int ro, int co, int kfirst, int n) {
        BandTile t;
        t.ro = ro;
        t.co = co;
        t.tb = ((co + kfirst) & ~1) - 2;
        t.i0 = co + kfirst - t.tb;
        t.n = n;
        t.cmin = co;
        return t;
    };
    const BandTile T1ad = make_tile(0, p.sw1, k1first, n1core), T1da = make_tile(p.sh1, 0, k1first, n1core),
                   T1dd = make_tile(p.sh1, p.sw1, k1first, n1core);
    const BandTile T2ad = make_tile(0, p.sw2, M0, nm), T2da = make_tile(p.sh2, 0, M0, nm), T2dd = make_tile(p.sh2, p.sw2, M0, nm);

    // OUT1 (warps 4-7, all three level-1 bands) and OUT2 (warp >> 1 = band) lane constants
    const LaneOut L1ad = make_lane_out(T1ad, lane, p.NW), L1da = make_lane_out(T1da, lane, p.NW),
                  L1dd = make_lane_out(T1dd, lane, p.NW);
    const int o2b = warp >> 1;
    const BandTile T2mine = o2b == 0 ? T2ad : (o2b == 1 ? T2da : T2dd);
    const LaneOut L2mine = make_lane_out(T2mine, lane, p.NW);

    // ---- H1 task of this thread: level-1 column c of the lo (warp even) or hi (warp odd) rows
    const int h1_lohi = warp & 1;
    const int h1_c = (warp >> 1) * 32 + lane;
    const int h1_k = KV0 + h1_c;
    // tile indices of this column (negative: not staged).  Staged: columns k1first - 1 .. k1first + n1core - 1
    const bool h1_core = h1_c < nk && h1_k >= max(k1first - 1, 0) && h1_k < k1first + n1core;
    const int h1_ta = h1_core ? (h1_lohi ? T1da.co + h1_k - T1da.tb : T1ad.co + h1_k - T1ad.tb) : -1;  // ad | da
    const int h1_tb = h1_core && h1_lohi ? T1dd.co + h1_k - T1dd.tb : -1;                                // dd
    // ---- H2 task: level-2 row e, lo / hi, column kk = M0 - 1 + c2
    const int h2_e = warp >> 2, h2_lohi = (warp >> 1) & 1;
    const int h2_c = (warp & 1) * 32 + lane;
    const int h2_k = M0 - 1 + h2_c;
    const bool h2_on = h2_c < nm + 1 && h2_k >= 0;
    const int h2_ta = h2_on ? (h2_lohi ? T2da.co + h2_k - T2da.tb : T2ad.co + h2_k - T2ad.tb) : -1;
    const int h2_tb = h2_on && h2_lohi ? T2dd.co + h2_k - T2dd.tb : -1;
    double *ll2z = p.ll2 + (size_t)z * p.bh2 * p.bw2 + h2_k;

    double w1[F - 2], w2[F - 2];
#pragma unroll
    for (int i = 0; i < F - 2; ++i) w1[i] = w2[i] = 0.0;
    uint32_t mx = 0;

    // Schedule of one iteration (three CTA barriers):
    //   A: V1(step)            + H2(step - 1)
    //   B: H1(step)            + OUT2(step - 1)
    //   C: V2(step) warps 0-3  | OUT1(step) warps 4-7
    for (int step = -FILL1; step <= nsteps; ++step) {
        const int q = step + FILL1;  // input stage
        // ================= V1: 8 input rows -> 4 (lo, hi) row pairs
        if (q < nstages) {
            mbar_wait(bar0 + 8 * (q % F2_NST), (uint32_t)((q / F2_NST) & 1));
            const uint32_t base = in0 + (q % F2_NST) * STAGEB + in_off;
            double x[F2_SR];
#pragma unroll
            for (int i = 0; i < F2_SR; ++i) x[i] = f2_px<Tin>(base + i * ROWB, s_lut);
            if (step >= 0) {
#pragma unroll
                for (int j = 0; j < F2_SB; ++j) {
                    double lo = 0.0, hi = 0.0;
#pragma unroll
                    for (int t = 0; t < F; ++t) {
                        const int n = 2 * j + F - 1 - t;  // index into [w1 (F-2) | x (8)]
                        const double v = n < F - 2 ? w1[n < F - 2 ? n : 0] : x[n >= F - 2 ? n - (F - 2) : 0];
                        if (Wav<WID>::dec_lo(t) != 0.0) lo = fma(Wav<WID>::dec_lo(t), v, lo);
                        if (wav_dec_hi<WID>(t) != 0.0) hi = fma(wav_dec_hi<WID>(t), v, hi);
                    }
                    s_v1[(0 * F2_SB + j) * F2_CW + tid] = lo;
                    s_v1[(1 * F2_SB + j) * F2_CW + tid] = hi;
                }
            }
            // carry the last F-2 rows
            double nw[F - 2];
#pragma unroll
            for (int i = 0; i < F - 2; ++i) {
                const int n = i + F2_SR;
                nw[i] = n < F - 2 ? w1[n < F - 2 ? n : 0] : x[n >= F - 2 ? n - (F - 2) : 0];
            }
#pragma unroll
            for (int i = 0; i < F - 2; ++i) w1[i] = nw[i];
        }
        // ================= H2 of the previous step: level-2 rows m = R0 - 1 + 2 (u - FILL2) + e
        {
            const int u = step - 1 - LAGS;
            const int m = R0 - 1 + 2 * (u - FILL2) + h2_e;
            if (u >= FILL2 && u < nvs && h2_on && m < R1 && m >= 0) {
                const double2 *pairs = reinterpret_cast<const double2 *>(s_v2 + (h2_lohi * 2 + h2_e) * F2_NKP) + h2_c;
                double o_lo, o_hi;
                f2_hfilter<WID>(pairs, o_lo, o_hi);
                int32_t *trow = s_t2 + (m & 3) * F2_TW2;
                if (h2_lohi == 0) {
                    if (h2_c >= 1 && m >= R0) ll2z[(size_t)m * p.bw2] = o_lo;  // aa -> level 3
                    trow[0 * 4 * F2_TW2 + h2_ta] = __double2int_rz((unit ? o_hi : mq * o_hi) * qs);  // ad
                } else {
                    trow[1 * 4 * F2_TW2 + h2_ta] = __double2int_rz((unit ? o_lo : mq * o_lo) * qs);  // da
                    trow[2 * 4 * F2_TW2 + h2_tb] = __double2int_rz((unit ? o_hi : mq * o_hi) * qs);  // dd
                }
            }
        }
        __syncthreads();  // A: (lo, hi) rows of level 1 and level-2 tile rows visible; the input stage is consumed
        if (warp == 0 && q + F2_NST < nstages && q < nstages) issue_stage(q + F2_NST);

        // ================= H1: level-1 outputs of production rows PR0 + 4 step + j
        if (step >= 0 && step < nps && h1_c < nk) {
            const int rbase = PR0 + F2_SB * step;
            double o_lo[F2_SB], o_hi[F2_SB];
#pragma unroll
            for (int j = 0; j < F2_SB; ++j) {
                const double2 *pairs = reinterpret_cast<const double2 *>(s_v1 + (h1_lohi * F2_SB + j) * F2_CW) + h1_c;
                f2_hfilter<WID>(pairs, o_lo[j], o_hi[j]);
            }
            // rows past Pl hold nothing useful: their ring / tile slots are never read
            if (h1_lohi == 0) {
#pragma unroll
                for (int j = 0; j < F2_SB; ++j) s_ring[((rbase + j - PR0) & (RING - 1)) * F2_NKP + h1_c] = o_lo[j];  // aa
            }
            if (h1_ta >= 0) {
#pragma unroll
                for (int j = 0; j < F2_SB; ++j) {
                    int32_t *trow = s_t1 + ((rbase + j) & 7) * F2_TW1;
                    if (h1_lohi == 0) {
                        trow[0 * 8 * F2_TW1 + h1_ta] = __double2int_rz((unit ? o_hi[j] : mq * o_hi[j]) * qs);  // ad
                    } else {
                        trow[1 * 8 * F2_TW1 + h1_ta] = __double2int_rz((unit ? o_lo[j] : mq * o_lo[j]) * qs);  // da
                        trow[2 * 8 * F2_TW1 + h1_tb] = __double2int_rz((unit ? o_hi[j] : mq * o_hi[j]) * qs);  // dd
                    }
                }
            }
        }
        // ================= OUT2 of the previous step (warps 0-5: row e = warp & 1, band = warp >> 1)
        {
            const int u = step - 1 - LAGS;
            if (u >= FILL2 && u < nvs && warp < 6) {
                const int m = R0 - 1 + 2 * (u - FILL2) + (warp & 1);
                if (m >= R0 && m < R1) {
                    const int32_t *trow = s_t2 + (o2b * 4 + (m & 3)) * F2_TW2;
                    const int ar = T2mine.ro + m;
                    mx = max(mx, f2_copy_row(trow, cplane + (size_t)ar * p.Wc, L2mine, T2mine.i0, lane));
                    if (dpz && (ar & 1) && m >= 1 && (ar >> 1) < p.NH)
                        f2_cells_row(s_t2 + (o2b * 4 + ((m - 1) & 3)) * F2_TW2, trow, dpz + (size_t)(ar >> 1) * p.NW, L2mine,
                                     lane);
                }
            }
        }
        __syncthreads();  // B: ring rows and level-1 tile rows of this step visible

        // ================= V2 (warps 0-3) | OUT1 (warps 4-7)
        const int u = step - LAGS;  // V2 step
        if (warp < 4) {
            if (u >= 0 && u < nvs) {
                double y[F2_SB];
#pragma unroll
                for (int i = 0; i < F2_SB; ++i) {
                    const int v = VR0 + F2_SB * u + i;                       // virtual level-1 row
                    const int rr = (per || (unsigned)v < (unsigned)p.bh1) ? v : ext_index_f2(v, p.bh1, mode);  // the row that holds it
                    y[i] = s_ring[((rr - PR0) & (RING - 1)) * F2_NKP + c1map];
                }
                if (u >= FILL2) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double lo = 0.0, hi = 0.0;
#pragma unroll
                        for (int t = 0; t < F; ++t) {
                            const int n = 2 * e + F - 1 - t;  // index into [w2 (F-2) | y (4)]
                            const double v = n < F - 2 ? w2[n < F - 2 ? n : 0] : y[n >= F - 2 ? n - (F - 2) : 0];
                            if (Wav<WID>::dec_lo(t) != 0.0) lo = fma(Wav<WID>::dec_lo(t), v, lo);
                            if (wav_dec_hi<WID>(t) != 0.0) hi = fma(wav_dec_hi<WID>(t), v, hi);
                        }
                        s_v2[(0 * 2 + e) * F2_NKP + tid] = lo;
                        s_v2[(1 * 2 + e) * F2_NKP + tid] = hi;
                    }
                }
                double nw[F - 2];
#pragma unroll
                for (int i = 0; i < F - 2; ++i) {
                    const int n = i + F2_SB;
                    nw[i] = n < F - 2 ? w2[n < F - 2 ? n : 0] : y[n >= F - 2 ? n - (F - 2) : 0];
                }
#pragma unroll
                for (int i = 0; i < F - 2; ++i) w2[i] = nw[i];
            }
        } else if (step >= 0 && step < nps) {
            // warp 4 + j: production row j of this step, all three bands
            const int r = PR0 + F2_SB * step + (warp - 4);
            if (r <= Pl && r >= r1lo && r < r1hi && n1core > 0) {
                const bool has_prev = dpz && r >= 1 && r - 1 >= PR0;
                const int32_t *trow = s_t1 + (r & 7) * F2_TW1, *tprev = s_t1 + ((r - 1) & 7) * F2_TW1;
                // ad: array row r; da, dd: array row sh1 + r
                int32_t *g_top = cplane + (size_t)r * p.Wc, *g_bot = cplane + (size_t)(p.sh1 + r) * p.Wc;
                mx = max(mx, f2_copy_row(trow, g_top, L1ad, T1ad.i0, lane));
                mx = max(mx, f2_copy_row(trow + 8 * F2_TW1, g_bot, L1da, T1da.i0, lane));
                mx = max(mx, f2_copy_row(trow + 16 * F2_TW1, g_bot, L1dd, T1dd.i0, lane));
                if (has_prev && (r & 1) && (r >> 1) < p.NH)
                    f2_cells_row(tprev, trow, dpz + (size_t)(r >> 1) * p.NW, L1ad, lane);
                const int ab = p.sh1 + r;
                if (has_prev && (ab & 1) && (ab >> 1) < p.NH) {
                    uint8_t *drow = dpz + (size_t)(ab >> 1) * p.NW;
                    f2_cells_row(tprev + 8 * F2_TW1, trow + 8 * F2_TW1, drow, L1da, lane);
                    f2_cells_row(tprev + 16 * F2_TW1, trow + 16 * F2_TW1, drow, L1dd, lane);
                }
            }
        }
        __syncthreads();  // C: (lo, hi) rows of level 2 visible
    }
    if (p.maxabs) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        if (lane == 0 && mx) atomicMax(p.maxabs + z / p.C, mx);
    }
}

// ---------------------------------------------------------------- host side
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        else
            cudaGetLastError();
    }
    return fn;
}

struct F2Plan {
    int nstrips, NMs, nchunks, NRc;
    bool use_b;
};

// Mirrors the index arithmetic of dwt_fwd12_kernel for every strip and row chunk and checks that each column /
// row a needed output reads is one the CTA holds (box A or B; a level-1 column it computes; a ring row that is
// produced and not yet overwritten).  Degenerate geometries (planes a few filter lengths wide) fail here and take
// the level-by-level kernel.
template <int WID>
bool f2_validate(const spihtb_geom &g, const F2Plan &pl, int es)
{
    using Cfg = F2Cfg<WID>;
    constexpr int F = Cfg::F, HF = Cfg::HF, FILL2 = Cfg::FILL2, LAGS = Cfg::LAGS, RING = Cfg::RING;
    const int mode = g.mode;
    const bool per = mode == SPIHTB_MODE_PERIODIZATION;
    const int sft = per ? HF - 1 : 0;
    const int src_h = g.in_h[0], src_w = g.in_w[0], bh1 = g.band_h[0], bw1 = g.band_w[0], bh2 = g.band_h[1],
              bw2 = g.band_w[1];
    (void)src_h;
    for (int strip = 0; strip < pl.nstrips; ++strip) {
        const int M0 = strip * pl.NMs, M1 = std::min(M0 + pl.NMs, bw2), nm = M1 - M0;
        if (nm <= 0 || nm > Cfg::NM) return false;
        const int KV0 = 2 * (M0 - 1) + sft - (F - 2), nk = 2 * (nm + 1) + F - 2;
        const int XV0 = 2 * KV0 + sft - (F - 2), cw = 2 * nk + F - 2;
        const int AL = 16 / es;  // box starts are 16-byte aligned
        const int AX0 = XV0 - (((XV0 % AL) + AL) % AL);
        if (nk > Cfg::NK || XV0 + cw > AX0 + F2_CW) return false;
        // input columns: box A [AX0, AX0 + CW), the rest within one box B
        int bmin = 0x7fffffff, bmax = -0x7fffffff;
        for (int t = 0; t < cw; ++t) {
            const int rc = ext_index(XV0 + t, src_w, mode);
            if (!(rc >= AX0 && rc < AX0 + F2_CW)) {
                bmin = std::min(bmin, rc);
                bmax = std::max(bmax, rc);
            }
        }
        if (bmin != 0x7fffffff && bmax - (bmin - (((bmin % AL) + AL) % AL)) >= F2_CWB) return false;
        // level-2 outputs kk in [max(M0-1, 0), M1): level-1 columns they read must be computed here
        for (int kk = std::max(M0 - 1, 0); kk < M1; ++kk)
            for (int i = 0; i < F; ++i) {
                const int kv = 2 * kk + sft - (F - 2) + i;
                const int c = (per ? kv : ext_index(kv, bw1, mode)) - KV0;
                if (c < 0 || c >= nk) return false;
            }
        // stored level-1 columns
        const int n1core = std::max(0, std::min(2 * M1, bw1) - 2 * M0);
        if (2 * M0 - KV0 - 1 < 0 || 2 * M0 - KV0 + n1core > nk) return false;
    }
    if (2 * pl.NMs * pl.nstrips < bw1) return false;  // every level-1 column is stored by some strip
    for (int chunk = 0; chunk < pl.nchunks; ++chunk) {
        const int R0 = chunk * pl.NRc, R1 = std::min(R0 + pl.NRc, bh2);
        if (R1 <= R0) return false;
        const int VR0 = 2 * (R0 - 1) + sft - (F - 2), VRl = 2 * R1 + sft - 1;
        const int PR0 = per ? VR0 : std::max(VR0, 0);
        const int Pl = per ? VRl : std::min(VRl, bh1 - 1);
        const int nvs = (VRl - VR0 + 1 + F2_SB - 1) / F2_SB;
        if (Pl < PR0) return false;
        for (int u = 0; u < nvs; ++u) {
            const int step = u + LAGS;
            const int nps = (Pl - PR0 + 1 + F2_SB - 1) / F2_SB;
            const int newest = PR0 + F2_SB * std::min(step, nps - 1) + F2_SB - 1;  // last ring row H1 has written (a step's rows past Pl included)
            for (int i = 0; i < F2_SB; ++i) {
                const int v = VR0 + F2_SB * u + i;
                // is this virtual row read by a level-2 row that matters?  rows m in [max(R0-1,0), R1)
                const int mlo = (v - sft - 1 + 1) / 2 - 1, mhi = (v - sft + (F - 2)) / 2 + 1;
                bool used = false;
                for (int m = std::max(std::max(R0 - 1, 0), mlo); m <= std::min(R1 - 1, mhi); ++m)
                    if (v >= 2 * m + sft - (F - 2) && v <= 2 * m + sft + 1) used = true;
                if (!used) continue;
                const int rr = per ? v : ext_index(v, bh1, mode);
                if (rr < PR0 || rr > std::min(Pl, newest) || rr + RING <= newest) return false;
            }
        }
        (void)FILL2;
        // stored level-1 rows are produced
        const int r1lo = 2 * R0, r1hi = std::min(2 * R1, bh1);
        if (r1lo < r1hi && (r1lo - 1 < PR0 && r1lo > 0)) return false;
        if (r1hi - 1 > Pl) return false;
    }
    if (2 * pl.NRc * pl.nchunks < bh1) return false;
    return true;
}

template <int WID>
size_t f2_smem_bytes(size_t es)
{
    using Cfg = F2Cfg<WID>;
    size_t b = 0;
    b += (size_t)F2_NST * F2_SR * (((F2_CW + F2_CWB) * es + 127) / 128 * 128);
    b += 2 * F2_SB * F2_CW * sizeof(double);
    b += (size_t)Cfg::RING * F2_NKP * sizeof(double);
    b += 2 * 2 * F2_NKP * sizeof(double);
    b += 3 * 8 * F2_TW1 * sizeof(int32_t);
    b += 3 * 4 * F2_TW2 * sizeof(int32_t);
    b += (es == 1 ? 256 : 0) * sizeof(double);
    b += F2_NST * sizeof(uint64_t);
    b += 64;
    return b;
}

template <typename Tin, int WID>
int f2_launch(spihtb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const F2K &k, int nz)
{
    const size_t smem = f2_smem_bytes<WID>(sizeof(Tin));
    static bool attr_set = false;
    if (!attr_set) {
        SPIHTB_CUDA_CHECK(cudaFuncSetAttribute(dwt_fwd12_kernel<Tin, WID>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem));
        attr_set = true;
    }
    const long long nb = (long long)nz * k.nstrips * k.nchunks;
    if (nb > 0x7fffffffLL) {
        set_error("forward DWT grid too large");
        return SPIHTB_ESHAPE;
    }
    dwt_fwd12_kernel<Tin, WID><<<(unsigned)nb, F2_NT, smem, ctx->stream>>>(tmA, tmB, k);
    ctx->launches++;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

template <int WID>
int f2_run(spihtb_ctx *ctx, const void *src, int pixel_dtype, const XformArgs &x, int32_t *coeffs, double *ll2,
           const PyrFuse *pf, const double *u8lut, bool *done)
{
    using Cfg = F2Cfg<WID>;
    const spihtb_geom &g = x.g;
    const int nz = x.B * x.C;
    *done = false;
    const size_t es = pixel_dtype == SPIHTB_F64 ? 8 : (pixel_dtype == SPIHTB_U8 ? 1 : 4);
    const int src_h = g.in_h[0], src_w = g.in_w[0];
    if (g.levels < 3) return SPIHTB_OK;
    // periodization pads an odd-length level-1 band with a repeated sample before level 2: the band is then not the
    // periodic continuation the strip / chunk halos assume
    if (g.mode == SPIHTB_MODE_PERIODIZATION && ((g.band_h[0] | g.band_w[0]) & 1)) return SPIHTB_OK;
    if (((size_t)src_w * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0) return SPIHTB_OK;
    // Opt-in (SPIHTB_FUSED12=1): measured on B200 the kernel is correct but slower than the level-by-level pair
    // on every BASELINE shape (3.9 ms against 1.6 + 0.6 ms for 256 x 3 x 1024^2): the transform is issue-bound, and
    // staging both filter passes through shared memory costs more instructions than the register / shuffle
    // formulation saves in HBM traffic (profiles/r02_dwt_fwd12_*).
    {
        const char *e = getenv("SPIHTB_FUSED12");
        if (!e || atoi(e) == 0) return SPIHTB_OK;
    }
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return SPIHTB_OK;

    F2Plan pl;
    const int bw2 = g.band_w[1], bh2 = g.band_h[1];
    // widest strip whose input columns fit box A behind a 16-byte aligned start: cw + (16 / es - 1) <= CW
    const int nk_max = (F2_CW - (int)(16 / es - 1) - (Cfg::F - 2)) / 2;
    const int nm_max = std::min(Cfg::NM, (nk_max - (Cfg::F - 2)) / 2 - 1);
    if (nm_max < 8) return SPIHTB_OK;
    pl.nstrips = (bw2 + nm_max - 1) / nm_max;
    pl.NMs = (bw2 + pl.nstrips - 1) / pl.nstrips;
    // row chunks: enough CTAs for a few waves, chunks no shorter than 32 level-2 rows
    const long long want = 8LL * ctx->sm_count;
    int nch = (int)std::min<long long>((want + (long long)nz * pl.nstrips - 1) / ((long long)nz * pl.nstrips),
                                       std::max(1, bh2 / 32));
    if (const char *e = getenv("SPIHTB_F12_CHUNKS")) nch = std::max(1, atoi(e));
    nch = std::max(1, std::min(nch, bh2));
    pl.NRc = (bh2 + nch - 1) / nch;
    pl.nchunks = (bh2 + pl.NRc - 1) / pl.NRc;
    if (!f2_validate<WID>(g, pl, (int)es)) return SPIHTB_OK;

    // tensor maps over the pixel planes {columns, rows, planes}
    CUtensorMap tmA, tmB;
    const CUtensorMapDataType dt = pixel_dtype == SPIHTB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                   : (pixel_dtype == SPIHTB_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    const cuuint64_t dims[3] = {(cuuint64_t)src_w, (cuuint64_t)src_h, (cuuint64_t)nz};
    const cuuint64_t strides[2] = {(cuuint64_t)src_w * es, (cuuint64_t)src_w * src_h * es};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t boxA[3] = {F2_CW, 1, 1}, boxB[3] = {F2_CWB, 1, 1};
    CUresult r1 = enc(&tmA, dt, 3, const_cast<void *>(src), dims, strides, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tmB, dt, 3, const_cast<void *>(src), dims, strides, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return SPIHTB_OK;  // shape the TMA unit cannot address: other path

    F2K k;
    k.src_h = src_h; k.src_w = src_w;
    k.bh1 = g.band_h[0]; k.bw1 = g.band_w[0]; k.bh2 = bh2; k.bw2 = bw2;
    k.Hc = g.enc_h; k.Wc = g.enc_w;
    k.sh1 = g.off_h[0]; k.sw1 = g.off_w[0]; k.sh2 = g.off_h[1]; k.sw2 = g.off_w[1];
    k.mode = g.mode; k.C = x.C;
    k.nstrips = pl.nstrips; k.nchunks = pl.nchunks; k.NMs = pl.NMs; k.NRc = pl.NRc;
    k.use_b = 0;
    k.coeffs = coeffs;
    k.ll2 = ll2;
    k.dp = pf ? pf->dp : nullptr;
    k.maxabs = pf ? pf->maxabs : nullptr;
    k.NH = g.enc_h / 2; k.NW = g.enc_w / 2;
    for (int c = 0; c < 8; ++c) k.scale[c] = x.scale[c];
    k.q = x.q;
    k.u8lut = u8lut;
    int rc;
    if (pixel_dtype == SPIHTB_F64)
        rc = f2_launch<double, WID>(ctx, tmA, tmB, k, nz);
    else if (pixel_dtype == SPIHTB_U8)
        rc = f2_launch<uint8_t, WID>(ctx, tmA, tmB, k, nz);
    else
        rc = f2_launch<float, WID>(ctx, tmA, tmB, k, nz);
    if (rc) return rc;
    *done = true;
    return SPIHTB_OK;
}

}  // namespace

// Levels 1 and 2 in one kernel when the geometry allows it (*done says whether it ran): coefficient bands of both
// levels, pyramid cells and maxabs (pf), level-2 approximation -> ll2 [nz][band_h[1]][band_w[1]] float64.
int launch_forward_fused12(spihtb_ctx *ctx, const void *src, int pixel_dtype, const XformArgs &x, int32_t *coeffs,
                           double *ll2, const PyrFuse *pf, const double *u8lut, bool *done)
{
    switch (x.g.wavelet) {
        case SPIHTB_WAVELET_BIOR22: return f2_run<SPIHTB_WAVELET_BIOR22>(ctx, src, pixel_dtype, x, coeffs, ll2, pf, u8lut, done);
        case SPIHTB_WAVELET_BIOR44: return f2_run<SPIHTB_WAVELET_BIOR44>(ctx, src, pixel_dtype, x, coeffs, ll2, pf, u8lut, done);
        case SPIHTB_WAVELET_BIOR68: return f2_run<SPIHTB_WAVELET_BIOR68>(ctx, src, pixel_dtype, x, coeffs, ll2, pf, u8lut, done);
    }
    *done = false;
    return SPIHTB_OK;
}

}  // namespace spihtb
