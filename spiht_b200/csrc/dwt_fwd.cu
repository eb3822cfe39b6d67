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
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "wavelets.cuh"

namespace spihtb {

constexpr int FW_WARPS = 4;  // warps (= independent tasks) per CTA
// output rows per task at most (launch_level and fix_rects_level must agree): long chunks amortise the window
// fill on the small levels, shorter ones balance the big first level better (measured)
static inline int fw_rhmax(int bh) { return bh >= 384 ? 80 : 96; }

struct FwdK {
    const void *src;        // [nz][src_h][src_w] planes of Tin
    int src_h, src_w;
    int bh, bw;             // band size of this level
    double *dst_ll;         // [nz][bh][bw] scratch (null at the last level)
    long long src_ps, dst_ps;  // elements between consecutive planes of src / dst_ll (src_h src_w and bh bw, except in
                               // the tail kernel, whose planes keep private scratch of a fixed stride)
    int32_t *coeffs;        // [nz][Hc][Wc]
    int Hc, Wc, sh, sw;     // detail block offsets of this level
    int mode, C, last;
    int tiles_x, tiles_y;   // strips across, row chunks down
    int RH;                 // output rows per chunk
    long long ntasks;       // nz * tiles_y * tiles_x
    double scale[8];
    double q;
    const double *u8lut;    // device table k / 255.0, k = 0 .. 255 (uint8 pixels only)
    // fused pyramid base pass (PYR kernels): dp planes [nz][NH][NW] and the per-image maximum magnitude
    uint8_t *dp;
    uint32_t *maxabs;
    int NH, NW;
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
    if constexpr (BYTES == 4) {
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(saddr));
        memcpy(v, &r, 4);
    } else if constexpr (BYTES == 8) {
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

// A lane holds two adjacent int32 outputs (v0 at p[0], v1 at p[1]).  When p is 8-byte aligned
// (warp-uniform: lanes are 8 bytes apart) the pair goes out as one store; otherwise every lane takes its
// left neighbour's v1 and stores (v1', v0) at p - 1, and the row's two end columns go out alone.  `fl` is
// the lane's store plan, fixed for the strip (no branches in the streaming loop):
//   bit 0: both columns inside the band          bit 1: only the first one
//   bit 2: first column and the neighbour's v1   bit 3: first column, no neighbour before it
//   bit 4: second column with no neighbour after it
// `mis`: p is not 8-byte aligned (warp-uniform).  Returns the left neighbour's v1 (the fused pyramid pass pairs it with v0 when the band starts at an
// odd column).
enum : uint32_t { SP_BOTH = 1, SP_FIRST = 2, SP_PAIR_PREV = 4, SP_V0_ALONE = 8, SP_V1_ALONE = 16 };
__device__ __forceinline__ int32_t store_pair(int32_t *p, bool mis, int32_t v0, int32_t v1, uint32_t fl)
{
    const int32_t up = __shfl_up_sync(0xffffffffu, v1, 1);
    // four predicated stores, no address selects and no branches
    if (!mis && (fl & SP_BOTH)) *reinterpret_cast<int2 *>(p) = make_int2(v0, v1);
    if (mis && (fl & SP_PAIR_PREV)) *reinterpret_cast<int2 *>(p - 1) = make_int2(up, v0);
    if (fl & (mis ? SP_V0_ALONE : SP_FIRST)) p[0] = v0;
    if (mis && (fl & SP_V1_ALONE)) p[1] = v1;
    return up;
}

// 64-bit shuffle as two 32-bit shuffles on explicitly split halves
__device__ __forceinline__ double shfl_up_f64(double v, int d)
{
    const int lo = __shfl_up_sync(0xffffffffu, __double2loint(v), d);
    const int hi = __shfl_up_sync(0xffffffffu, __double2hiint(v), d);
    return __hiloint2double(hi, lo);
}

// Fused pyramid base pass.  A tree node (a, b) owns the 2x2 cell rows 2a, 2a+1 x cols 2b, 2b+1 of the
// coefficient array; its dp byte is 1 + floor(log2 max|x|) over the cell (pyramid.cu).  While a warp
// streams a band, a lane sees both columns of a cell (its own pair, or its left neighbour's second column
// and its own first when the band starts at an odd column) and keeps the maximum over the cell's first
// row; on the cell's second row it writes the byte.  Cells that straddle two row chunks, two strips or
// two bands are left to pyr_fix_kernel, which recomputes them from the finished array (fix_rects_level).
struct CellTrack {
    uint32_t prev;  // cell maximum over the previous row
    int off;        // byte offset (within the plane's dp) of the cell this lane completes next
};
// cm: maximum over the lane's cell columns in the current row; odd: the row is the second row of its cell
// (warp-uniform); wr: this lane's cell has both columns in this strip and both rows in this chunk
__device__ __forceinline__ void cell_row(CellTrack &ct, uint32_t cm, bool odd, bool wr, uint8_t *dpz, int NW)
{
    const uint32_t m = max(ct.prev, cm);
    ct.prev = cm;
    // one predicated byte store (as inline PTX: left to itself the compiler branches around the store and its
    // address arithmetic three times per row, and the branches cost more than the few instructions they skip)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u8 [%0], %1;\n\t}"
        :
        : "l"(dpz + ct.off), "r"(plane1(m)), "r"((uint32_t)(odd && wr))
        : "memory");
    ct.off += odd ? NW : 0;
}

// pixel -> float64.  uint8 pixels are images as stored on disk; the reference's loader (spiht/utils.py:12-20,
// imload) scales them with `im / 255` in float64: the 256 quotients come from a table the host fills with the
// same IEEE division (exact, and no division in the streaming loop).
template <typename Tin>
__device__ __forceinline__ double px_f64(Tin v, const double *)
{
    return (double)v;
}
template <>
__device__ __forceinline__ double px_f64<uint8_t>(uint8_t v, const double *lut)
{
    return lut[v];
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
    // resident CTAs per SM the register allocation must allow (bior2.2: 4 x 128 threads x 128 registers)
    static constexpr int MINB = HF == 3 ? 4 : 1;
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
template <typename Tin, int WID, int NP, bool UNIT_M, bool LAST, bool PYR>
__device__ __forceinline__ void dwt_fwd_task(const FwdK &p, int tx, int ty, int z, uint32_t ring, bool aligned,
                                             const double *lut)
{
    constexpr int F = Wav<WID>::F;
    constexpr int HF = F / 2;
    constexpr int NC = 2 * NP;
    static_assert((HF - 1) % NP == 0, "halo must be a whole number of lanes");
    constexpr int HL = (HF - 1) / NP;           // lanes that only feed their neighbours
    constexpr int NOUT = 32 * NP - (HF - 1);
    constexpr int DEPTH = FwdCfg<WID>::DEPTH;
    // uint8 pixels: cp.async cannot copy single bytes, so a lane that goes element by element (boundary map)
    // copies the aligned 32-bit word around each pixel into a 4-byte slot and picks the byte when it reads the
    // row back; rows are 4-byte aligned (the launcher converts other widths to float64 first)
    constexpr bool U8 = sizeof(Tin) == 1;
    constexpr int ES = U8 ? 4 : (int)sizeof(Tin);   // ring bytes per element
    constexpr int VFRAG = NC * (int)sizeof(Tin);    // bytes a lane copies as one vector
    constexpr int FRAG = NC * ES;                   // ring bytes of one lane's row fragment
    constexpr int SLOT = 2 * 32 * FRAG;             // one row pair of the warp
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int src_h = p.src_h, src_w = p.src_w, mode = p.mode;
    const int sft = mode == SPIHTB_MODE_PERIODIZATION ? (HF - 1) : 0;  // even for every supported wavelet
    const int k0 = tx * NOUT, r0 = ty * p.RH;
    const int nrows = min(p.RH, p.bh - r0);
    const int kf = k0 - (HF - 1) + NP * lane;  // this lane's first pair index = its first output column
    const int gc = 2 * kf + sft;               // its first input column
    const int gr0 = 2 * r0 - (F - 2) + sft;    // first input row of the chunk's window

    const Tin *plane = static_cast<const Tin *>(p.src) + (size_t)z * (size_t)p.src_ps;
    const bool lane_vec = aligned && VFRAG >= 4 && gc >= 0 && gc + NC <= src_w;
    const bool warp_vec = __all_sync(0xffffffffu, lane_vec);
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
        if ((uint32_t)gr > (uint32_t)(src_h - 2)) {  // gr < 0 or gr + 1 >= src_h; rare (warp-uniform): boundary rows go through the extension map
            pa = plane + (size_t)ext_index_slow(gr, src_h, mode) * src_w;
            pb = plane + (size_t)ext_index_slow(gr + 1, src_h, mode) * src_w;
        }
        const uint32_t sa = my + slot_off, sb = sa + 32 * FRAG;
        constexpr int VB = VFRAG >= 16 ? 16 : (VFRAG >= 8 ? 8 : 4);
        if (warp_vec) {  // interior strips: every lane copies its fragment as vectors (no per-lane predicates)
#pragma unroll
            for (int o = 0; o < VFRAG; o += VB) {
                cp_async<VB>(sa + o, reinterpret_cast<const char *>(pa + gc) + o);
                cp_async<VB>(sb + o, reinterpret_cast<const char *>(pb + gc) + o);
            }
        } else if (lane_vec) {
#pragma unroll
            for (int o = 0; o < VFRAG; o += VB) {
                cp_async<VB>(sa + o, reinterpret_cast<const char *>(pa + gc) + o);
                cp_async<VB>(sb + o, reinterpret_cast<const char *>(pb + gc) + o);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int cc = U8 ? (col[c] & ~3) : col[c];
                cp_async<ES>(sa + c * ES, pa + cc);
                cp_async<ES>(sb + c * ES, pb + cc);
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
        if constexpr (U8) {
            if (lane_vec) {  // NP == 2: the lane's four pixels are one word
                uint32_t wa, wb;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa) : "r"(sa));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb) : "r"(sa + 32 * FRAG));
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    a[c] = (Tin)((wa >> (8 * (c & 3))) & 0xffu);
                    b[c] = (Tin)((wb >> (8 * (c & 3))) & 0xffu);
                }
            } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    uint32_t wa, wb;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa) : "r"(sa + c * ES));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb) : "r"(sa + 32 * FRAG + c * ES));
                    a[c] = (Tin)((wa >> (8 * (col[c] & 3))) & 0xffu);
                    b[c] = (Tin)((wb >> (8 * (col[c] & 3))) & 0xffu);
                }
            }
        } else {
            lds_frag<Tin, NC>(sa, a);
            lds_frag<Tin, NC>(sa + 32 * FRAG, b);
        }
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
            w[2 * j][c] = px_f64<Tin>(a[c], lut);
            w[2 * j + 1][c] = px_f64<Tin>(b[c], lut);
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
    // coefficient rows: one running pointer (array row r0 + i, column kf); the bands sit at offsets that
    // depend on the launch parameters only:  ad +sw (rows lo, cols hi: top right), da +sh Wc (rows hi, cols
    // lo: bottom left), dd +sh Wc + sw, and at the last level aa +0 (the LL corner)
    int32_t *prow = p.coeffs + (size_t)z * p.Hc * Wc + (ptrdiff_t)r0 * Wc + kf;
    const int o_ad = p.sw, o_da = p.sh * Wc, o_dd = p.sh * Wc + p.sw;
    constexpr bool ll_scratch = !LAST;
    double *p_ll = ll_scratch ? p.dst_ll + (size_t)z * (size_t)p.dst_ps + (size_t)r0 * p.bw + kf : nullptr;
    const int bw = p.bw;
    // the lane's store plan (see store_pair) and magnitude masks, fixed for the strip
    const bool ok0 = col_ok[0], ok1 = NP == 2 ? col_ok[NP - 1] : false;
    uint32_t fl = 0;
    if (ok0 && ok1) fl |= SP_BOTH;
    if (ok0 && !ok1) fl |= SP_FIRST;
    if (ok0 && ok_prev) fl |= SP_PAIR_PREV;
    if (ok0 && !ok_prev) fl |= SP_V0_ALONE;
    if (ok1 && !ok_next) fl |= SP_V1_ALONE;
    // The high-pass column filter of every supported wavelet ends one pair earlier than the low-pass one (the
    // taps at pair offset HF-1 are zero), so the last column of the lane *before* the first storing lane is
    // already correct in the two bands that start at column sw: a cell of those bands that straddles two strips
    // is completed by the right-hand strip, and the fix-up pass needs none of their interior columns.
    constexpr bool HI_SHORT = wav_dec_hi<WID>(2 * (HF - 1)) == 0.0 && wav_dec_hi<WID>(2 * (HF - 1) + 1) == 0.0;
    const bool ok_prev_hi = HI_SHORT ? (lane >= 1 && lane >= HL && kf - 1 >= 0 && kf - 1 < p.bw) : ok_prev;
    const uint32_t k0m = ok0 ? ~0u : 0u, k1m = ok1 ? ~0u : 0u, kpm = ok_prev ? ~0u : 0u,
                   kpm_hi = ok_prev_hi ? ~0u : 0u;
    // fused pyramid base pass: one cell tracker per detail band (band rows from ro, columns from co)
    CellTrack ct_ad, ct_da, ct_dd;
    uint32_t mx = 0;
    const int pyrNW = p.NW;
    uint8_t *dpz = PYR ? p.dp + (size_t)z * p.NH * p.NW : nullptr;
    bool wr_lo, wr_hi;  // lane's cell complete in this strip: bands starting at column 0 / at column sw
    auto cell_init = [&](CellTrack &ct, int ro, int co, bool &wr, bool okp) {
        const int c = co + kf, a = ro + r0;
        ct.prev = 0;
        ct.off = (a >> 1) * p.NW + (c >> 1);
        if (NP == 2)
            wr = (c & 1) ? (ok0 && okp) : (ok0 && ok1);
        else
            wr = (c & 1) && ok0 && okp;
    };
    cell_init(ct_ad, 0, p.sw, wr_hi, ok_prev_hi);
    cell_init(ct_da, p.sh, 0, wr_lo, ok_prev);
    cell_init(ct_dd, p.sh, p.sw, wr_hi, ok_prev_hi);
    const int par_ad = r0 & 1, par_lo = (p.sh + r0) & 1;  // parity of the chunk's first row in the array
    const bool codd_hi = ((p.sw + kf) & 1) != 0, codd_lo = (kf & 1) != 0;  // first column odd (warp-uniform for NP == 2)
    // cell maximum of one band row and the running maximum of everything stored
    auto cell_max = [&](int32_t v0, int32_t v1, int32_t up, bool codd, uint32_t kp) -> uint32_t {
        const uint32_t m0 = absu(v0) & k0m;
        if (NP == 2) {
            const uint32_t t = max(m0, absu(v1) & k1m);
            mx = max(mx, t);
            return codd ? max(absu(up) & kp, m0) : t;
        }
        mx = max(mx, m0);
        return max(absu(up) & kp, m0);
    };

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
                    w[(2 * u + F - 2) % F][c] = px_f64<Tin>(a[c], lut);
                    w[(2 * u + F - 1) % F][c] = px_f64<Tin>(b[c], lut);
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
                        xlo[j][eo] = used ? shfl_up_f64(lo[2 * idx + eo], d) : 0.0;
                        xhi[j][eo] = used ? shfl_up_f64(hi[2 * idx + eo], d) : 0.0;
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
            if (row_ok) {  // warp-uniform; false only for the padding rows after the chunk's last row
                int32_t u_ad, u_da, u_dd;
                if constexpr (NP == 2) {
                    // both columns of a lane go out as one 8-byte store (whole 32-byte sectors per warp)
                    const int mrow = (int)(reinterpret_cast<uintptr_t>(prow) >> 2);  // bit 0: row pointer odd
                    u_ad = store_pair(prow + o_ad, ((mrow + o_ad) & 1) != 0, q_ad[0], q_ad[1], fl);
                    u_da = store_pair(prow + o_da, ((mrow + o_da) & 1) != 0, q_da[0], q_da[1], fl);
                    u_dd = store_pair(prow + o_dd, ((mrow + o_dd) & 1) != 0, q_dd[0], q_dd[1], fl);
                    if (ll_scratch) {
                        const bool pair = (fl & SP_BOTH) && (reinterpret_cast<uintptr_t>(p_ll) & 15) == 0;
                        if (pair) *reinterpret_cast<double2 *>(p_ll) = make_double2(f_aa[0], f_aa[1]);
                        if (ok0 && !pair) p_ll[0] = f_aa[0];
                        if (ok1 && !pair) p_ll[1] = f_aa[1];
                    } else {
                        store_pair(prow, (mrow & 1) != 0, q_aa[0], q_aa[1], fl);
                    }
                } else {
                    if (ok0) {
                        prow[o_ad] = q_ad[0];
                        prow[o_da] = q_da[0];
                        prow[o_dd] = q_dd[0];
                        if (ll_scratch)
                            p_ll[0] = f_aa[0];
                        else
                            prow[0] = q_aa[0];
                    }
                    if constexpr (PYR) {
                        u_ad = __shfl_up_sync(FULL, q_ad[0], 1);
                        u_da = __shfl_up_sync(FULL, q_da[0], 1);
                        u_dd = __shfl_up_sync(FULL, q_dd[0], 1);
                    }
                }
                if constexpr (PYR) {
                    const bool not_first = i > 0;
                    const bool odd_ad = ((i ^ par_ad) & 1) != 0, odd_lo = ((i ^ par_lo) & 1) != 0;
                    cell_row(ct_ad, cell_max(q_ad[0], q_ad[NP - 1], u_ad, codd_hi, kpm_hi), odd_ad, wr_hi && not_first, dpz, pyrNW);
                    cell_row(ct_da, cell_max(q_da[0], q_da[NP - 1], u_da, codd_lo, kpm), odd_lo, wr_lo && not_first, dpz, pyrNW);
                    cell_row(ct_dd, cell_max(q_dd[0], q_dd[NP - 1], u_dd, codd_hi, kpm_hi), odd_lo, wr_hi && not_first, dpz, pyrNW);
                    if (!ll_scratch) {
                        mx = max(mx, absu(q_aa[0]) & k0m);
                        if (NP == 2) mx = max(mx, absu(q_aa[NP - 1]) & k1m);
                    }
                }
            }
            prow += Wc;
            if (ll_scratch) p_ll += bw;
        }
    }
    if constexpr (PYR) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
        if (lane == 0 && mx) atomicMax(p.maxabs + z / p.C, mx);
    }
    cp_async_wait<0>();  // the ring is free again (the prefetches past the chunk's last row are dropped)
}

// ---- the coarse levels of one plane in one CTA ------------------------------------------------------------
// From the level whose plane is a few dozen tasks on, a launch per level is all launch latency and tail: the
// bench workload spends 0.5 ms in levels 3..7 for 3 % of the coefficients.  Here a CTA of FT_WARPS warps owns
// one (image, channel) plane and runs those levels back to back: the tasks of a level are dealt to the warps
// (same tiles, same dwt_fwd_task as the per-level kernel, so the results are bit-identical), a CTA barrier
// separates the levels (a level reads the approximation plane the previous one wrote; stores invalidate the
// SM's own L1 lines, so the cp.async reads see them).  Planes are independent: while one CTA is down in its
// latency-bound 12 x 12 level, its neighbours stream level 2.
constexpr int FT_WARPS = 8;
constexpr int FT_MAXLV = 10;
struct FwdTail {
    FwdK lv[FT_MAXLV];
    int nlv;
};
template <int WID, int NP, bool UNIT_M, bool PYR>
__global__ void __launch_bounds__(FT_WARPS * 32, (Wav<WID>::F == 6 ? 2 : 1)) dwt_fwd_tail_kernel(const __grid_constant__ FwdTail p)
{
    constexpr int F = Wav<WID>::F;
    constexpr int NOUT = 32 * NP - (F / 2 - 1);
    constexpr int DEPTH = FwdCfg<WID>::DEPTH;
    constexpr int WARP_RING = DEPTH * 2 * 32 * 2 * NP * (int)sizeof(double);
    extern __shared__ __align__(16) unsigned char s_tail_ring[];
    const int warp = threadIdx.x >> 5;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_tail_ring) + warp * WARP_RING;
    const int z = blockIdx.x;
    for (int li = 0; li < p.nlv; ++li) {
        const FwdK &k = p.lv[li];
        const int ntp = k.tiles_x * k.tiles_y;
        const int sft = k.mode == SPIHTB_MODE_PERIODIZATION ? (F / 2 - 1) : 0;
        for (int t = warp; t < ntp; t += FT_WARPS) {
            const int tx = t % k.tiles_x, ty = t / k.tiles_x;
            const int gc_first = 2 * (tx * NOUT - (F / 2 - 1)) + sft;
            constexpr int VB = 2 * NP * (int)sizeof(double) >= 16 ? 16 : 8;
            const long long base = (long long)reinterpret_cast<uintptr_t>(k.src) + (long long)z * k.src_ps * 8LL;
            const bool aligned = ((size_t)k.src_w * 8) % VB == 0 && (base + (long long)gc_first * 8LL) % VB == 0;
            if (k.last)
                dwt_fwd_task<double, WID, NP, UNIT_M, true, PYR>(k, tx, ty, z, ring, aligned, nullptr);
            else
                dwt_fwd_task<double, WID, NP, UNIT_M, false, PYR>(k, tx, ty, z, ring, aligned, nullptr);
        }
        __syncthreads();
    }
}

template <typename Tin, int WID, int NP, bool UNIT_M, bool LAST, bool PYR>
__global__ void __launch_bounds__(FW_WARPS * 32, FwdCfg<WID>::MINB) dwt_fwd_level_kernel(const FwdK p)
{
    constexpr int F = Wav<WID>::F;
    constexpr int NOUT = 32 * NP - (F / 2 - 1);
    constexpr int DEPTH = FwdCfg<WID>::DEPTH;
    constexpr int WARP_RING = DEPTH * 2 * 32 * 2 * NP * (sizeof(Tin) == 1 ? 4 : (int)sizeof(Tin));
    __shared__ __align__(16) unsigned char s_ring[FW_WARPS * WARP_RING];
    __shared__ double s_lut[sizeof(Tin) == 1 ? 256 : 1];  // k / 255 for uint8 pixels
    if constexpr (sizeof(Tin) == 1) {
        for (int k = threadIdx.x; k < 256; k += FW_WARPS * 32) s_lut[k] = p.u8lut[k];
        __syncthreads();
    }
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
    constexpr int VB = 2 * NP * (int)sizeof(Tin) >= 16 ? 16 : (2 * NP * (int)sizeof(Tin) >= 8 ? 8 : 4);  // vector bytes
    const long long base = (long long)reinterpret_cast<uintptr_t>(p.src) +
                           (long long)z * p.src_ps * (long long)sizeof(Tin);
    const bool aligned = ((size_t)p.src_w * sizeof(Tin)) % VB == 0 &&
                         (base + (long long)gc_first * (long long)sizeof(Tin)) % VB == 0;
    dwt_fwd_task<Tin, WID, NP, UNIT_M, LAST, PYR>(p, tx, ty, z, ring, aligned, s_lut);
}

// zero the gaps coeffs_to_array leaves between a level's off-diagonal blocks
// and the square of coarser levels: rows [bh,sh) x cols [sw,sw+bw) and
// rows [sh,sh+bh) x cols [bw,sw).  The host lists them as rectangles; a warp takes
// one row of one rectangle (blockDim (32, 8), blockIdx.x over the rows of all
// rectangles, blockIdx.y over the planes), so there is no division per element.
// With the fused pyramid pass the node cells that lie entirely inside a gap get
// their zero byte here as well (every other cell is written by the transform's
// epilogue or by pyr_fix_kernel).
struct GapK {
    int32_t *coeffs;
    uint8_t *dp;  // null without the fused pyramid pass
    int Hc, Wc, NH, NW, nz, nrect, nrows;
    int r0[2 * SPIHTB_MAX_LEVELS], r1[2 * SPIHTB_MAX_LEVELS], c0[2 * SPIHTB_MAX_LEVELS], c1[2 * SPIHTB_MAX_LEVELS];
    int first[2 * SPIHTB_MAX_LEVELS + 1];  // running row count
};
__global__ void __launch_bounds__(256) gap_fill_kernel(const __grid_constant__ GapK p)
{
    const int gr = blockIdx.x * 8 + threadIdx.y;
    if (gr >= p.nrows) return;
    int q = 0;
    while (q + 1 < p.nrect && gr >= p.first[q + 1]) ++q;
    const int r = p.r0[q] + (gr - p.first[q]), c0 = p.c0[q], c1 = p.c1[q];
    // the cell row this coefficient row opens, when both of its rows lie in the rectangle
    const bool cell_row = p.dp && !(r & 1) && r + 1 < p.r1[q] && (r >> 1) < p.NH;
    for (int z = blockIdx.y; z < p.nz; z += gridDim.y) {
        int32_t *row = p.coeffs + ((size_t)z * p.Hc + r) * p.Wc;
        for (int c = c0 + threadIdx.x; c < c1; c += 32) row[c] = 0;
        if (cell_row) {
            uint8_t *drow = p.dp + ((size_t)z * p.NH + (r >> 1)) * p.NW;
            for (int b = ((c0 + 1) >> 1) + threadIdx.x; 2 * b + 1 < c1 && b < p.NW; b += 32) drow[b] = 0;
        }
    }
}

// uint8 pixels whose rows are not 4-byte aligned: imload's im / 255 as a separate pass
__global__ void __launch_bounds__(256) u8_to_f64_kernel(const uint8_t *__restrict__ src, double *__restrict__ dst, size_t n)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
        dst[t] = (double)src[t] / 255.0;
}

// tiles of a level (shared by launch_level, the tail kernel and fix_rects_level)
template <int WID>
static void level_tiles(FwdK &k)
{
    constexpr int NOUT = FwdCfg<WID>::NOUT;
    const int RHMAX = fw_rhmax(k.bh);
    k.tiles_x = (k.bw + NOUT - 1) / NOUT;
    k.tiles_y = (k.bh + RHMAX - 1) / RHMAX;
    k.RH = (k.bh + k.tiles_y - 1) / k.tiles_y;
}

template <typename Tin, int WID>
static int launch_level(spihtb_ctx *ctx, FwdK k, int nz)
{
    constexpr int NP = FwdCfg<WID>::NP;
    level_tiles<WID>(k);   // balanced row chunks (no nearly empty tail chunk)
    k.ntasks = (long long)k.tiles_x * k.tiles_y * nz;
    const long long nb = (k.ntasks + FW_WARPS - 1) / FW_WARPS;
    if (nb > 0x7fffffffLL) {
        set_error("forward DWT grid too large");
        return SPIHTB_ESHAPE;
    }
    bool unit = true;
    for (int c = 0; c < k.C && c < 8; ++c) unit = unit && k.scale[c] == 1.0;
    const dim3 grid((unsigned)nb), block(FW_WARPS * 32);
    auto go = [&](auto unit_c, auto last_c) {
        constexpr bool U = decltype(unit_c)::value, LA = decltype(last_c)::value;
        if (k.dp)
            dwt_fwd_level_kernel<Tin, WID, NP, U, LA, true><<<grid, block, 0, ctx->stream>>>(k);
        else
            dwt_fwd_level_kernel<Tin, WID, NP, U, LA, false><<<grid, block, 0, ctx->stream>>>(k);
    };
    if (unit && k.last)
        go(std::true_type{}, std::true_type{});
    else if (unit)
        go(std::true_type{}, std::false_type{});
    else if (k.last)
        go(std::false_type{}, std::true_type{});
    else
        go(std::false_type{}, std::false_type{});
    ctx->launches++;
    return SPIHTB_OK;
}

template <int WID>
static int launch_tail(spihtb_ctx *ctx, FwdTail &t, int nz)
{
    constexpr int NP = FwdCfg<WID>::NP;
    constexpr int DEPTH = FwdCfg<WID>::DEPTH;
    constexpr size_t smem = (size_t)FT_WARPS * DEPTH * 2 * 32 * 2 * NP * sizeof(double);
    bool unit = true;
    for (int c = 0; c < t.lv[0].C && c < 8; ++c) unit = unit && t.lv[0].scale[c] == 1.0;
    const bool pyr = t.lv[0].dp != nullptr;
    for (int i = 0; i < t.nlv; ++i) {
        level_tiles<WID>(t.lv[i]);
        t.lv[i].ntasks = (long long)t.lv[i].tiles_x * t.lv[i].tiles_y * nz;
    }
    auto go = [&](auto unit_c, auto pyr_c) -> int {
        constexpr bool U = decltype(unit_c)::value, PY = decltype(pyr_c)::value;
        static bool attr_set = false;
        if (!attr_set) {
            SPIHTB_CUDA_CHECK(cudaFuncSetAttribute(dwt_fwd_tail_kernel<WID, NP, U, PY>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        dwt_fwd_tail_kernel<WID, NP, U, PY><<<(unsigned)nz, FT_WARPS * 32, smem, ctx->stream>>>(t);
        return SPIHTB_OK;
    };
    int rc;
    if (unit && pyr)
        rc = go(std::true_type{}, std::true_type{});
    else if (unit)
        rc = go(std::true_type{}, std::false_type{});
    else if (pyr)
        rc = go(std::false_type{}, std::true_type{});
    else
        rc = go(std::false_type{}, std::false_type{});
    if (rc) return rc;
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

// Cells of the node grid the fused base pass leaves open (a superset is harmless: the fix-up pass
// recomputes a cell from the finished array).  Per level and detail band (rows from `ro`, columns from
// `co`): the cell row of a chunk's first row when that row is odd and of its last row when that is even;
// the cell column of a strip's first column when odd and of its last column when even; and the LL block.
// whole: the level was written by the fused two-level kernel (dwt_fwd2.cu), which completes every cell inside a
// band itself: only the band's own edges remain.
template <int WID>
static void fix_rects_level(std::vector<FixRect> &out, int bh, int bw, int sh, int sw, int NH, int NW, bool whole)
{
    const int NOUT = whole ? bw : FwdCfg<WID>::NOUT;
    const int RHMAX = whole ? bh : fw_rhmax(bh);
    const int tiles_x = (bw + NOUT - 1) / NOUT, tiles_y = (bh + RHMAX - 1) / RHMAX;
    const int RH = (bh + tiles_y - 1) / tiles_y;
    auto add = [&](int a0, int a1, int b0, int b1) {
        a1 = std::min(a1, NH);
        b1 = std::min(b1, NW);
        if (a0 < a1 && b0 < b1) out.push_back(FixRect{a0, a1, b0, b1});
    };
    const int bands[3][2] = {{0, sw}, {sh, 0}, {sh, sw}};  // ad, da, dd: (ro, co)
    for (const auto &bd : bands) {
        const int ro = bd[0], co = bd[1];
        const int b0 = co >> 1, b1 = ((co + bw - 1) >> 1) + 1;
        const int a0 = ro >> 1, a1 = ((ro + bh - 1) >> 1) + 1;
        for (int ty = 0; ty < tiles_y; ++ty) {
            const int r0 = ty * RH, nrows = std::min(RH, bh - r0);
            if (nrows <= 0) break;
            const int first = ro + r0, last = ro + r0 + nrows - 1;
            if (first & 1) add(first >> 1, (first >> 1) + 1, b0, b1);
            if (!(last & 1)) add(last >> 1, (last >> 1) + 1, b0, b1);
        }
        // (the bands that start at column sw complete their strip-straddling cells themselves when the
        // high-pass filter is one pair shorter, see HI_SHORT in dwt_fwd_task: only the band's own edges remain)
        constexpr int HF = Wav<WID>::F / 2;
        constexpr bool HI_SHORT = wav_dec_hi<WID>(2 * (HF - 1)) == 0.0 && wav_dec_hi<WID>(2 * (HF - 1) + 1) == 0.0;
        const bool self = HI_SHORT && co != 0;
        for (int tx = 0; tx < tiles_x; ++tx) {
            const int first = co + tx * NOUT, last = co + std::min((tx + 1) * NOUT, bw) - 1;
            if ((first & 1) && !(self && tx > 0)) add(a0, a1, first >> 1, (first >> 1) + 1);
            if (!(last & 1) && !(self && tx + 1 < tiles_x)) add(a0, a1, last >> 1, (last >> 1) + 1);
        }
    }
}

static int upload_fix_rects(spihtb_ctx *ctx, const spihtb_geom &g, bool fused12, const FixRect **rects,
                            const uint32_t **prefix, int *nrect, uint32_t *total)
{
    // key: everything the rectangle list depends on
    int32_t key[32] = {0};
    key[0] = g.wavelet; key[1] = g.levels; key[2] = g.enc_h; key[3] = g.enc_w; key[4] = g.ll_h; key[5] = g.ll_w;
    for (int l = 0; l < g.levels && l < 6; ++l) {
        key[6 + 4 * l] = g.band_h[l]; key[7 + 4 * l] = g.band_w[l];
        key[8 + 4 * l] = g.off_h[l]; key[9 + 4 * l] = g.off_w[l];
    }
    key[30] = fused12 ? 1 : 0;
    key[31] = 0x5eee;
    std::vector<int32_t> &h = ctx->fix_host;
    const bool hit = h.size() >= 34 && memcmp(h.data(), key, sizeof(key)) == 0 && ctx->fix.p;
    if (!hit) {
        const int NH = g.enc_h / 2, NW = g.enc_w / 2;
        std::vector<FixRect> r;
        for (int l = 0; l < g.levels; ++l) {
            switch (g.wavelet) {
                case SPIHTB_WAVELET_BIOR22:
                    fix_rects_level<SPIHTB_WAVELET_BIOR22>(r, g.band_h[l], g.band_w[l], g.off_h[l], g.off_w[l], NH, NW,
                                                             fused12 && l < 2);
                    break;
                case SPIHTB_WAVELET_BIOR44:
                    fix_rects_level<SPIHTB_WAVELET_BIOR44>(r, g.band_h[l], g.band_w[l], g.off_h[l], g.off_w[l], NH, NW,
                                                             fused12 && l < 2);
                    break;
                default:
                    fix_rects_level<SPIHTB_WAVELET_BIOR68>(r, g.band_h[l], g.band_w[l], g.off_h[l], g.off_w[l], NH, NW,
                                                             fused12 && l < 2);
                    break;
            }
        }
        r.push_back(FixRect{0, std::min((g.ll_h + 1) / 2, NH), 0, std::min((g.ll_w + 1) / 2, NW)});  // LL block
        // The edges of the gaps as well.  gap_fill_kernel writes the cells that lie inside ONE gap rectangle and the
        // band rectangles above cover the cells a gap shares with a band, but a cell can also straddle two gaps of
        // different levels: the gap below a level's 'ad' band starts at column off_w, and when that is odd the cell
        // (off_w - 1, off_w) has its other column in a gap of the next coarser level (found by
        // tools/fuzz_encode_paths.py on 3 x 45 x 113, bior6.8, 4 levels: off_w[1] = 75; no BASELINE shape has it).
        auto gap_edges = [&](int r0, int r1, int c0, int c1) {
            if (r0 >= r1 || c0 >= c1) return;
            const int a0 = r0 >> 1, a1 = std::min(((r1 - 1) >> 1) + 1, NH), b0 = c0 >> 1, b1 = std::min(((c1 - 1) >> 1) + 1, NW);
            auto add = [&](int x0, int x1, int y0, int y1) {
                x1 = std::min(x1, NH);
                y1 = std::min(y1, NW);
                if (x0 < x1 && y0 < y1) r.push_back(FixRect{x0, x1, y0, y1});
            };
            if (r0 & 1) add(a0, a0 + 1, b0, b1);
            if (!((r1 - 1) & 1)) add(a1 - 1, a1, b0, b1);
            if (c0 & 1) add(a0, a1, b0, b0 + 1);
            if (!((c1 - 1) & 1)) add(a0, a1, b1 - 1, b1);
        };
        for (int l = 0; l < g.levels; ++l) {
            gap_edges(g.band_h[l], g.off_h[l], g.off_w[l], g.off_w[l] + g.band_w[l]);
            gap_edges(g.off_h[l], g.off_h[l] + g.band_h[l], g.band_w[l], g.off_w[l]);
        }
        h.assign(key, key + 32);
        h.push_back((int32_t)r.size());
        h.push_back(0);
        uint64_t tot = 0;
        std::vector<uint32_t> pre;
        for (const FixRect &q : r) {
            pre.push_back((uint32_t)tot);
            tot += (uint64_t)(q.a1 - q.a0) * (q.b1 - q.b0);
            h.push_back(q.a0); h.push_back(q.a1); h.push_back(q.b0); h.push_back(q.b1);
        }
        pre.push_back((uint32_t)tot);
        if (tot > 0x7fffffffull) {
            set_error("pyramid fix-up list too large");
            return SPIHTB_ESHAPE;
        }
        h[33] = (int32_t)tot;
        for (uint32_t v : pre) h.push_back((int32_t)v);
        int rc = ctx->ensure(ctx->fix, h.size() * sizeof(int32_t) + 256);
        if (rc) return rc;
        SPIHTB_CUDA_CHECK(cudaMemcpyAsync(ctx->fix.p, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                                          ctx->stream));
    }
    *nrect = h[32];
    *total = (uint32_t)h[33];
    const int32_t *d = static_cast<const int32_t *>(ctx->fix.p);
    *rects = reinterpret_cast<const FixRect *>(d + 34);
    *prefix = reinterpret_cast<const uint32_t *>(d + 34 + 4 * (size_t)h[32]);
    return SPIHTB_OK;
}

int launch_forward(spihtb_ctx *ctx, const void *pixels, const XformArgs &x, int32_t *coeffs, const PyrFuse *pf)
{
    const spihtb_geom &g = x.g;
    const int nz = x.B * x.C;
    const int L = g.levels;
    const bool generic = wavelet_is_generic(g.wavelet);
    if (generic && pf) {
        set_error("the fused pyramid base pass is not available for wavelet id %d", g.wavelet);
        return SPIHTB_EINVAL;
    }
    // scratch: two float64 approximation planes (level 1 output is the largest)
    const size_t ll1 = (size_t)nz * g.band_h[0] * g.band_w[0] * sizeof(double);
    const size_t ll2 = L > 1 ? (size_t)nz * g.band_h[1] * g.band_w[1] * sizeof(double) : 0;
    int rc = ctx->ensure(ctx->tmpb, ll2 + 256);
    if (rc) return rc;

    const double *u8lut = nullptr;
    const void *conv = nullptr;  // uint8 pixels converted to float64 up front (unaligned rows)
    if (x.pixel_dtype == SPIHTB_U8 && x.color != SPIHTB_COLOR_IPT &&
        (g.w % 4 != 0 || (reinterpret_cast<uintptr_t>(pixels) & 3) != 0)) {
        const size_t n = (size_t)nz * g.h * g.w;
        rc = ctx->ensure(ctx->io2, n * sizeof(double) + 256);
        if (rc) return rc;
        const unsigned nb = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 32);
        u8_to_f64_kernel<<<nb, 256, 0, ctx->stream>>>(static_cast<const uint8_t *>(pixels),
                                                      static_cast<double *>(ctx->io2.p), n);
        ctx->launches++;
        conv = ctx->io2.p;
    } else if (x.pixel_dtype == SPIHTB_U8 && x.color != SPIHTB_COLOR_IPT) {
        if (!ctx->u8lut.p) {
            rc = ctx->ensure(ctx->u8lut, 256 * sizeof(double));
            if (rc) return rc;
            double h[256];
            for (int k = 0; k < 256; ++k) h[k] = (double)k / 255.0;  // utils.py:19  im / 255
            SPIHTB_CUDA_CHECK(cudaMemcpyAsync(ctx->u8lut.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
            SPIHTB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // h is on the stack
        }
        u8lut = static_cast<const double *>(ctx->u8lut.p);
    }
    const void *src = conv ? conv : pixels;
    bool src_is_f64 = x.pixel_dtype == SPIHTB_F64 || conv;
    if (x.color == SPIHTB_COLOR_IPT) {
        if (x.C != 3) {
            set_error("IPT colour model needs 3 channels, got %d", x.C);
            return SPIHTB_EINVAL;
        }
        const size_t plane = (size_t)g.h * g.w;
        rc = ctx->ensure(ctx->io2, (size_t)nz * plane * sizeof(double) + 256);
        if (rc) return rc;
        rc = launch_rgb_to_ipt(ctx, pixels, src_is_f64 ? SPIHTB_F64 : x.pixel_dtype, static_cast<double *>(ctx->io2.p),
                               plane, x.B);
        if (rc) return rc;
        src = ctx->io2.p;
        src_is_f64 = true;
    }

    // One launch per level over the whole batch.  (Running the first levels image group by image group,
    // so that a group's float64 approximation planes stay in L2 for the next level, was measured slower
    // on B200: the short launches leave the SMs idle at every kernel boundary.)
    auto make_k = [&](int l, int z0) -> FwdK {
        const bool in_f64 = l > 0 || x.pixel_dtype == SPIHTB_F64 || x.color == SPIHTB_COLOR_IPT || conv;
        const bool in_u8 = !in_f64 && x.pixel_dtype == SPIHTB_U8;
        const size_t esz = in_f64 ? sizeof(double) : (in_u8 ? 1 : sizeof(float));
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
        k.dp = pf ? pf->dp + (size_t)z0 * (g.enc_h / 2) * (g.enc_w / 2) : nullptr;
        k.maxabs = pf ? pf->maxabs : nullptr;  // indexed by z / C: z0 is a multiple of C
        k.NH = g.enc_h / 2;
        k.NW = g.enc_w / 2;
        k.u8lut = u8lut;
        k.tiles_x = k.tiles_y = k.RH = 0;
        k.ntasks = 0;
        k.src_ps = (long long)k.src_h * k.src_w;
        k.dst_ps = (long long)k.bh * k.bw;
        return k;
    };
    auto run_level = [&](int l, int z0, int nzg) -> int {
        const bool in_f64 = l > 0 || x.pixel_dtype == SPIHTB_F64 || x.color == SPIHTB_COLOR_IPT || conv;
        const bool in_u8 = !in_f64 && x.pixel_dtype == SPIHTB_U8;
        const FwdK k = make_k(l, z0);
        const int st = l == 0 ? 0 : 1;
        if (generic) {   // the rest of the bior family: separable kernels with the taps as launch parameters (dwt_gen.cu)
            GenFwdLevel a;
            a.src = k.src;
            a.src_dtype = in_f64 ? SPIHTB_F64 : (in_u8 ? SPIHTB_U8 : SPIHTB_F32);
            a.src_h = k.src_h; a.src_w = k.src_w; a.bh = k.bh; a.bw = k.bw;
            a.dst_ll = k.dst_ll;
            a.coeffs = k.coeffs;
            a.Hc = k.Hc; a.Wc = k.Wc; a.sh = k.sh; a.sw = k.sw; a.mode = k.mode; a.C = k.C; a.last = k.last;
            for (int c = 0; c < 8; ++c) a.scale[c] = k.scale[c];
            a.q = k.q;
            ctx->stage_begin(st);
            const int r = launch_gen_fwd_level(ctx, g.wavelet, a, nzg);
            ctx->stage_end(st);
            return r;
        }
        ctx->stage_begin(st);
        const int r = in_f64 ? launch_level_w<double>(ctx, g.wavelet, k, nzg)
                             : (in_u8 ? launch_level_w<uint8_t>(ctx, g.wavelet, k, nzg)
                                      : launch_level_w<float>(ctx, g.wavelet, k, nzg));
        ctx->stage_end(st);
        return r;
    };
    if (pf) {
        // per-image maxima start from 0.  Every cell byte is written exactly by one of: the epilogue of a
        // level, gap_fill_kernel (cells inside a gap), pyr_fix_kernel (everything that straddles).  The tests
        // set SPIHTB_DEBUG_POISON so that a cell none of them reaches cannot pass on a stale value.
        ctx->stage_begin(2);
        if (getenv("SPIHTB_DEBUG_POISON"))
            SPIHTB_CUDA_CHECK(cudaMemsetAsync(pf->dp, 0xff, (size_t)nz * (g.enc_h / 2) * (g.enc_w / 2), ctx->stream));
        SPIHTB_CUDA_CHECK(cudaMemsetAsync(pf->maxabs, 0, sizeof(uint32_t) * x.B, ctx->stream));
        ctx->stage_end(2);
    }
    // levels 1 and 2 in one kernel (TMA-staged tiles, the level-1 approximation stays in shared memory) when the
    // geometry allows it; otherwise level by level through the float64 scratch planes
    bool fused12 = false;
    {
        const bool in_f64 = x.pixel_dtype == SPIHTB_F64 || x.color == SPIHTB_COLOR_IPT || conv;
        const int sdt = in_f64 ? SPIHTB_F64 : x.pixel_dtype;
        if (!generic) {
            ctx->stage_begin(0);
            rc = launch_forward_fused12(ctx, src, sdt, x, coeffs, static_cast<double *>(ctx->tmpb.p), pf, u8lut, &fused12);
            ctx->stage_end(0);
            if (rc) return rc;
        }
    }
    ctx->last_forward_fused12 = fused12;
    if (fused12) {  // level 3 writes its approximation to the first scratch plane
        rc = ctx->ensure(ctx->tmpa, (size_t)nz * g.band_h[2] * g.band_w[2] * sizeof(double) + 256);
        if (rc) return rc;
    } else {
        rc = ctx->ensure(ctx->tmpa, ll1 + 256);
        if (rc) return rc;
        rc = run_level(0, 0, nz);
        if (rc) return rc;
    }
    bool gap_forked = false;
    {
        GapK gk;
        gk.coeffs = coeffs;
        gk.dp = pf ? pf->dp : nullptr;
        gk.Hc = g.enc_h;
        gk.Wc = g.enc_w;
        gk.NH = g.enc_h / 2;
        gk.NW = g.enc_w / 2;
        gk.nz = nz;
        gk.nrect = 0;
        gk.nrows = 0;
        auto add = [&](int r0, int r1, int c0, int c1) {
            if (r0 >= r1 || c0 >= c1) return;
            const int q = gk.nrect++;
            gk.r0[q] = r0; gk.r1[q] = r1; gk.c0[q] = c0; gk.c1[q] = c1;
            gk.first[q] = gk.nrows;
            gk.nrows += r1 - r0;
        };
        for (int l = 0; l < L; ++l) {
            add(g.band_h[l], g.off_h[l], g.off_w[l], g.off_w[l] + g.band_w[l]);
            add(g.off_h[l], g.off_h[l] + g.band_h[l], g.band_w[l], g.off_w[l]);
        }
        gk.first[gk.nrect] = gk.nrows;
        if (gk.nrect) {
            // independent of every DWT level (disjoint parts of the array and of the cell planes): it runs on the
            // side stream, beside the levels after the first (which leave bandwidth unused), and is joined before the fix-up pass / the end of the call
            SPIHTB_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            SPIHTB_CUDA_CHECK(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
            gap_fill_kernel<<<dim3((gk.nrows + 7) / 8, std::min(nz, 48)), dim3(32, 8), 0, ctx->aux>>>(gk);
            ctx->launches++;
            SPIHTB_CUDA_CHECK(cudaEventRecord(ctx->ev_join, ctx->aux));
            gap_forked = true;
        }
    }
    // per-level launches while a level has many tasks per plane, then all remaining levels of a plane in one CTA
    // (dwt_fwd_tail_kernel); SPIHTB_NO_TAIL=1 keeps one launch per level
    {
        const int F = wavelet_flen(g.wavelet);
        const int nout = (F == 6 ? 2 : 1) * 32 - (F / 2 - 1);
        auto tasks_per_plane = [&](int l) {
            const int rhmax = fw_rhmax(g.band_h[l]);
            return ((g.band_w[l] + nout - 1) / nout) * ((g.band_h[l] + rhmax - 1) / rhmax);
        };
        const bool tail_on = getenv("SPIHTB_NO_TAIL") == nullptr && !generic;
        // opt-in (SPIHTB_TAIL_TASKS=n: levels with at most n tasks per plane go to the tail kernel): measured on B200
        // the tail is slower than one launch per level at every threshold -- a plane's coarse
        // levels are a serial chain of short, latency-bound streams, and one launch per level runs all planes abreast
        const int tail_max = getenv("SPIHTB_TAIL_TASKS") ? atoi(getenv("SPIHTB_TAIL_TASKS")) : 0;
        int l = fused12 ? 2 : 1;
        for (; l < L; ++l) {
            if (tail_on && L - l >= 2 && L - l <= FT_MAXLV && tasks_per_plane(l) <= tail_max) break;
            rc = run_level(l, 0, nz);
            if (rc) return rc;
        }
        if (l < L) {
            // Planes run through their levels independently, so the shared ping-pong planes (whose plane stride
            // changes with the level) cannot be used: a plane at level l + 2 would write where another still reads
            // level l.  Every plane gets two private scratch planes of the tail's largest approximation band.
            const long long S = (((long long)g.band_h[l] * g.band_w[l]) + 1) / 2 * 2;
            rc = ctx->ensure(ctx->tail, (size_t)nz * 2 * S * sizeof(double) + 256);
            if (rc) return rc;
            double *ts = static_cast<double *>(ctx->tail.p);
            FwdTail t;
            t.nlv = L - l;
            for (int i = 0; i < t.nlv; ++i) {
                FwdK &k = t.lv[i];
                k = make_k(l + i, 0);
                if (i > 0) {           // reads what the tail's previous level wrote
                    k.src = ts + ((i - 1) & 1) * S;
                    k.src_ps = 2 * S;
                }
                if (!k.last) {
                    k.dst_ll = ts + (i & 1) * S;
                    k.dst_ps = 2 * S;
                }
            }
            ctx->stage_begin(1);
            switch (g.wavelet) {
                case SPIHTB_WAVELET_BIOR22: rc = launch_tail<SPIHTB_WAVELET_BIOR22>(ctx, t, nz); break;
                case SPIHTB_WAVELET_BIOR44: rc = launch_tail<SPIHTB_WAVELET_BIOR44>(ctx, t, nz); break;
                default: rc = launch_tail<SPIHTB_WAVELET_BIOR68>(ctx, t, nz); break;
            }
            ctx->stage_end(1);
            if (rc) return rc;
        }
    }
    (void)src_is_f64;
    if (gap_forked) SPIHTB_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // join the gap fill
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    if (pf) {
        const FixRect *rects;
        const uint32_t *prefix;
        int nrect;
        uint32_t total;
        rc = upload_fix_rects(ctx, g, fused12, &rects, &prefix, &nrect, &total);
        if (rc) return rc;
        ctx->stage_begin(2);
        rc = launch_pyr_fix(ctx, coeffs, nz, g.enc_h, g.enc_w, pf->dp, rects, prefix, nrect, total);
        ctx->stage_end(2);
        if (rc) return rc;
    }
    return SPIHTB_OK;
}

}  // namespace spihtb
