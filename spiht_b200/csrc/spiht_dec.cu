// SPIHT decoder (replaces src/encoder_decoder.rs:307-454), one CTA per stream.
//
// The decoder's control flow depends on the bits it reads, so record
// boundaries are not known up front.  Each list pass is parsed as follows.
//
//   LIP pass (records "0" | "1 sign"): fully parallel.  Any 0 bit is followed
//   by a record start, so inside the pass "bit p starts a record" <=> the run of
//   1 bits ending at p-1 has even length.  Every thread takes a 32-bit window,
//   the run parity entering each window comes from a scan over 2-state
//   transition functions, and popcounts of the start masks give every record
//   its list index.
//
//   LIS pass (records "0" | B:"1" | A:"1 c c c c", c = "0" | "1 sign"): only a
//   fired A set is longer than one bit.  One thread walks the pass, but it only
//   stops at fired A sets: it ANDs the bitstream with the list's set-type mask
//   (shifted by the bits consumed so far) and jumps from one fired A set to the
//   next with a find-first-set, recording the length of its child bits.  All
//   other work -- reading every entry's own bit at (entry index + prefix sum of
//   the child lengths), the child bits, value writes, list appends through a
//   CTA-wide scan, stable compaction of the retained entries -- is a parallel
//   map over the entries.  The FIFO order of the reference is reproduced
//   generation by generation as in the encoder.
//
//   Refinement pass: a parallel map (entry e reads bit p0 + e).
//
// Values are written straight into the coefficient array, as the reference
// does, so pad bits and duplicated coordinates behave identically.  Decoding
// stops at the first bit position >= 8 * nbytes (pop_bit!, :314-325): an
// action is applied only if every bit it needs lies below that.
//
// Odd LL sizes: the reference's LL-root offspring rule (:50-62) then gives some
// cells two parents, so whole subtrees sit in the lists twice and two list
// entries write the same cell.  Those writes must land in list order (pad bits
// or foreign streams can make the two entries disagree), so writes to cells of
// duplicated subtrees are queued in list order (ordered compaction through the
// CTA scan) and applied by one thread after each parallel step.  Even LL sizes
// (all BASELINE configs) never take that path.
#include <algorithm>
#include <type_traits>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace spihtb {

constexpr int DEC_NT = 512;
constexpr int DEC_NW = DEC_NT / 32;
constexpr int DEC_CH = 2048;           // LIS entries per chain round
constexpr int DEC_PLW = DEC_CH * 9 / 32 + 8;  // staged stream words per round (an entry takes at most 9 bits)
constexpr int DEC_SLACK = 8 * DEC_NT + 64;
constexpr int DEC_DQ = 4 * DEC_NT;     // ordered write queue (odd LL sizes only)
// pipelined rounds: warp w issues on scheduler (b + w) % 4, b = the hardware slot of the CTA's warp 0.  Warp 0 walks; the
// warps that share its scheduler (4, 8, 12) stay idle meanwhile (all fifteen others applying: 3.97 ms per 256 images;
// twelve: 3.86).  Two CTAs share an SM, and on this part the second one's slots start one scheduler further (hardware
// warp ids 0..3 / 16..19 observed), so each CTA also idles the warps on the OTHER walker's scheduler -- residue 1 for
// the first CTA, 3 for the second, told apart by the hardware warp id (3.77 -> 3.72 ms; with the wrong residue the result
// is the same and only the gain is lost: 3.76).  Eight warps apply; that costs nothing (twelve or eight: 3.86 / 3.84).
// Giving the second CTA's WALKER another warp instead was slower every time it was tried (3.92 .. 4.13 ms).
constexpr int DEC_PIPE_NW = DEC_NW / 2;
constexpr int DEC_PIPE_NT = DEC_PIPE_NW * 32;
__device__ __forceinline__ bool dec_pipe_worker2(int wid, int idle2) { return (wid & 3) != 0 && (wid & 3) != idle2; }
__device__ __forceinline__ int dec_pipe_wid2(int wid, int idle2)
{
    // residues at work: idle2 == 1 -> {2, 3}; idle2 == 3 -> {1, 2}
    const int r = wid & 3;
    return ((wid >> 2) << 1) | (idle2 == 1 ? r - 2 : r - 1);
}

struct DecK {
    const uint32_t *in;
    uint64_t in_stride_words;
    const uint64_t *nbytes;
    const int32_t *n;
    int B, C, H, W, ll_h, ll_w;
    KeyFmt kf;
    int32_t *out;
    uint8_t *blk;  // optional block marks [B*C][BH][BW] (64x64 blocks of the array holding a coefficient)
    int BH, BW;
    uint32_t *lip, *lsp, *lis;  // lip: 2 buffers per slot; lis: 3 buffers per slot
    size_t pix_cap, lis_cap;
    unsigned int *counter;
    // decode_with_metadata (encoder_decoder.rs:616-841): one row of 8 int32 per bit position, the decoder state
    // just before that bit is read; rows [B][meta_rows][8], zero-filled by the launcher
    int32_t *meta;
    uint64_t meta_rows;
    int level;                 // number of detail levels (slices.other_slices.len())
    int top_ei, top_ej;        // slices.top_slice.end_i / end_j
    const int32_t *slices;     // device [level][3][4]: start_i, end_i, start_j, end_j in the caller's order da, ad, dd
    int32_t *meta_err;         // set to 1 when the reference would have panicked (slice index out of range)
    int walk_variant;
    int pipe;   // apply round r-1 under the walk of round r
    // lazy zero fill (spihtb_decode_images with the scratch-coefficients option): the launcher zeroes only the
    // top-left fine_h0 x fine_w0 corner of every plane (everything but the finest detail bands); an image's CTA zeroes
    // the rest of its planes itself the first time it meets a coefficient there.
    // fine_h0 == 0: the whole array was zeroed by the launcher.
    // fine_h0 / fine_w0 are even (a 2x2 child block never straddles them); fine_ht / fine_wt are the rows / columns
    // where the finest bands really start (at most one less).  The image zeroes its finest bands when a child block
    // lies past fine_h0 / fine_w0, or when any coefficient at or past fine_ht / fine_wt becomes significant -- so an
    // image with unzeroed finest bands has no non-zero coefficient there at all, marks nothing in blk1, and the level-1 inverse
    // never reads them.
    int fine_h0, fine_w0, fine_ht, fine_wt;
    uint8_t *blk1;   // marks of the finest bands alone (coefficients at or past fine_ht / fine_wt), or null
    uint8_t *l1_any; // [B]: 1 = the image marked something in blk1
};

// ---- decode_with_metadata helpers ----------------------------------------------------------------------
// depth and filter of a coefficient from its position (even LL sizes: every cell has one parent).  The
// reference carries them along the lists (CoefficientMetadata, encoder_decoder.rs:123-151): an LL root has
// depth = level and filter LL; its offspring take the filter from the root's parity (get_offspring_filter) and
// depth - 1; below that the filter is inherited and the depth falls by one per generation (u8, wrapping).
__device__ __forceinline__ void meta_depth_filter(uint32_t i, uint32_t j, uint32_t ll_h, uint32_t ll_w, int level,
                                                  uint32_t &depth, uint32_t &filter)
{
    if (i < ll_h && j < ll_w) {
        depth = (uint32_t)level & 0xffu;
        filter = 0;
        return;
    }
    uint32_t d = 0;
    while (i >= 2 * ll_h || j >= 2 * ll_w) {
        i >>= 1;
        j >>= 1;
        ++d;
    }
    const bool io = i >= ll_h, jo = j >= ll_w;   // the LL root's row / column is odd
    filter = io && jo ? 3u : (!io && jo ? 2u : 1u);
    depth = (uint32_t)(level - 1 - (int)d) & 0xffu;
}
// row `p` of image b's table (get_local_position :593-613 in float32, `as i32` saturating; assign_metadata :663-682)
__device__ __forceinline__ void meta_row(const DecK &p, int32_t *mrows, uint64_t pos, int action, uint32_t k, uint32_t i,
                                         uint32_t j, int n, int32_t value)
{
    uint32_t depth, filter;
    meta_depth_filter(i, j, (uint32_t)p.ll_h, (uint32_t)p.ll_w, p.level, depth, filter);
    float lh, lw;
    if (depth == ((uint32_t)p.level & 0xffu)) {
        lh = __fdiv_rn((float)i, (float)p.top_ei);
        lw = __fdiv_rn((float)j, (float)p.top_ej);
    } else {
        const uint32_t di = (uint32_t)(p.level - 1 - (int)depth) & 0xffu;
        if (di >= (uint32_t)p.level) {
            *p.meta_err = 1;
            return;
        }
        const int32_t *sl = p.slices + ((size_t)di * 3 + (filter - 1)) * 4;
        lh = __fdiv_rn(__fsub_rn((float)i, (float)sl[0]), (float)(uint32_t)(sl[1] - sl[0]));
        lw = __fdiv_rn(__fsub_rn((float)j, (float)sl[2]), (float)(uint32_t)(sl[3] - sl[2]));
    }
    int4 a, c;
    a.x = action;
    a.y = __float2int_rz(__fsub_rn(__fmul_rn(lh, 200000.0f), 100000.0f));
    a.z = __float2int_rz(__fsub_rn(__fmul_rn(lw, 200000.0f), 100000.0f));
    a.w = (int32_t)k;
    c.x = (int32_t)filter;
    c.y = (int32_t)depth;
    c.z = n;
    c.w = value;
    int4 *row = reinterpret_cast<int4 *>(mrows + pos * 8);
    row[0] = a;
    row[1] = c;
}

// per-phase cycle counters of image 0's CTA (tools/dec_phases.py): debug builds only (-DSPIHTB_PROF); release
// builds compile them out and do not export spihtb_debug_dec_prof
#ifdef SPIHTB_PROF
__device__ unsigned long long g_dec_prof[16];
__device__ unsigned long long g_dec_img[1024][4];  // per image: total cycles, walk cycles, smid, start clock
#define DEC_PROF_T0() const long long _t0 = clock64()
#define DEC_PROF_ADD(slot)                                                        \
    do {                                                                          \
        if (tid == 0 && b == 0) g_dec_prof[slot] += (unsigned long long)(clock64() - _t0); \
    } while (0)
#define DEC_PROF_CNT(slot, v)                                  \
    do {                                                       \
        if (tid == 0 && b == 0) g_dec_prof[slot] += (v);       \
    } while (0)
#define DEC_PROF_MARK(name) const long long name = clock64()
#define DEC_PROF_SINCE(slot, name)                                                        \
    do {                                                                                  \
        if (tid == 0 && b == 0) g_dec_prof[slot] += (unsigned long long)(clock64() - name); \
    } while (0)
#else
#define DEC_PROF_T0() do { } while (0)
#define DEC_PROF_ADD(slot) do { } while (0)
#define DEC_PROF_CNT(slot, v) do { } while (0)
#define DEC_PROF_MARK(name) do { } while (0)
#define DEC_PROF_SINCE(slot, name) do { } while (0)
#endif

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct BitRow {
    const uint32_t *row;
    uint64_t nwords;
    uint64_t nbits;
    __device__ __forceinline__ uint32_t word(uint64_t idx) const { return idx < nwords ? __ldg(row + idx) : 0u; }
    __device__ __forceinline__ uint64_t get64(uint64_t p) const
    {
        const uint64_t idx = p >> 5;
        const int sh = (int)(p & 31);
        const uint32_t w0 = word(idx), w1 = word(idx + 1), w2 = word(idx + 2);
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        return (uint64_t)lo | ((uint64_t)hi << 32);
    }
    __device__ __forceinline__ uint32_t get32(uint64_t p) const
    {
        const uint64_t idx = p >> 5;
        return __funnelshift_r(word(idx), word(idx + 1), (int)(p & 31));
    }
    __device__ __forceinline__ uint32_t bit(uint64_t p) const { return (word(p >> 5) >> (p & 31)) & 1u; }
};

// Is (y,x) inside a subtree that the reference's lists hold twice?  Cells with
// two LL-root parents: row ll_h (odd ll_h) x cols [ll_w, 2 ll_w) and col ll_w
// (odd ll_w) x rows [ll_h, 2 ll_h); their descendants follow the dyadic rule.
__device__ __forceinline__ bool in_dup_subtree(uint32_t y, uint32_t x, uint32_t ll_h, uint32_t ll_w)
{
    for (int t = 0; t < 32; ++t) {
        const uint32_t yy = y >> t, xx = x >> t;
        if (yy < ll_h && xx < ll_w) return false;
        if ((ll_h & 1u) && yy == ll_h && xx >= ll_w && xx < 2 * ll_w) return true;
        if ((ll_w & 1u) && xx == ll_w && yy >= ll_h && yy < 2 * ll_h) return true;
    }
    return false;
}

// set_bit (encoder_decoder.rs:14-29): set / clear bit n of the magnitude, keep the sign
__device__ __forceinline__ void refine_cell(int32_t *cell, int n, uint32_t bit)
{
    const int32_t x = *cell;
    const uint32_t m = 1u << n;
    uint32_t mag = absu(x);
    mag = bit ? (mag | m) : (mag & ~m);
    *cell = x >= 0 ? (int32_t)mag : -(int32_t)mag;
}

// One thread: apply queued writes in list order.  (Cells in a duplicated subtree; their blocks are marked by
// the caller's mark() at queueing time.)
__device__ void apply_queue(const uint2 *dq, uint32_t cnt, const KeyFmt &kf, int32_t *rec, uint32_t H, uint32_t W,
                            int n)
{
    for (uint32_t q = 0; q < cnt; ++q) {
        const uint2 op = dq[q];
        uint32_t k, i, j;
        key_unpack(kf, op.x, k, i, j);
        int32_t *cell = rec + ((size_t)k * H + i) * W + j;
        if (op.x >> 31)
            refine_cell(cell, n, op.y);
        else
            *cell = (int32_t)op.y;
    }
}

// ---- barrier and scan for the threads that apply a round: the whole CTA, or (PIPE) warps 1..15 on named barrier 1
// while warp 0 walks the next round
template <bool PIPE>
__device__ __forceinline__ void dec_sync()
{
    if (PIPE)
        asm volatile("bar.sync 1, %0;" ::"n"(DEC_PIPE_NT) : "memory");
    else
        __syncthreads();
}
template <bool PIPE>
__device__ __forceinline__ uint64_t dec_exscan(uint64_t v, uint64_t *warp_tot, uint64_t &total, int idle2 = 1)
{
    if (!PIPE) return block_exscan<DEC_NT>(v, warp_tot, total);
    constexpr int NW = DEC_PIPE_NW;
    const int lane = threadIdx.x & 31, wid = dec_pipe_wid2((int)(threadIdx.x >> 5), idle2);
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    dec_sync<true>();  // previous users of warp_tot are done
    if (lane == 31) warp_tot[wid] = inc;
    dec_sync<true>();
    uint64_t base = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        const uint64_t t = warp_tot[q];
        if (q < wid) base += t;
        tot += t;
    }
    total = tot;
    return base + inc - v;
}

// ---- lazy zero fill: zero the finest detail bands of one image (rows from fh on whole, rows above from column fw
// on); t = index of the calling thread among the nt that call.  Out of line: it runs at most twice per image, and
// inlined into the hot loops its address arithmetic cost the decoder registers it does not have (spills)
__device__ __noinline__ void dec_zero_fine(int32_t *rec, uint32_t C, uint32_t H, uint32_t W, uint32_t fh, uint32_t fw,
                                           int t, int nt)
{
    for (uint32_t kk = 0; kk < C; ++kk) {
        int32_t *pl = rec + (size_t)kk * H * W;
        const size_t lo = (size_t)fh * W, hi = (size_t)H * W;
        // 16-byte stores over the aligned middle, scalar stores at the ends
        int32_t *z0 = pl + lo;
        const size_t head = min(hi - lo, (size_t)((16u - (uint32_t)((uintptr_t)z0 & 15u)) & 15u) >> 2);
        const size_t nq = (hi - lo - head) >> 2, done = head + (nq << 2);
        if ((size_t)t < head) z0[t] = 0;
        int4 *zq = reinterpret_cast<int4 *>(z0 + head);
        for (size_t q = t; q < nq; q += nt) zq[q] = make_int4(0, 0, 0, 0);
        if (done + t < hi - lo) z0[done + t] = 0;
        for (uint32_t r = (uint32_t)t >> 5; r < fh; r += nt / 32)
            for (uint32_t cc = fw + (t & 31); cc < W; cc += 32) pl[(size_t)r * W + cc] = 0;
    }
}

// ---- LIP parse: 2-state automaton over the bits of a pass -------------------
// state s = 1: the next bit starts a record; s = 0: the next bit is a sign bit.
// next(s, b) = !(s & b).  A transition function over a bit window is stored as
// 2 bits: bit x = state after the window when entered in state x.
__device__ __forceinline__ uint32_t lip_fn_compose(uint32_t first, uint32_t then)
{
    return ((then >> (first & 1u)) & 1u) | (((then >> ((first >> 1) & 1u)) & 1u) << 1);
}

// Exclusive scan of transition functions over the CTA, applied to `s_in`
// (the state entering thread 0's window).  Returns the state entering this
// thread's window and, in `s_out`, the state after the last window.
__device__ __forceinline__ uint32_t lip_state_scan(uint32_t fn, uint32_t s_in, uint32_t *warp_fn, uint32_t &s_out)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = fn;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc = lip_fn_compose(t, inc);
    }
    __syncthreads();  // previous users of warp_fn are done
    if (lane == 31) warp_fn[wid] = inc;
    __syncthreads();
    uint32_t s = s_in, s_mine = s_in;
#pragma unroll
    for (int q = 0; q < DEC_NW; ++q) {
        if (q == wid) s_mine = s;
        s = (warp_fn[q] >> s) & 1u;
    }
    s_out = s;
    // state entering this lane = exclusive prefix within the warp applied to the warp's entry state
    const uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? s_mine : ((prev >> s_mine) & 1u);
}

// ---- the serial walk over the LIS records of one round (one thread).  Out of line so that its register
// allocation does not depend on the rest of the kernel (which runs at the 64-register cap): the loop is pure
// instruction latency, and a spilled value in it costs more than the call.
// a_sw: staged stream words (MSB-first, zero from the end of the stream on), q: table address of the round's first
// bit (s_lp + pos mod 32; the table is 32-byte aligned, so q mod 32 is the shift of the stream window), a_x: per-entry
// child-length bytes to fill, a_t: set-type words (MSB-first), cnt: entries of the round.  Returns the bits consumed
// (every entry one, every fired A set its child bits).
__device__ __noinline__ uint32_t lis_walk(uint32_t a_sw, uint32_t q, uint32_t a_x, uint32_t a_t, uint32_t cnt)
{
    const uint32_t a_swe = a_sw + (DEC_PLW - 1) * 4;  // last staged word
    const uint32_t a_x31 = a_x + 31u;
    const uint32_t q0 = q;
    uint32_t r0 = lds_u32(a_sw), r1 = lds_u32(a_sw + 4), r2 = lds_u32(a_sw + 8);
    a_sw += 12;
    uint32_t t0 = lds_u32(a_t), t1 = lds_u32(a_t + 4), t2 = lds_u32(a_t + 8);
    a_t += 12;
    uint32_t e = 0;                                 // entries consumed in this round
    uint32_t nextq = (q & ~31u) + 32, nexte = 32;   // where the windows run out of their first word
#pragma unroll 1
    while (e < cnt) {
        const uint32_t w0 = __funnelshift_l(r1, r0, q);
        const uint32_t tt = __funnelshift_l(t1, t0, e);
        const uint32_t m = w0 & tt & 0xffffff00u;
        uint32_t hb;
        asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(m));  // 31 - distance of the fired set
        const uint32_t at = q - hb;                     // table address of the set, minus 31
        const uint32_t len = lds_u8(at + 31u);
        const uint32_t e1 = e - hb + 32u;
        if (__builtin_expect(m != 0, 1)) {
            sts_u8(a_x31 + e - hb, len);
            e = e1;
            q = at + len + 32u;
        } else {  // no fired A set within the next 24 entries
            const uint32_t sh = min(24u, cnt - e);
            e += sh;
            q += sh;
        }
        if (__builtin_expect(q >= nextq || e >= nexte, 0)) {
            if (q >= nextq) {
                nextq += 32;
                r0 = r1;
                r1 = r2;
                r2 = lds_u32(min(a_sw, a_swe));
                a_sw += 4;
            }
            if (e >= nexte) {
                nexte += 32;
                t0 = t1;
                t1 = t2;
                t2 = lds_u32(a_t);
                a_t += 4;
            }
        }
    }
    return q - q0;
}

// ---- the walk as it ships: the same steps with NO branch inside.  The loop above spends a third of its time in
// control flow: a lone warp pays the full fetch bubble of every BRA / BSSY / BSYNC, and a step has three.  Here the
// word rotations are selects (the word after the window is loaded every step, ahead of need), a step past the
// round's end is a no-op (the set-type window is zero there and cnt - e is 0), so UNROLL steps run as straight-line
// code and the loop tests its exit once per UNROLL steps.  The stream window is aligned to the position right after
// the fired bit (known as soon as the set is located) so that only one shift -- by the child length -- sits behind
// the length load.  35 instead of 25 instructions per step, 78 instead of 126 cycles: 5.8 -> 3.4 M cycles per 1024^2
// image at 0.5 bpp (without the alignment trick, UNROLL 1 / 2 / 4 / 8: 4.9 / 4.0 / 3.8 / 3.6 M).
// Measured and dropped on the way: windows from byte-granular tables in shared memory instead of registers (28
// instructions, but a second load on the dependency chain: 4.6 M cycles); the walker warp of the SM's second CTA on
// another scheduler (decode kernel 4.30 -> 4.46 ms); every thread walking one 32-entry segment from a candidate
// position, resolved by following the true path through the exits (bit-exact; 512 walkers saturate the schedulers
// that one walker leaves idle: 5.3 M cycles).
template <int UNROLL>
__device__ __noinline__ uint32_t lis_walk_pa(uint32_t a_sw, uint32_t q, uint32_t a_x, uint32_t a_t, uint32_t cnt)
{
    const uint32_t a_x31 = a_x + 31u;
    const uint32_t q0 = q;
    uint32_t r0 = lds_u32(a_sw), r1 = lds_u32(a_sw + 4), r2 = lds_u32(a_sw + 8);
    a_sw += 12;
    uint32_t t0 = lds_u32(a_t), t1 = lds_u32(a_t + 4), t2 = lds_u32(a_t + 8);
    a_t += 12;
    uint32_t e = 0;
    uint32_t A = q, le = 0;   // next bit: A + le (le <= 8); the window registers hold the word of A in r0
    uint32_t nextq = (q & ~31u) + 32, nexte = 32;
#pragma unroll 1
    while (e < cnt) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t rn = lds_u32(a_sw);
            const uint32_t tn = lds_u32(a_t);
            const uint32_t whi = __funnelshift_l(r1, r0, A), wlo = __funnelshift_l(r2, r1, A);
            const uint32_t tts = __funnelshift_l(t1, t0, e) & 0xffffff00u;
            const uint32_t w0 = __funnelshift_l(wlo, whi, le);
            const uint32_t m = w0 & tts;
            uint32_t hb;
            asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(m));
            const uint32_t qq = A + le;
            const uint32_t at = qq - hb;
            const uint32_t len = lds_u8(at + 31u);
            const bool f = m != 0;
            const uint32_t sh = min(24u, cnt - e);
            if (f) sts_u8(a_x31 + e - hb, len);
            e = f ? e - hb + 32u : e + sh;
            A = f ? at + 32u : qq + sh;
            le = f ? len : 0u;
            const bool rq = A >= nextq, re = e >= nexte;
            r0 = rq ? r1 : r0;
            r1 = rq ? r2 : r1;
            r2 = rq ? rn : r2;
            a_sw += rq ? 4u : 0u;
            nextq += rq ? 32u : 0u;
            t0 = re ? t1 : t0;
            t1 = re ? t2 : t1;
            t2 = re ? tn : t2;
            a_t += re ? 4u : 0u;
            nexte += re ? 32u : 0u;
        }
    }
    return A + le - q0;
}

// META: also fill the decode_with_metadata table.  The parse then runs one bit position further than the data
// (limit + 1): the reference assigns the row of the bit it is about to read before it finds the data exhausted,
// so the table has one more row than there are bits.  That phantom bit can change no coefficient -- a record that
// needs a bit beyond it is cut exactly as at the real end -- except through a refinement, which is therefore
// applied only below the real limit.
template <bool META>
__global__ void __launch_bounds__(DEC_NT, META ? 1 : 2) spiht_decode_kernel(const DecK p)
{
    __shared__ uint32_t s_tmask[DEC_CH / 32 + 4];  // A sets with offspring (a fired record carries child bits)
    __shared__ __align__(16) uint8_t s_x2[2][DEC_CH];           // child-bit length of fired A sets, 0 elsewhere (double-buffered)
    __shared__ uint32_t s_grp2[2][DEC_CH / 32];    // exclusive prefix of s_x per 32 entries
    __shared__ uint32_t s_cnt[4];                  // list counters handed back by the applying warps
    __shared__ uint32_t s_fine_done;               // lazy zero fill: the image's finest detail bands are zeroed
    __shared__ uint64_t s_wtot[DEC_NW];
    __shared__ uint32_t s_wfn[DEC_NW];
    __shared__ uint64_t s_chain_p;
    // staged for the chain: the stream words the round can reach, bit-reversed (MSB-first), and for every
    // bit position the child-bit length of a fired A record whose "1" sits there
    __shared__ uint32_t s_sw2[2][DEC_PLW];
    __shared__ __align__(32) uint8_t s_lp[DEC_PLW * 32];
    __shared__ uint32_t s_na;  // A sets with offspring in the round
    __shared__ uint8_t s_len[256];  // child-bit length of a fired A record from its next 8 bits
    __shared__ int s_img;
    __shared__ uint2 s_dq[DEC_DQ];  // {cell key | refine flag << 31, value or bit}

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const KeyFmt kf = p.kf;
    const uint32_t H = p.H, W = p.W, ll_h = p.ll_h, ll_w = p.ll_w, C = p.C;
    uint32_t *lipA = p.lip + (size_t)blockIdx.x * 2 * p.pix_cap;
    uint32_t *lipB = lipA + p.pix_cap;
    uint32_t *lsp = p.lsp + (size_t)blockIdx.x * p.pix_cap;
    uint32_t *R = p.lis + (size_t)blockIdx.x * 3 * p.lis_cap;
    uint32_t *G0 = R + p.lis_cap;
    uint32_t *G1 = G0 + p.lis_cap;
    const bool has_dups = ((ll_h | ll_w) & 1u) != 0;
    if (tid < 256) {  // index: the 8 bits after the fire bit, first bit in bit 7
        uint32_t cb = (uint32_t)tid << 24, len = 4;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (cb >> 31) {
                cb <<= 2;
                ++len;
            } else {
                cb <<= 1;
            }
        }
        s_len[tid] = (uint8_t)len;
    }
    if (tid == 0) s_na = 0;
    __shared__ int s_idle2;
    if (tid == 0) {
        uint32_t hw;
        asm("mov.u32 %0, %%warpid;" : "=r"(hw));
        s_idle2 = hw < 16 ? 1 : 3;
    }
    __syncthreads();
    const int idle2 = s_idle2;
    // the walk may run past the end of a short stream into words / table bytes no round has staged yet
    for (int w = tid; w < 2 * DEC_PLW; w += DEC_NT) (&s_sw2[0][0])[w] = 0;
    for (int w = tid; w < DEC_PLW * 8; w += DEC_NT) reinterpret_cast<uint32_t *>(s_lp)[w] = 0;
    // bits per thread in a LIP round; the ordered write queue bounds it when cells can be duplicated
    const int BPT = has_dups ? 4 : 32;

    for (;;) {
        if (tid == 0) s_img = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int b = s_img;
        __syncthreads();
        if (b >= p.B) break;

        BitRow br;
        br.row = p.in + (size_t)b * p.in_stride_words;
        br.nwords = p.in_stride_words;
        br.nbits = p.nbytes[b] * 8ull;
        if (br.nbits > br.nwords * 32ull) br.nbits = br.nwords * 32ull;
        const uint64_t data_limit = br.nbits;
        const uint64_t limit = META ? data_limit + 1 : data_limit;
        [[maybe_unused]] int32_t *mrows = META ? p.meta + (size_t)b * p.meta_rows * 8 : nullptr;
        int32_t *rec = p.out + (size_t)b * C * H * W;
        int n = p.n[b];
        n = n < 0 ? 0 : (n > 31 ? 31 : n);
        // ---- list initialisation (encoder_decoder.rs:329-348)
        uint32_t *lip = lipA, *lip_alt = lipB;
        const uint32_t T0 = ll_h * ll_w * C;
        uint32_t lip_len = T0, lsp_len = 0, r_len = 0;
        for (uint32_t base = 0; base < T0; base += DEC_NT) {
            const uint32_t t = base + tid;
            const bool valid = t < T0;
            uint32_t key = 0;
            bool root = false;
            if (valid) {
                const uint32_t k = t % C, ij = t / C;
                const uint32_t j = ij % ll_w, i = ij / ll_w;
                key = key_pack(kf, k, i, j);
                lip[t] = key;
                root = ((i | j) & 1u) != 0;
            }
            uint64_t tot;
            const uint64_t ex = block_exscan<DEC_NT>(root ? 1ull : 0ull, s_wtot, tot);
            if (root) R[r_len + (uint32_t)ex] = 0x80000000u | key;
            r_len += (uint32_t)tot;
        }
        __syncthreads();

        uint64_t pos = 0;  // uniform: next unread bit
        if (tid == 0) s_fine_done = p.fine_h0 == 0 ? 1u : 0u;   // the finest detail bands of this image are zeroed
        DEC_PROF_MARK(_timg);
#ifdef SPIHTB_PROF
        unsigned long long walk_cycles = 0, last_walk = 0, prev_tb = 0;
#endif
        for (; pos < limit; --n) {
            const int32_t basev = n == 0 ? 1 : (int32_t)((1u << (n - 1)) + (1u << n));
            const uint32_t lsp_len0 = lsp_len;

            // ================= LIP pass (encoder_decoder.rs:355-377)
            {
                uint32_t done_e = 0, keep = 0;
                while (done_e < lip_len && pos < limit) {
                    DEC_PROF_T0();
                    // this thread's window: BPT bits at wp; bits at or past `limit` are not data
                    const uint64_t wp = pos + (uint64_t)tid * BPT;
                    uint32_t nv = wp < limit ? (uint32_t)min((uint64_t)BPT, limit - wp) : 0u;
                    const uint32_t vmask = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
                    const uint32_t wbits = (nv ? br.get32(wp) : 0u) & vmask;
                    // transition function of the window (all BPT bits; invalid ones read as 0)
                    uint32_t fn;
                    {
                        const uint32_t full = BPT >= 32 ? 0xffffffffu : ((1u << BPT) - 1u);
                        if ((wbits & full) == full) {
                            fn = (BPT & 1) ? 0x1u : 0x2u;  // all ones: the state flips BPT times
                        } else {
                            // after the highest 0 bit the state is 1; then it flips once per trailing 1
                            const uint32_t inv = ~wbits & full;
                            const int top0 = 31 - __clz((int)inv);
                            const uint32_t ones = (uint32_t)(BPT - 1 - top0);
                            const uint32_t s = (ones & 1u) ? 0u : 1u;
                            fn = s | (s << 1);
                        }
                    }
                    uint32_t s_last;
                    uint32_t s = lip_state_scan(fn, 1u, s_wfn, s_last);
                    // start mask of the window
                    uint32_t smask = 0;
                    for (int i = 0; i < BPT; ++i) {
                        smask |= s << i;
                        s = ((wbits >> i) & s & 1u) ^ 1u;
                    }
                    smask &= vmask;
                    const uint32_t nrec = __popc(smask);
                    uint64_t tot;
                    const uint32_t e0 = (uint32_t)block_exscan<DEC_NT>((uint64_t)nrec, s_wtot, tot);
                    const uint32_t rem = lip_len - done_e;
                    const uint32_t take = min((uint32_t)tot, rem);
                    // records of this window that belong to the pass: the first `mine`
                    const uint32_t mine = e0 >= take ? 0u : min(nrec, take - e0);
                    uint32_t use = smask;
                    for (uint32_t q = nrec; q > mine; --q) use &= ~(0x80000000u >> __clz((int)use));
                    // newly significant: start bit set and the sign bit below the limit
                    const uint32_t sigm = use & wbits;
                    uint32_t okm = sigm;
                    if (nv && nv <= 32 && wp + nv >= limit) {
                        // the window holds the last valid bit: a record starting there has no sign bit
                        okm &= ~(1u << (nv - 1));
                    }
                    if constexpr (META) {
                        // ... and with the phantom position appended, neither has the record that starts on the
                        // last real bit: its sign would be the phantom
                        if (data_limit >= 1 && data_limit - 1 >= wp && data_limit - 1 - wp < 32)
                            okm &= ~(1u << (uint32_t)(data_limit - 1 - wp));
                    }
                    const uint32_t nsig = __popc(okm), nkeep = __popc(use & ~wbits);
                    // sign bits: bit i+1 of the window, the last one from the next word
                    const uint32_t nextbits = (uint32_t)(br.get64(wp) >> 1);
                    uint32_t ndef = 0, defm = 0;
                    if (has_dups) {
                        uint32_t mm = okm;
                        uint32_t idx = e0;
                        uint32_t um = use;
                        while (um) {
                            const int bpos = __ffs((int)um) - 1;
                            um &= um - 1;
                            if ((mm >> bpos) & 1u) {
                                const uint32_t key = lip[done_e + idx];
                                uint32_t k, i, j;
                                key_unpack(kf, key, k, i, j);
                                if (in_dup_subtree(i, j, ll_h, ll_w)) {
                                    defm |= 1u << bpos;
                                    ++ndef;
                                }
                            }
                            ++idx;
                        }
                    }
                    const uint64_t pack = (uint64_t)nsig | ((uint64_t)nkeep << 20) | ((uint64_t)ndef << 40);
                    uint64_t tot2;
                    const uint64_t ex = block_exscan<DEC_NT>(pack, s_wtot, tot2);
                    {
                        uint32_t os = lsp_len + (uint32_t)(ex & 0xfffff);
                        uint32_t ok = keep + (uint32_t)((ex >> 20) & 0xfffff);
                        uint32_t od = (uint32_t)(ex >> 40);
                        uint32_t idx = done_e + e0;
                        uint32_t um = use;
                        while (um) {
                            const int bpos = __ffs((int)um) - 1;
                            um &= um - 1;
                            const uint32_t key = lip[idx++];
                            if constexpr (META) {
                                uint32_t k, i, j;
                                key_unpack(kf, key, k, i, j);
                                const int32_t cur = rec[((size_t)k * H + i) * W + j];
                                meta_row(p, mrows, wp + bpos, 0, k, i, j, n, cur);
                                if (((wbits >> bpos) & 1u) && wp + bpos + 1 < limit)
                                    meta_row(p, mrows, wp + bpos + 1, 1, k, i, j, n, cur);
                            }
                            if ((wbits >> bpos) & 1u) {
                                if ((okm >> bpos) & 1u) {
                                    uint32_t k, i, j;
                                    key_unpack(kf, key, k, i, j);
                                    const int32_t val = ((nextbits >> bpos) & 1u) ? basev : -basev;
                                    if ((defm >> bpos) & 1u)
                                        s_dq[od++] = make_uint2(key, (uint32_t)val);
                                    else
                                        rec[((size_t)k * H + i) * W + j] = val;
                                    lsp[os++] = key;
                                }
                            } else {
                                lip_alt[ok++] = key;
                            }
                        }
                    }
                    if (has_dups) {
                        __syncthreads();
                        if (tid == 0) apply_queue(s_dq, (uint32_t)(tot2 >> 40), kf, rec, H, W, n);
                        __syncthreads();
                    }
                    lsp_len += (uint32_t)(tot2 & 0xfffff);
                    keep += (uint32_t)((tot2 >> 20) & 0xfffff);
                    done_e += take;
                    // next unread bit: the start of record `take` if the window holds it, else the
                    // bit after the window (plus the sign bit the window's last record still owes)
                    if ((uint32_t)tot > take) {
                        // position of start number `take`: in the thread with e0 <= take < e0 + nrec
                        if (e0 <= take && take < e0 + nrec) {
                            uint32_t um = smask;
                            for (uint32_t q = e0; q < take; ++q) um &= um - 1;
                            s_chain_p = wp + (uint64_t)(__ffs((int)um) - 1);
                        }
                        __syncthreads();
                        pos = s_chain_p;
                        __syncthreads();
                    } else {
                        pos = pos + (uint64_t)DEC_NT * BPT + (s_last ? 0u : 1u);
                    }
                    DEC_PROF_ADD(0);
                    DEC_PROF_CNT(8, 1);
                }
                if (done_e < lip_len) break;  // the stream ended inside the pass
                uint32_t *t = lip;
                lip = lip_alt;
                lip_alt = t;
                lip_len = keep;
            }
            if (pos >= limit) break;

            // ================= LIS pass (encoder_decoder.rs:379-436), generation by generation
            bool ended = false;
            {
                uint32_t *cur = R, *nxt = G0;
                uint32_t cur_len = r_len, rkeep = 0;
                int gen = 0;
                const bool pipe = !META && !has_dups && p.pipe != 0;
                while (cur_len > 0 && !ended) {
                    uint32_t nxt_len = 0;
                    // One round = up to DEC_CH entries of the generation: stage (all threads), walk (one thread), apply (parallel).
                    // Pipelined form (even LL sizes, no metadata table): the staged words, child lengths and group prefixes are
                    // double-buffered, and while thread 0 walks round r eight warps on the schedulers no walker uses apply
                    // round r-1 behind a named barrier of their own; the list counters they advance come back through
                    // shared memory.
                    auto stage_round = [&](uint32_t ebase, uint32_t cnt, int buf) {
                        uint8_t *sx = s_x2[buf];
                        uint32_t *ssw = s_sw2[buf];
                        DEC_PROF_T0();
                        // set-type mask of this round's entries; clear the child-length bytes
                        constexpr int TMI = (DEC_CH / 32 + 3 + DEC_NW - 1) / DEC_NW;  // iterations per warp
                        uint32_t tkey[TMI];
#pragma unroll
                        for (int it = 0; it < TMI; ++it) {  // all loads in flight before the first is used
                            const uint32_t e = (uint32_t)(wid + it * DEC_NW) * 32 + lane;
                            tkey[it] = e < cnt ? cur[ebase + e] : 0u;
                        }
#pragma unroll
                        for (int it = 0; it < TMI; ++it) {
                            const uint32_t w = (uint32_t)(wid + it * DEC_NW);
                            if (w >= (cnt + 31) / 32 + 3) break;
                            const uint32_t e = w * 32 + lane;
                            const uint32_t key = tkey[it];
                            bool a_with_children = false;
                            if (key >> 31) {
                                uint32_t k, i, j, ci, cj;
                                key_unpack(kf, key, k, i, j);
                                a_with_children = offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj);
                            }
                            const uint32_t m = __ballot_sync(0xffffffffu, a_with_children);
                            if (lane == 0) {
                                s_tmask[w] = __brev(m);  // MSB-first for the chain
                                if (m) atomicAdd(&s_na, (uint32_t)__popc(m));
                            }
                            if (e < DEC_CH) sx[e] = 0;
                        }
                        __syncthreads();
                        // stage what the chain reads (all threads): the stream words the round can reach,
                        // bit-reversed (MSB-first) and cleared from `limit` on, then for every bit position
                        // the child-bit length of a fired A record whose "1" sits there
                        {
                            const uint64_t remaining = pos < limit ? limit - pos : 0ull;
                            const uint64_t reach = (uint64_t)cnt + 8ull * s_na;
                            const uint32_t bound = (uint32_t)(reach < remaining ? reach : remaining);
                            const uint32_t npw = min((uint32_t)DEC_PLW - 2, (((uint32_t)(pos & 31) + bound + 31) >> 5) + 3);
                            const uint64_t wbase = pos >> 5;
                            for (uint32_t w = tid; w < npw + 2; w += DEC_NT) {
                                const uint64_t gw = wbase + w;
                                uint32_t x = __brev(br.word(gw));
                                const uint64_t first = gw << 5;  // stream position of the word's first bit
                                if (first + 32 > limit) x = first >= limit ? 0u : (x & ~(0xffffffffu >> (uint32_t)(limit - first)));
                                ssw[w] = x;
                            }
                            __syncthreads();
                            for (uint32_t w = wid; w < npw; w += DEC_NW) {
                                // lane l: record whose "1" is bit l (MSB-first) of word w; its child bits follow
                                const uint32_t v = __funnelshift_lc(ssw[w + 1], ssw[w], lane + 1) >> 24;
                                s_lp[w * 32 + lane] = s_len[v];
                            }
                        }
                        __syncthreads();
                        DEC_PROF_ADD(1);
                    };
                    auto walk_round = [&](uint32_t cnt, int buf) {
                        uint8_t *sx = s_x2[buf];
                        uint32_t *ssw = s_sw2[buf];
                        // ---- chain: one thread jumps from fired A set to fired A set
                        if (tid == 0) {
                            DEC_PROF_MARK(_tc);
                            // Everything is kept MSB-first so that the next fired A set is one
                            // count-leading-zeros away.  r0, r1 (r2 prefetched): stream words, sS = consumed
                            // bits of r0; t0, t1 (t2): set-type words, sT likewise; all from shared memory
                            // (32-bit addresses).  A step looks at the next 24 entries, so it consumes at most
                            // 24 + 8 bits and one word rotation per step is enough.  The record length is one
                            // byte load from the per-position table, under whose latency the bookkeeping
                            // issues; bits from `limit` on are staged as zeros, so the walk needs no end test
                            // of its own: it runs out of entries one bit per entry.
                            // A lone warp issues roughly one instruction every four cycles here, so the
                            // step is written for instruction count: the stream position q is kept as the
                            // table address itself (the table is 32-byte aligned, so q mod 32 is the shift
                            // of the stream window and the funnel shift wraps it), the entry index e is the
                            // shift of the set-type window, and both word rotations and the step without a
                            // fired set sit behind one rarely taken branch.
                            const uint32_t wa = (uint32_t)__cvta_generic_to_shared(ssw);
                            const uint32_t wq = (uint32_t)__cvta_generic_to_shared(s_lp) + (uint32_t)(pos & 31);
                            const uint32_t wx = (uint32_t)__cvta_generic_to_shared(sx);
                            const uint32_t wt = (uint32_t)__cvta_generic_to_shared(s_tmask);
                            const uint32_t used = p.walk_variant == 0 ? lis_walk(wa, wq, wx, wt, cnt)
                                                                      : lis_walk_pa<8>(wa, wq, wx, wt, cnt);
                            // bits consumed: every entry one, every fired A set its child bits
                            s_chain_p = pos + used;
                            s_na = 0;
                            DEC_PROF_SINCE(2, _tc);
#ifdef SPIHTB_PROF
                            last_walk = (unsigned long long)(clock64() - _tc);
                            walk_cycles += last_walk;
#endif
                            DEC_PROF_CNT(11, 1);
                        }
                    };
                    auto apply_round = [&](auto pipe_c, uint32_t ebase, uint32_t cnt, uint64_t p_base, int buf) {
                        constexpr bool PIPE = decltype(pipe_c)::value;
                        constexpr int NT = PIPE ? DEC_PIPE_NT : DEC_NT;
                        // index among the threads at work here
                        const int tid = PIPE ? dec_pipe_wid2((int)(threadIdx.x >> 5), idle2) * 32 + (int)(threadIdx.x & 31) : (int)threadIdx.x;
                        const uint8_t *sx = s_x2[buf];
                        const uint32_t *ssw = s_sw2[buf];
                        uint32_t *sgrp = s_grp2[buf];
                        // ---- prefix of the child lengths per 32 entries
                        {
                            uint32_t v = 0;
                            if ((uint32_t)tid < (cnt + 31) / 32) {
                                const uint32_t *xw = reinterpret_cast<const uint32_t *>(sx) + tid * 8;
#pragma unroll
                                for (int q = 0; q < 8; ++q) v += __dp4a(xw[q], 0x01010101u, 0u);
                            }
                            static_assert(DEC_CH / 32 <= NT, "one thread per 32-entry group");
                            uint64_t tot;
                            const uint64_t ex = dec_exscan<PIPE>((uint64_t)v, s_wtot, tot, idle2);
                            if (tid < DEC_CH / 32) sgrp[tid] = (uint32_t)ex;
                        }
                        dec_sync<PIPE>();
                        DEC_PROF_MARK(_tb);
                        // ---- every entry: its own bit, then the fired sets' records
                        for (uint32_t eb = 0; eb < cnt; eb += NT) {
                            const uint32_t e = eb + tid;
                            [[maybe_unused]] const bool lazy_live = !has_dups && s_fine_done == 0;   // uniform per batch
                            const bool valid = e < cnt;
                            // bit position: entries before e take one bit each plus the child bits of fired A sets
                            const uint32_t xe = valid ? sx[e] : 0u;
                            uint32_t inc = xe;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                                if (lane >= d) inc += t;
                            }
                            const uint64_t pe = p_base + e + (valid ? sgrp[e >> 5] : 0u) + (inc - xe);
                            const bool avail = valid && pe < limit;
                            const uint32_t key = valid ? cur[ebase + e] : 0u;
                            // the entry's bit and the 32 after it, from the staged words (MSB-first; staged from the
                            // word of p_base on, so every bit the round can reach is there)
                            const uint32_t so = (uint32_t)(p_base & 31) + (uint32_t)(pe - p_base);
                            const bool fired = avail && ((ssw[so >> 5] << (so & 31)) >> 31);
                            uint32_t k = 0, i = 0, j = 0, ci = 0, cj = 0;
                            uint32_t nlsp = 0, nlip = 0, nnext = 0, sigmask = 0, sgnmask = 0, nread = 0;
                            uint32_t ndef = 0, defmask = 0;
                            const bool isA = (key >> 31) != 0;
                            if constexpr (META) {
                                if (avail) {
                                    key_unpack(kf, key, k, i, j);
                                    meta_row(p, mrows, pe, isA ? 2 : 5, k, i, j, n, rec[((size_t)k * H + i) * W + j]);
                                }
                            }
                            if (fired) {
                                key_unpack(kf, key, k, i, j);
                                const bool has = offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj);
                                if (isA) {
                                    uint64_t q = pe + 1;
                                    uint32_t cbits = __brev(__funnelshift_l(ssw[((so + 1) >> 5) + 1], ssw[(so + 1) >> 5], so + 1));
                                    bool cut = false;
                                    if (has) {
#pragma unroll
                                        for (int c4 = 0; c4 < 4; ++c4) {
                                            if (cut) break;
                                            if (q >= limit) { cut = true; break; }
                                            const uint32_t sg = cbits & 1u;
                                            if constexpr (META) {
                                                const uint32_t y = ci + (c4 >> 1), x = cj + (c4 & 1);
                                                const int32_t cur = rec[((size_t)k * H + y) * W + x];
                                                meta_row(p, mrows, q, 3, k, y, x, n, cur);
                                                if (sg && q + 1 < limit) meta_row(p, mrows, q + 1, 4, k, y, x, n, cur);
                                            }
                                            cbits >>= 1;
                                            ++q;
                                            if (sg) {
                                                if (q >= data_limit) { cut = true; break; }   // the sign must be a real bit
                                                sigmask |= 1u << c4;
                                                sgnmask |= (cbits & 1u) << c4;
                                                cbits >>= 1;
                                                ++q;
                                                ++nlsp;
                                            } else {
                                                ++nlip;
                                            }
                                            ++nread;
                                        }
                                    }
                                    if (!cut && has_desc_past_offspring(i, j, H, W)) nnext = 1;
                                    // lazy zero fill: a record with a child in the finest detail bands (the field of the
                                    // ordered-write count is free: odd LL sizes never run lazily)
#ifndef SPIHTB_NO_LAZY
                                    if (lazy_live && nread && (ci >= (uint32_t)p.fine_h0 || cj >= (uint32_t)p.fine_w0)) ndef = 1;
#endif
                                    if (has_dups && nlsp) {
                                        for (uint32_t c4 = 0; c4 < nread; ++c4)
                                            if ((sigmask & (1u << c4)) &&
                                                in_dup_subtree(ci + (c4 >> 1), cj + (c4 & 1), ll_h, ll_w)) {
                                                defmask |= 1u << c4;
                                                ++ndef;
                                            }
                                    }
                                } else {
                                    nnext = has ? 4 : 0;
                                }
                            }
                            const uint64_t pack = (uint64_t)nlsp | ((uint64_t)nlip << 12) | ((uint64_t)nnext << 24) |
                                                  ((uint64_t)ndef << 36) | ((uint64_t)(avail && !fired) << 48);
                            uint64_t tot;
                            const uint64_t ex = dec_exscan<PIPE>(pack, s_wtot, tot, idle2);
#ifndef SPIHTB_NO_LAZY
                            if (lazy_live && ((tot >> 36) & 0xfff)) {
                                // first coefficient of the image in the finest bands: zero them, all threads at work
                                // here, then go on (the flag is only read between scans: the barriers inside them order it)
                                dec_zero_fine(rec, C, H, W, (uint32_t)p.fine_h0, (uint32_t)p.fine_w0, tid, NT);
                                if (tid == 0) s_fine_done = 1;
                                dec_sync<PIPE>();
                            }
#endif
                            if (avail && !fired) R[rkeep + (uint32_t)(ex >> 48)] = key;
                            if (fired) {
                                uint32_t os = lsp_len + (uint32_t)(ex & 0xfff);
                                uint32_t oi = lip_len + (uint32_t)((ex >> 12) & 0xfff);
                                const uint32_t on = nxt_len + (uint32_t)((ex >> 24) & 0xfff);
                                uint32_t od = (uint32_t)((ex >> 36) & 0xfff);
                                if (isA) {
                                    for (uint32_t c4 = 0; c4 < nread; ++c4) {
                                        const uint32_t y = ci + (c4 >> 1), x = cj + (c4 & 1);
                                        const uint32_t ck = key_pack(kf, k, y, x);
                                        if (sigmask & (1u << c4)) {
                                            const int32_t val = (sgnmask & (1u << c4)) ? basev : -basev;
                                            if (defmask & (1u << c4))
                                                s_dq[od++] = make_uint2(ck, (uint32_t)val);
                                            else
                                                rec[((size_t)k * H + y) * W + x] = val;
                                            lsp[os++] = ck;
                                        } else {
                                            lip[oi++] = ck;
                                        }
                                    }
                                    if (nnext) nxt[on] = key & 0x7fffffffu;
                                } else if (nnext) {
#pragma unroll
                                    for (int c4 = 0; c4 < 4; ++c4)
                                        nxt[on + c4] = 0x80000000u | key_pack(kf, k, ci + (c4 >> 1), cj + (c4 & 1));
                                }
                            }
                            lsp_len += (uint32_t)(tot & 0xfff);
                            lip_len += (uint32_t)((tot >> 12) & 0xfff);
                            nxt_len += (uint32_t)((tot >> 24) & 0xfff);
                            rkeep += (uint32_t)(tot >> 48);
                            if (has_dups) {
                                dec_sync<PIPE>();
                                if (tid == 0) apply_queue(s_dq, (uint32_t)((tot >> 36) & 0xfff), kf, rec, H, W, n);
                                dec_sync<PIPE>();
                            }
                        }
                        dec_sync<PIPE>();  // s_x / s_tmask / s_grp are rewritten by the next round
                        DEC_PROF_SINCE(3, _tb);
                        DEC_PROF_CNT(12, (cnt + NT - 1) / NT);
#ifdef SPIHTB_PROF
                        // how much of the parallel work could run under the next round's walk: rounds that are not
                        // the last of their generation (slot 13: their batch cycles; slot 14: min(batch cycles of
                        // round r-1, walk cycles of round r) summed; slot 15: generations)
                        {
                            const unsigned long long tb_c = (unsigned long long)(clock64() - _tb);
                            if (tid == 0 && b == 0) {
                                if (ebase > 0) g_dec_prof[14] += prev_tb < last_walk ? prev_tb : last_walk;
                                if (ebase + DEC_CH < cur_len) g_dec_prof[13] += tb_c;
                                if (ebase == 0) g_dec_prof[15] += 1;
                            }
                            prev_tb = tb_c;
                        }
#endif
                        if (PIPE && tid == 0) {
                            s_cnt[0] = lsp_len;
                            s_cnt[1] = lip_len;
                            s_cnt[2] = nxt_len;
                            s_cnt[3] = rkeep;
                        }
                    };
                    bool pend = false;   // a walked round whose entries are not applied yet
                    uint32_t pend_ebase = 0, pend_cnt = 0;
                    uint64_t pend_pbase = 0;
                    int buf = 0;
                    for (uint32_t ebase = 0; ebase < cur_len && !ended; ebase += DEC_CH, buf ^= 1) {
                        const uint32_t cnt = min((uint32_t)DEC_CH, cur_len - ebase);
                        stage_round(ebase, cnt, buf);
                        if constexpr (!META) {
                            if (pipe) {
                                if (tid < 32)
                                    walk_round(cnt, buf);
                                else if (pend && dec_pipe_worker2(wid, idle2))
                                    apply_round(std::true_type{}, pend_ebase, pend_cnt, pend_pbase, buf ^ 1);
                                __syncthreads();
                                if (pend) {
                                    lsp_len = s_cnt[0];
                                    lip_len = s_cnt[1];
                                    nxt_len = s_cnt[2];
                                    rkeep = s_cnt[3];
                                }
                                pend = true;
                                pend_ebase = ebase;
                                pend_cnt = cnt;
                                pend_pbase = pos;
                                pos = s_chain_p;
                                if (pos >= limit) ended = true;
                                continue;
                            }
                        }
                        walk_round(cnt, buf);
                        __syncthreads();
                        const uint64_t p_base = pos;
                        pos = s_chain_p;
                        apply_round(std::false_type{}, ebase, cnt, p_base, buf);
                        if (pos >= limit) ended = true;
                    }
                    if constexpr (!META) {
                        if (pend) {   // the last walked round of the generation
                            if (dec_pipe_worker2(wid, idle2)) apply_round(std::true_type{}, pend_ebase, pend_cnt, pend_pbase, buf ^ 1);
                            __syncthreads();
                            lsp_len = s_cnt[0];
                            lip_len = s_cnt[1];
                            nxt_len = s_cnt[2];
                            rkeep = s_cnt[3];
                        }
                    }
                    uint32_t *old = cur;
                    cur = nxt;
                    nxt = gen == 0 ? G1 : old;
                    cur_len = nxt_len;
                    ++gen;
                }
                r_len = rkeep;
            }
            if (ended) break;

            // ================= refinement (encoder_decoder.rs:439-444)
            DEC_PROF_T0();
            if (!has_dups) {
                for (uint32_t e = tid; e < lsp_len0; e += DEC_NT) {
                    const uint64_t q = pos + e;
                    if (q < limit) {
                        uint32_t k, i, j;
                        key_unpack(kf, lsp[e], k, i, j);
                        int32_t *cell = rec + ((size_t)k * H + i) * W + j;
                        if constexpr (META) meta_row(p, mrows, q, 6, k, i, j, n, *cell);
                        if (q < data_limit) refine_cell(cell, n, br.bit(q));
                    }
                }
            } else {
                for (uint32_t eb = 0; eb < lsp_len0; eb += DEC_NT) {
                    const uint32_t e = eb + tid;
                    const uint64_t q = pos + e;
                    bool defer = false;
                    uint32_t key = 0, bit = 0;
                    if (e < lsp_len0 && q < limit) {
                        key = lsp[e];
                        bit = br.bit(q);
                        uint32_t k, i, j;
                        key_unpack(kf, key, k, i, j);
                        defer = in_dup_subtree(i, j, ll_h, ll_w);
                        if (!defer) refine_cell(rec + ((size_t)k * H + i) * W + j, n, bit);
                    }
                    uint64_t tot;
                    const uint64_t ex = block_exscan<DEC_NT>(defer ? 1ull : 0ull, s_wtot, tot);
                    if (defer) s_dq[(uint32_t)ex] = make_uint2(key | 0x80000000u, bit);
                    __syncthreads();
                    if (tid == 0) apply_queue(s_dq, (uint32_t)tot, kf, rec, H, W, n);
                    __syncthreads();
                }
            }
            pos += lsp_len0;
            __syncthreads();
            DEC_PROF_ADD(4);
            if (n == 0) break;
        }
        __syncthreads();
        // every coefficient that became non-zero is in the LSP: mark the 64x64 blocks of the array they lie in
        // (benign race: every writer stores 1)
        if (p.blk) {
            uint8_t *blkb = p.blk + (size_t)b * C * p.BH * p.BW;
            constexpr int MU = 8;  // keys in flight per thread (the loop is pure load latency otherwise)
            int late = 0;          // a coefficient on the first row / column of the finest bands (inside the zeroed corner)
            for (uint32_t e0 = tid; e0 < lsp_len; e0 += DEC_NT * MU) {
                uint32_t key[MU];
#pragma unroll
                for (int u = 0; u < MU; ++u) key[u] = e0 + u * DEC_NT < lsp_len ? lsp[e0 + u * DEC_NT] : 0xffffffffu;
#pragma unroll
                for (int u = 0; u < MU; ++u) {
                    if (e0 + u * DEC_NT < lsp_len) {
                        uint32_t k, i, j;
                        key_unpack(kf, key[u], k, i, j);
                        blkb[((size_t)k * p.BH + (i >> 6)) * p.BW + (j >> 6)] = 1;
                        if (p.blk1 && (i >= (uint32_t)p.fine_ht || j >= (uint32_t)p.fine_wt)) {
                            p.blk1[(((size_t)b * C + k) * p.BH + (i >> 6)) * p.BW + (j >> 6)] = 1;
                            if (p.l1_any) p.l1_any[b] = 1;
                            late = 1;
                        }
                    }
                }
            }
            // lazy zero fill: the image never read a child past the zeroed corner, but a coefficient on the first row or
            // column of the finest bands (inside the corner) is non-zero, so the level-1 inverse will read those bands:
            // zero them now (nothing was ever written there)
            if (__syncthreads_or(late) && !s_fine_done)
                dec_zero_fine(rec, C, H, W, (uint32_t)p.fine_h0, (uint32_t)p.fine_w0, tid, DEC_NT);
        }
        DEC_PROF_SINCE(5, _timg);
#ifdef SPIHTB_PROF
        if (tid == 0 && b < 1024) {
            uint32_t smid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            g_dec_img[b][0] = (unsigned long long)(clock64() - _timg);
            g_dec_img[b][1] = walk_cycles;
            g_dec_img[b][2] = smid;
            uint32_t hwid;
            asm("mov.u32 %0, %%warpid;" : "=r"(hwid));
            g_dec_img[b][3] = hwid;
        }
#endif
    }
}

#ifdef SPIHTB_PROF
// debug: read and clear the phase counters (tools/dec_phases.py)
extern "C" int spihtb_debug_dec_img(unsigned long long *out, int n)
{
    if (n > 1024) n = 1024;
    if (cudaMemcpyFromSymbol(out, g_dec_img, sizeof(unsigned long long) * 4 * n) != cudaSuccess) return SPIHTB_ECUDA;
    return SPIHTB_OK;
}
extern "C" int spihtb_debug_dec_prof(unsigned long long *out16)
{
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out16, g_dec_prof, sizeof(z)) != cudaSuccess) return SPIHTB_ECUDA;
    if (cudaMemcpyToSymbol(g_dec_prof, z, sizeof(z)) != cudaSuccess) return SPIHTB_ECUDA;
    return SPIHTB_OK;
}
#endif

// zero the top-left h0 x w0 corner of every plane (lazy zero fill: everything but the finest detail bands).  A CTA
// takes every gridDim.x-th group of eight rows of one plane (one warp per row); planes in grid.y (strided when there are
// more than 65535).
__global__ void __launch_bounds__(256) zero_corner_kernel(int32_t *out, int nplanes, int H, int W, int h0, int w0)
{
    const int lane = threadIdx.x & 31;
    for (int z = blockIdx.y; z < nplanes; z += gridDim.y) {
        for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < h0; row += gridDim.x * 8) {
            int32_t *r = out + ((size_t)z * H + row) * W;
            // 16-byte stores over the aligned middle of the row segment, scalar stores at its ends
            const int head = (int)(((16u - (uint32_t)((uintptr_t)r & 15u)) & 15u) >> 2);
            const int a0 = min(head, w0), nq = (w0 - a0) >> 2, a1 = a0 + 4 * nq;
            if (lane < a0) r[lane] = 0;
            int4 *q = reinterpret_cast<int4 *>(r + a0);
            for (int t = lane; t < nq; t += 32) q[t] = make_int4(0, 0, 0, 0);
            if (a1 + lane < w0) r[a1 + lane] = 0;
        }
    }
}

int launch_decode(spihtb_ctx *ctx, const DecArgs &a)
{
    DecK k;
    k.meta = a.meta;
    k.meta_rows = a.meta_rows;
    k.level = a.level;
    k.top_ei = a.top_ei;
    k.top_ej = a.top_ej;
    k.slices = a.slices;
    k.meta_err = a.meta_err;
    if (a.meta && ((a.ll_h | a.ll_w) & 1)) {
        set_error("decode_with_metadata needs even LL sizes (got %dx%d): with an odd LL band cells have two parents "
                  "and the per-coefficient depth / filter of the reference's lists cannot be told from the position",
                  a.ll_h, a.ll_w);
        return SPIHTB_EGEOM;
    }
    if (!make_keyfmt(a.C, a.H, a.W, &k.kf)) {
        set_error("shape c=%d h=%d w=%d does not fit a 31-bit packed list entry", a.C, a.H, a.W);
        return SPIHTB_ESHAPE;
    }
    if ((a.in_stride & 7) != 0 || a.in_stride == 0) {
        set_error("in_stride must be a positive multiple of 8 bytes");
        return SPIHTB_EINVAL;
    }
    k.in = reinterpret_cast<const uint32_t *>(a.in);
    k.in_stride_words = a.in_stride / 4;
    k.nbytes = a.nbytes;
    k.n = a.n;
    k.B = a.B; k.C = a.C; k.H = a.H; k.W = a.W; k.ll_h = a.ll_h; k.ll_w = a.ll_w;
    k.out = a.out;
    k.blk = a.blk;
    k.BH = (a.H + 63) / 64;
    k.BW = (a.W + 63) / 64;

    k.walk_variant = 1;   // SPIHTB_WALK=0: the branchy loop (A-B measurements)
    if (const char *e = getenv("SPIHTB_WALK")) k.walk_variant = atoi(e);
    k.pipe = 1;
    if (const char *e = getenv("SPIHTB_DEC_PIPE")) k.pipe = atoi(e);
    int occ = 1;
    if (a.meta)
        SPIHTB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spiht_decode_kernel<true>, DEC_NT, 0));
    else
        SPIHTB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spiht_decode_kernel<false>, DEC_NT, 0));
    if (occ < 1) occ = 1;
    const int slots = std::min(a.B, ctx->sm_count * occ);

    const uint64_t T0 = (uint64_t)a.ll_h * a.ll_w * a.C;
    const uint64_t chw = (uint64_t)a.C * a.H * a.W;
    const uint64_t budget = a.in_stride * 8;
    const uint64_t pix_cap = std::min<uint64_t>(chw + T0, T0 + budget) + DEC_SLACK;
    const uint64_t lis_shape = (uint64_t)a.C * (a.H / 2 + 2) * (a.W / 2 + 2) * 5 / 4 + T0;
    const uint64_t lis_cap = std::min<uint64_t>(lis_shape, T0 + budget) + DEC_SLACK;
    k.pix_cap = pix_cap;
    k.lis_cap = lis_cap;
    const size_t per_slot = (pix_cap * 3 + lis_cap * 3) * sizeof(uint32_t);
    int rc = ctx->ensure(ctx->lists, per_slot * slots + 256);
    if (rc) return rc;
    rc = ctx->ensure(ctx->misc, 256);
    if (rc) return rc;
    uint32_t *base = static_cast<uint32_t *>(ctx->lists.p);
    k.lis = base;
    k.lip = base + (size_t)slots * lis_cap * 3;
    k.lsp = k.lip + (size_t)slots * pix_cap * 2;
    k.counter = static_cast<unsigned int *>(ctx->misc.p);

    ctx->stage_begin(5);
    // (zeroing each image's array inside the kernel, under the first passes, was measured slower than this memset:
    // 5.56 against 5.41 ms per 256 images -- one CTA cannot issue 13 MB of stores as fast as the copy engine path)
    k.fine_h0 = k.fine_w0 = 0;
    k.fine_ht = a.fine_h0;
    k.fine_wt = a.fine_w0;
    k.blk1 = a.blk1;
    k.l1_any = a.l1_any;
    if (a.lazy_zero && a.blk1 && a.fine_h0 > 0 && a.fine_w0 > 0 && a.fine_h0 < a.H && a.fine_w0 < a.W && !a.meta &&
        !((a.ll_h | a.ll_w) & 1)) {
        // lazy: only the corner that holds everything but the finest detail bands; an image that reaches those bands
        // zeroes them itself (spiht_decode_kernel)
        k.fine_h0 = std::min(a.H, (a.fine_h0 + 1) & ~1);
        k.fine_w0 = std::min(a.W, (a.fine_w0 + 1) & ~1);
        // (cudaMemset3DAsync over 768 planes of 539-column rows took 3 ms; this kernel runs at memory speed)
        {
            const dim3 grid((unsigned)std::max(1, std::min(8, (k.fine_h0 + 63) / 64)), (unsigned)std::min<long long>((long long)a.B * a.C, 65535));
            zero_corner_kernel<<<grid, 256, 0, ctx->stream>>>(a.out, a.B * a.C, a.H, a.W, k.fine_h0, k.fine_w0);
            ctx->launches++;
        }
    } else {
        SPIHTB_CUDA_CHECK(cudaMemsetAsync(a.out, 0, sizeof(int32_t) * (size_t)a.B * a.C * a.H * a.W, ctx->stream));
    }
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(k.counter, 0, sizeof(unsigned int), ctx->stream));
    if (a.meta)
        spiht_decode_kernel<true><<<slots, DEC_NT, 0, ctx->stream>>>(k);
    else
        spiht_decode_kernel<false><<<slots, DEC_NT, 0, ctx->stream>>>(k);
    ctx->launches++;
    ctx->stage_end(5);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
