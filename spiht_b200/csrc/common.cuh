// Shared device/host helpers for libspiht_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/spiht_b200.h"

namespace spihtb {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
#define SPIHTB_CUDA_CHECK(expr)                                                        \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            spihtb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                     \
            return SPIHTB_ECUDA;                                                       \
        }                                                                              \
    } while (0)

// ---------------------------------------------------------------- context
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace spihtb

namespace spihtb {
constexpr int PROF_RING = 64;
struct StageProf {
    cudaEvent_t a[PROF_RING], b[PROF_RING];
    bool made = false;
    int64_t recorded = 0, harvested = 0;
    double acc_ms = 0.0;
};
}  // namespace spihtb

struct spihtb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    // side stream for work that is independent of the main chain of a call (the gap fill of the forward
    // transform runs beside the first DWT level); fork / join events
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_order = nullptr;  // orders the shared workspaces across a change of stream (spihtb_set_stream)
    // Group pipeline of spihtb_encode_images: the batch is cut into groups of images and group g runs on
    // sub-context g % nsubs -- its own streams and workspaces -- so that the (latency-bound) coder of one group
    // overlaps the (bandwidth-bound) transform of the next.  A sub-context runs the transform and the pyramid on
    // `stream` and the coder on `coder_stream` (higher priority: its few CTAs are placed as soon as an SM has room).
    static constexpr int MAX_SUBS = 4;
    spihtb_ctx *subs[MAX_SUBS] = {nullptr, nullptr, nullptr, nullptr};
    int nsubs = 0;
    bool is_sub = false;
    cudaStream_t coder_stream = nullptr;                 // sub-contexts only
    cudaEvent_t ev_pyr = nullptr, ev_done = nullptr;     // sub-contexts: pyramid ready / group finished
    cudaEvent_t ev_in = nullptr;                         // parent: inputs ready on the caller's stream
    int64_t launches = 0;
    // grow-only device workspaces
    spihtb::DevBuf pyr;      // DP / LP planes + LL-root planes + per-image max
    spihtb::DevBuf lists;    // coder list storage, one slice per resident CTA
    spihtb::DevBuf misc;     // counters, small per-image arrays
    spihtb::DevBuf tmpa, tmpb;  // DWT approximation ping-pong (float64)
    spihtb::DevBuf tail;        // forward tail kernel: two private approximation planes per (image, channel) plane
    spihtb::DevBuf io;       // staging for the host-pointer entry points
    spihtb::DevBuf io2;
    spihtb::DevBuf u8lut;    // k / 255.0 for uint8 pixels
    spihtb::DevBuf blk;      // marks of the 64x64 blocks the decoder wrote into (spihtb_decode_images)
    spihtb::DevBuf fix;      // rectangles of the pyramid fix-up pass (forward transform with fused base pass)
    std::vector<int32_t> fix_host;  // their host copy: [key (32 ints)] [nrect, total] [rects] [prefix]
    std::vector<uint8_t> host_out;  // spihtb_encode result
    bool profiling = false;
    bool scratch_coeffs = false;  // SPIHTB_OPT_SCRATCH_COEFFS
    bool last_forward_fused12 = false;  // the last forward transform ran levels 1+2 in the fused TMA kernel
    spihtb::StageProf prof[SPIHTB_NSTAGES];
    int ensure(spihtb::DevBuf &b, size_t bytes);
    void stage_begin(int s);
    void stage_end(int s);
    void harvest(int s, bool all);
};

namespace spihtb {

// ---------------------------------------------------------------- list entries
// A list entry packs a coefficient coordinate into 31 bits:
//   [ j : bj ][ i : bi ][ k : bk ]   bit 31 = set type (1 = A / D-set, 0 = B / L-set)
struct KeyFmt {
    int sj;  // shift of i  (= bits of j)
    int sk;  // shift of k  (= bits of j + bits of i)
    uint32_t mj, mi;
};

static inline int bits_for(uint32_t n)  // bits needed for values 0..n-1
{
    int b = 0;
    while ((1ull << b) < n) ++b;
    return b;
}

static inline bool make_keyfmt(int c, int h, int w, KeyFmt *f)
{
    int bj = bits_for((uint32_t)w), bi = bits_for((uint32_t)h), bk = bits_for((uint32_t)c);
    if (bj + bi + bk > 31) return false;
    f->sj = bj;
    f->sk = bj + bi;
    f->mj = (1u << bj) - 1;
    f->mi = (1u << bi) - 1;
    return true;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t key_pack(const KeyFmt &f, uint32_t k, uint32_t i, uint32_t j)
{
    return j | (i << f.sj) | (k << f.sk);
}
__device__ __forceinline__ void key_unpack(const KeyFmt &f, uint32_t key, uint32_t &k, uint32_t &i, uint32_t &j)
{
    j = key & f.mj;
    i = (key >> f.sj) & f.mi;
    k = (key & 0x7fffffffu) >> f.sk;
}

__device__ __forceinline__ uint32_t absu(int32_t x) { return x < 0 ? (uint32_t)(-x) : (uint32_t)x; }
// 1 + floor(log2 v), 0 for v == 0
__device__ __forceinline__ uint32_t plane1(uint32_t v) { return 32u - (uint32_t)__clz((int)v); }

// (max as f32).log2() as u8  (encoder_decoder.rs:166).  The f32 log2 is taken
// as the correctly rounded one: computed in double, rounded to float once.
__device__ __forceinline__ int max_n_of(uint32_t max_abs)
{
    if (max_abs == 0) return 0;
    float f = __uint2float_rn(max_abs);
    float l = __double2float_rn(log2((double)f));
    if (!(l > 0.f)) return 0;
    return (int)l;
}

// Offspring of node (i,j) (encoder_decoder.rs:43-75).  Returns false when the
// node has none; otherwise (ci,cj) is the top-left corner of its 2x2 block.
__device__ __forceinline__ bool offspring_corner(uint32_t i, uint32_t j, uint32_t h, uint32_t w,
                                                 uint32_t ll_h, uint32_t ll_w, uint32_t &ci, uint32_t &cj)
{
    if (i < ll_h && j < ll_w) {
        if (((i | j) & 1u) == 0) return false;
        ci = (i & 1u) * ll_h + (i & ~1u);
        cj = (j & 1u) * ll_w + (j & ~1u);
        return true;
    }
    if (2 * i + 1 >= h || 2 * j + 1 >= w) return false;
    ci = 2 * i;
    cj = 2 * j;
    return true;
}
// encoder_decoder.rs:7-12
__device__ __forceinline__ bool has_desc_past_offspring(uint32_t i, uint32_t j, uint32_t h, uint32_t w)
{
    return !((i * 2 + 1) * 2 + 1 >= h || (j * 2 + 1) * 2 + 1 >= w);
}

// ---------------------------------------------------------------- block scan
// Exclusive scan of a 64-bit packed counter vector over the CTA.  `warp_tot`
// is shared scratch of NT/32 entries.  Returns the exclusive prefix and the
// CTA total.  Two __syncthreads(); safe to call back to back.
template <int NT>
__device__ __forceinline__ uint64_t block_exscan(uint64_t v, uint64_t *warp_tot, uint64_t &total)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();  // previous users of warp_tot are done
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint64_t base = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        uint64_t t = warp_tot[q];
        if (q < wid) base += t;
        tot += t;
    }
    total = tot;
    return base + inc - v;
}
// Same scan with two barriers and far fewer instructions per thread: warp 0 scans the warp totals.
// `buf` is shared scratch of 2 x (NT/32 + 1) entries, `parity` a per-thread (CTA-uniform) toggle that
// alternates the two halves so that back-to-back calls need no extra barrier.
template <int NT, typename T>
__device__ __forceinline__ T block_exscan2(T v, T (*buf)[NT / 32 + 1], int &parity, T &total)
{
    constexpr int NW = NT / 32;
    static_assert(NW <= 32, "one warp scans the warp totals");
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T *b = buf[parity];
    parity ^= 1;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) b[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const T t = lane < NW ? b[lane] : (T)0;
        T ti = t;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            T u = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += u;
        }
        if (lane < NW) b[lane] = ti - t;
        if (lane == NW - 1) b[NW] = ti;
    }
    __syncthreads();
    total = b[NW];
    return b[wid] + inc - v;
}
#endif  // __CUDACC__

}  // namespace spihtb
