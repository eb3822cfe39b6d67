// Descendant-max pyramid: for every tree node the bit-plane at which its D-set
// (all descendants) and L-set (grand-descendants and below) first become
// significant.  Replaces the recursive scans is_set_sig / is_l_sig of the
// reference (src/encoder_decoder.rs:78-121) with one streaming pass over the
// coefficient array plus a few tiny passes over the quarter-size node grid.
//
//   dp[z][i][j] = 1 + floor(log2 max|x| over all descendants of (i,j))   (0: none / all zero)
//   lp[z][i][j] = same over grand-descendants and below
// for the dyadic rule (encoder_decoder.rs:65-74) on the node grid
// [0,h/2) x [0,w/2); dpll/lpll hold the same for the LL roots, whose offspring
// follow the block rule of encoder_decoder.rs:44-62.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace spihtb {

// ---- base pass: reads every coefficient once (HBM-bound) ------------------
// blockDim (32, 8); a warp owns one row pair.  A lane reads the two adjacent
// columns of its cell in both rows (the 2x2 block of one tree node), so a cell
// needs no exchange and a warp writes 32 consecutive plane bytes.
constexpr int PYR_UNROLL = 4;

__global__ void __launch_bounds__(256) pyr_base_kernel(const int32_t *__restrict__ coeffs, int H, int W, int NH, int NW,
                                                       int gy, int C, uint8_t *__restrict__ dp,
                                                       uint8_t *__restrict__ lp, uint32_t *__restrict__ maxabs)
{
    __shared__ uint32_t s_max[8];
    const int lane = threadIdx.x, wy = threadIdx.y;
    const int by = blockIdx.x % gy;
    const int z = blockIdx.x / gy;  // image * C + channel

    const int ip = by * 8 + wy;  // row pair
    const int r0 = 2 * ip;
    uint32_t wmax = 0;
    if (r0 < H) {
        const int32_t *p0 = coeffs + (size_t)z * H * W + (size_t)r0 * W;
        const bool has1 = r0 + 1 < H;
        const int32_t *p1 = has1 ? p0 + W : p0;
        uint8_t *drow = dp + ((size_t)z * NH + (ip < NH ? ip : 0)) * NW;
        uint8_t *lrow = lp + ((size_t)z * NH + (ip < NH ? ip : 0)) * NW;  // cleared: ring-1 nodes have no L-set
        const int ncell = (W + 1) >> 1;  // cells incl. a half cell in the last odd column
        for (int jb = 0; jb < ncell; jb += 32 * PYR_UNROLL) {
            uint32_t m[PYR_UNROLL];
#pragma unroll
            for (int u = 0; u < PYR_UNROLL; ++u) {
                const int j = jb + u * 32 + lane;
                const int c = 2 * j;
                uint32_t v = 0;
                if (c < W) {
                    v = max(absu(__ldg(p0 + c)), absu(__ldg(p1 + c)));
                    if (c + 1 < W) v = max(v, max(absu(__ldg(p0 + c + 1)), absu(__ldg(p1 + c + 1))));
                }
                m[u] = v;
            }
#pragma unroll
            for (int u = 0; u < PYR_UNROLL; ++u) {
                const int j = jb + u * 32 + lane;
                wmax = max(wmax, m[u]);
                if (ip < NH && j < NW) {
                    drow[j] = (uint8_t)plane1(m[u]);
                    lrow[j] = 0;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
    if (lane == 0) s_max[wy] = wmax;
    __syncthreads();
    if (wy == 0 && lane == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t = max(t, s_max[q]);
        if (t) atomicMax(maxabs + z / C, t);
    }
}

// ---- ring t >= 2: nodes whose deepest child chain has length t ------------
// node (i,j) of ring [0,RH) x [0,RW) minus the deeper rings [0,IH) x [0,IW):
//   lp = max dp over its four child nodes, dp = max(dp, lp)
// 8 / 4 bytes at any alignment from aligned 32-bit words (reads up to 3 bytes past the range: the planes
// are followed by other planes of the same workspace)
__device__ __forceinline__ uint2 ld8_any(const uint8_t *p)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const int sh = 8 * (int)(reinterpret_cast<uintptr_t>(p) & 3);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
// per-byte maximum of two words whose bytes are all below 128 (plane numbers are at most 32): the borrow-free
// subtraction (a | 0x80..) - b leaves bit 7 of a byte set where a >= b.  (The __vmaxu4 intrinsic is emulated
// on this architecture and several times longer.)
__device__ __forceinline__ uint32_t bmax4(uint32_t a, uint32_t b)
{
    const uint32_t ge = (((a | 0x80808080u) - b) >> 7) & 0x01010101u;
    const uint32_t m = ge * 0xffu;  // 0xff in the bytes where a >= b
    return (a & m) | (b & ~m);
}
__device__ __forceinline__ uint32_t ld4_any(const uint8_t *p)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const int sh = 8 * (int)(reinterpret_cast<uintptr_t>(p) & 3);
    return __funnelshift_r(w[0], w[1], sh);
}

// One quad = four consecutive nodes (i, j0 .. j0+3) of a ring in one plane.  quad_load gathers their eight child
// bytes in each of two rows and their own four bytes; quad_store writes lp = max over the children and
// dp = max(dp, lp) for the nodes of the quad that belong to the ring (`member`).
struct Quad {
    uint2 a, b;
    uint32_t own;
};
__device__ __forceinline__ uint32_t quad_members(int i, int j0, int RW, int IH, int IW)
{
    uint32_t member = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (j0 + q < RW && !(i < IH && j0 + q < IW)) member |= 1u << q;  // outside the deeper rings (and (0,0))
    return member;
}
__device__ __forceinline__ Quad quad_load(const uint8_t *d, int NH, int NW, int i, int j0)
{
    Quad q;
    const size_t c00 = (size_t)(2 * i) * NW + 2 * j0;
    q.a = ld8_any(d + c00);
    q.b = 2 * i + 1 < NH ? ld8_any(d + c00 + NW) : make_uint2(0u, 0u);  // 2i < NH holds for every ring node
    q.own = ld4_any(d + (size_t)i * NW + j0);
    return q;
}
__device__ __forceinline__ void quad_store(const Quad &q, uint8_t *d, uint8_t *l, int NW, int i, int j0, uint32_t member)
{
    // child columns 2 j0 .. 2 j0 + 7 that exist
    const int ncol = min(8, NW - 2 * j0);
    const uint32_t mlo = ncol >= 4 ? 0xffffffffu : (0xffffffffu >> (8 * (4 - ncol)));
    const uint32_t mhi = ncol >= 8 ? 0xffffffffu : (ncol <= 4 ? 0u : (0xffffffffu >> (8 * (8 - ncol))));
    const uint32_t lo = bmax4(q.a.x, q.b.x) & mlo, hi = bmax4(q.a.y, q.b.y) & mhi;
    const uint32_t tl = bmax4(lo, lo >> 8), th = bmax4(hi, hi >> 8);  // bytes 0, 2: node maxima
    const uint32_t m4 = __byte_perm(tl, th, 0x6420);
    const uint32_t d4 = bmax4(q.own, m4);
    uint8_t *lq = l + (size_t)i * NW + j0, *dq = d + (size_t)i * NW + j0;
    if (member == 0xfu && (reinterpret_cast<uintptr_t>(lq) & 3) == 0) {
        *reinterpret_cast<uint32_t *>(lq) = m4;
        *reinterpret_cast<uint32_t *>(dq) = d4;
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (member & (1u << t)) {
                lq[t] = (uint8_t)(m4 >> (8 * t));
                dq[t] = (uint8_t)(d4 >> (8 * t));
            }
        }
    }
}

// one large ring (ring 2 is a quarter of the node grid): blockDim (32, 8); a thread takes four consecutive
// nodes of a row -- their children are 8 consecutive bytes in each of two rows -- in PYR_ZPT planes with all
// loads in flight together; byte-wise maxima by SWAR arithmetic (bmax4)
constexpr int PYR_ZPT = 4;
__global__ void __launch_bounds__(256) pyr_ring_kernel(int NH, int NW, int RH, int RW, int IH, int IW, int nz,
                                                       uint8_t *__restrict__ dp, uint8_t *__restrict__ lp)
{
    const int j0 = (blockIdx.x * 32 + threadIdx.x) * 4, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= RH || j0 >= RW) return;
    const uint32_t member = quad_members(i, j0, RW, IH, IW);
    if (!member) return;
    const size_t plane = (size_t)NH * NW;
    for (int z0 = blockIdx.z * PYR_ZPT; z0 < nz; z0 += gridDim.z * PYR_ZPT) {
        Quad q[PYR_ZPT];
#pragma unroll
        for (int u = 0; u < PYR_ZPT; ++u)
            if (z0 + u < nz) q[u] = quad_load(dp + (size_t)(z0 + u) * plane, NH, NW, i, j0);
#pragma unroll
        for (int u = 0; u < PYR_ZPT; ++u)
            if (z0 + u < nz)
                quad_store(q[u], dp + (size_t)(z0 + u) * plane, lp + (size_t)(z0 + u) * plane, NW, i, j0, member);
    }
}

// the small rings t_start.. and the LL roots of one plane per CTA (each ring is a quarter of the one
// before it: the rings of a plane are chained with barriers instead of one launch per ring)
__global__ void __launch_bounds__(256) pyr_rest_kernel(const int32_t *__restrict__ coeffs, int H, int W, int NH, int NW,
                                                       int ll_h, int ll_w, int t_start, uint8_t *__restrict__ dp,
                                                       uint8_t *__restrict__ lp, uint8_t *__restrict__ dpll,
                                                       uint8_t *__restrict__ lpll)
{
    const int z = blockIdx.x;
    uint8_t *d = dp + (size_t)z * NH * NW;
    uint8_t *l = lp + (size_t)z * NH * NW;
    for (int t = t_start;; ++t) {
        const long long s = 1LL << (t - 1);
        const int RH = (int)((NH + s - 1) / s), RW = (int)((NW + s - 1) / s);
        if (RH <= 1 && RW <= 1) break;
        const int IH = (int)((NH + 2 * s - 1) / (2 * s)), IW = (int)((NW + 2 * s - 1) / (2 * s));
        // four nodes per thread and step, two steps in flight (the chain of rings is latency-bound)
        const int qpr = (RW + 3) / 4, nq = RH * qpr;
        for (int r = threadIdx.x; r < nq; r += 2 * blockDim.x) {
            Quad q[2];
            int qi[2], qj[2];
            uint32_t mem[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int rr = r + u * blockDim.x;
                qi[u] = rr / qpr;
                qj[u] = 4 * (rr - qi[u] * qpr);
                mem[u] = rr < nq ? quad_members(qi[u], qj[u], RW, IH, IW) : 0u;
                if (mem[u]) q[u] = quad_load(d, NH, NW, qi[u], qj[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (mem[u]) quad_store(q[u], d, l, NW, qi[u], qj[u], mem[u]);
        }
        __syncthreads();  // the next ring reads this ring's dp
    }
    // LL roots (encoder_decoder.rs:44-62)
    const int32_t *a = coeffs + (size_t)z * H * W;
    for (int r = threadIdx.x; r < ll_h * ll_w; r += blockDim.x) {
        const uint32_t i = r / ll_w, j = r % ll_w;
        uint32_t ci, cj, dd = 0, ll = 0;
        if (offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t y = ci + (q >> 1), x = cj + (q & 1);
                dd = max(dd, plane1(absu(a[(size_t)y * W + x])));
                if (y < (uint32_t)NH && x < (uint32_t)NW) ll = max(ll, (uint32_t)d[(size_t)y * NW + x]);
            }
            dd = max(dd, ll);
        }
        dpll[(size_t)z * ll_h * ll_w + r] = (uint8_t)dd;
        lpll[(size_t)z * ll_h * ll_w + r] = (uint8_t)ll;
    }
}

// ---- cells the fused base pass of the forward transform could not finish (dwt_fwd.cu: fix_rects) -----
struct FixK {
    const int32_t *coeffs;
    int H, W, NH, NW, nz;
    uint8_t *dp;
    const FixRect *rects;
    const uint32_t *prefix;  // [nrect + 1] running cell counts
    int nrect;
    uint32_t total;
};
__global__ void __launch_bounds__(256) pyr_fix_kernel(const FixK p)
{
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= p.total) return;
    int lo = 0, hi = p.nrect;  // last rect with prefix[r] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p.prefix + mid) <= t)
            lo = mid;
        else
            hi = mid;
    }
    const FixRect r = p.rects[lo];
    const uint32_t local = t - __ldg(p.prefix + lo), w = (uint32_t)(r.b1 - r.b0);
    const int a = r.a0 + (int)(local / w), b = r.b0 + (int)(local % w);
    for (int z = blockIdx.y; z < p.nz; z += gridDim.y) {
        const int32_t *x = p.coeffs + ((size_t)z * p.H + 2 * a) * p.W + 2 * b;  // 2a+1 < H, 2b+1 < W inside the node grid
        const uint32_t m = max(max(absu(x[0]), absu(x[1])), max(absu(x[p.W]), absu(x[p.W + 1])));
        p.dp[((size_t)z * p.NH + a) * p.NW + b] = (uint8_t)plane1(m);
    }
}

int launch_pyr_fix(spihtb_ctx *ctx, const int32_t *coeffs, int nz, int H, int W, uint8_t *dp, const FixRect *dev_rects,
                   const uint32_t *dev_prefix, int nrect, uint32_t total)
{
    if (nrect == 0 || total == 0) return SPIHTB_OK;
    FixK k;
    k.coeffs = coeffs;
    k.H = H; k.W = W; k.NH = H / 2; k.NW = W / 2; k.nz = nz;
    k.dp = dp;
    k.rects = dev_rects;
    k.prefix = dev_prefix;
    k.nrect = nrect;
    k.total = total;
    pyr_fix_kernel<<<dim3((total + 255) / 256, std::min(nz, 65535)), 256, 0, ctx->stream>>>(k);
    ctx->launches++;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

int launch_pyramid(spihtb_ctx *ctx, const int32_t *coeffs, int B, int C, int H, int W, int ll_h, int ll_w,
                   uint8_t *dp, uint8_t *lp, uint8_t *dpll, uint8_t *lpll, uint32_t *maxabs, bool base_done)
{
    cudaStream_t st = ctx->stream;
    const int NH = H / 2, NW = W / 2, nz = B * C;
    if (!base_done) {
        SPIHTB_CUDA_CHECK(cudaMemsetAsync(maxabs, 0, sizeof(uint32_t) * B, st));
        const int gy = ((H + 1) / 2 + 7) / 8;
        const long long nb = (long long)gy * nz;
        if (nb > 0x7fffffffLL) {
            set_error("pyramid grid too large");
            return SPIHTB_ESHAPE;
        }
        ctx->stage_begin(2);
        pyr_base_kernel<<<(unsigned)nb, dim3(32, 8), 0, st>>>(coeffs, H, W, NH, NW, gy, C, dp, lp, maxabs);
        ctx->launches++;
        ctx->stage_end(2);
    }
    ctx->stage_begin(3);
    {
        // large rings: one launch each over all planes; from the first ring of at most 64 Ki nodes on, the
        // remaining rings and the LL roots of a plane are chained inside one CTA
        int t = 2;
        for (;; ++t) {
            const long long s = 1LL << (t - 1);
            const int RH = (int)((NH + s - 1) / s), RW = (int)((NW + s - 1) / s);
            if (RH <= 1 && RW <= 1) break;
            if (t > 2 && (long long)RH * RW <= 65536) break;
            const int IH = (int)((NH + 2 * s - 1) / (2 * s)), IW = (int)((NW + 2 * s - 1) / (2 * s));
            const dim3 grid((RW + 127) / 128, (RH + 7) / 8, std::min((nz + PYR_ZPT - 1) / PYR_ZPT, 65535));
            pyr_ring_kernel<<<grid, dim3(32, 8), 0, st>>>(NH, NW, RH, RW, IH, IW, nz, dp, lp);
            ctx->launches++;
        }
        pyr_rest_kernel<<<(unsigned)nz, 256, 0, st>>>(coeffs, H, W, NH, NW, ll_h, ll_w, t, dp, lp, dpll, lpll);
        ctx->launches++;
    }
    ctx->stage_end(3);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

// ---- per-image maximum magnitude of a coefficient batch (spihtb_max_abs): sizes the stream rows of an
// untruncated encode.  gridDim.y = images; blocks of an image stride over its coefficients.
__global__ void __launch_bounds__(256) max_abs_kernel(const int32_t *__restrict__ coeffs, size_t per_image,
                                                      uint32_t *__restrict__ out)
{
    __shared__ uint32_t s_max[8];
    const int32_t *p = coeffs + (size_t)blockIdx.y * per_image;
    uint32_t m = 0;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < per_image; t += (size_t)gridDim.x * 256)
        m = max(m, absu(__ldg(p + t)));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t = max(t, s_max[q]);
        if (t) atomicMax(out + blockIdx.y, t);
    }
}

int launch_max_abs(spihtb_ctx *ctx, const int32_t *coeffs, int B, size_t per_image, uint32_t *out)
{
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(uint32_t) * (size_t)B, ctx->stream));
    const unsigned gx = (unsigned)std::min<size_t>((per_image + 256 * 8 - 1) / (256 * 8), 1024);
    max_abs_kernel<<<dim3(gx ? gx : 1, (unsigned)B), 256, 0, ctx->stream>>>(coeffs, per_image, out);
    ctx->launches++;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
