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
                                                       uint32_t *__restrict__ maxabs)
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
                if (ip < NH && j < NW) drow[j] = (uint8_t)plane1(m[u]);
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
__global__ void __launch_bounds__(256) pyr_up_kernel(int NH, int NW, int RH, int RW, int IH, int IW, int nz,
                                                     uint8_t *__restrict__ dp, uint8_t *__restrict__ lp)
{
    const size_t per = (size_t)RH * RW;
    const size_t total = per * nz;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(t / per);
        const int r = (int)(t % per);
        const int i = r / RW, j = r % RW;
        if (i < IH && j < IW) continue;  // deeper ring (or the self-referential node (0,0))
        uint8_t *d = dp + (size_t)z * NH * NW;
        uint32_t l = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int ci = 2 * i + (q >> 1), cj = 2 * j + (q & 1);
            if (ci < NH && cj < NW) l = max(l, (uint32_t)d[(size_t)ci * NW + cj]);
        }
        size_t o = (size_t)i * NW + j;
        lp[(size_t)z * NH * NW + o] = (uint8_t)l;
        d[o] = (uint8_t)max((uint32_t)d[o], l);
    }
}

// ---- LL roots --------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_ll_kernel(const int32_t *__restrict__ coeffs, int H, int W, int NH, int NW,
                                                     int ll_h, int ll_w, int nz, const uint8_t *__restrict__ dp,
                                                     uint8_t *__restrict__ dpll, uint8_t *__restrict__ lpll)
{
    const int per = ll_h * ll_w;
    const size_t total = (size_t)per * nz;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(t / per);
        const int r = (int)(t % per);
        const uint32_t i = r / ll_w, j = r % ll_w;
        uint32_t ci, cj, dd = 0, ll = 0;
        if (offspring_corner(i, j, H, W, ll_h, ll_w, ci, cj)) {
            const int32_t *a = coeffs + (size_t)z * H * W;
            const uint8_t *d = dp + (size_t)z * NH * NW;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t y = ci + (q >> 1), x = cj + (q & 1);
                dd = max(dd, plane1(absu(a[(size_t)y * W + x])));
                if (y < (uint32_t)NH && x < (uint32_t)NW) ll = max(ll, (uint32_t)d[(size_t)y * NW + x]);
            }
            dd = max(dd, ll);
        }
        dpll[t] = (uint8_t)dd;
        lpll[t] = (uint8_t)ll;
    }
}

int launch_pyramid(spihtb_ctx *ctx, const int32_t *coeffs, int B, int C, int H, int W, int ll_h, int ll_w,
                   uint8_t *dp, uint8_t *lp, uint8_t *dpll, uint8_t *lpll, uint32_t *maxabs)
{
    cudaStream_t st = ctx->stream;
    const int NH = H / 2, NW = W / 2, nz = B * C;
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(maxabs, 0, sizeof(uint32_t) * B, st));
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(lp, 0, (size_t)nz * NH * NW, st));  // ring-1 nodes have no L-set
    {
        const int gy = ((H + 1) / 2 + 7) / 8;
        const long long nb = (long long)gy * nz;
        if (nb > 0x7fffffffLL) {
            set_error("pyramid grid too large");
            return SPIHTB_ESHAPE;
        }
        ctx->stage_begin(2);
        pyr_base_kernel<<<(unsigned)nb, dim3(32, 8), 0, st>>>(coeffs, H, W, NH, NW, gy, C, dp, maxabs);
        ctx->launches++;
        ctx->stage_end(2);
    }
    ctx->stage_begin(3);
    for (int t = 2;; ++t) {
        const long long s = 1LL << (t - 1);
        const int RH = (int)((NH + s - 1) / s), RW = (int)((NW + s - 1) / s);
        if (RH <= 1 && RW <= 1) break;
        const int IH = (int)((NH + 2 * s - 1) / (2 * s)), IW = (int)((NW + 2 * s - 1) / (2 * s));
        const size_t total = (size_t)RH * RW * nz;
        const unsigned nb = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
        pyr_up_kernel<<<nb, 256, 0, st>>>(NH, NW, RH, RW, IH, IW, nz, dp, lp);
        ctx->launches++;
    }
    {
        const size_t total = (size_t)ll_h * ll_w * nz;
        const unsigned nb = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
        pyr_ll_kernel<<<nb, 256, 0, st>>>(coeffs, H, W, NH, NW, ll_h, ll_w, nz, dp, dpll, lpll);
        ctx->launches++;
    }
    ctx->stage_end(3);
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
