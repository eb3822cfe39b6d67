// RGB <-> IPT as stand-alone passes (replaces spiht/color_models.py:6-13 -> colour.convert(.., 'RGB', 'IPT') and
// back, colour-science 0.4.4).  Called by spihtb_convert_color and by the image path: launch_forward runs
// rgb_to_ipt_kernel into a float64 scratch image before the level-1 transform (dwt_fwd.cu) and launch_inverse runs
// ipt_to_rgb_kernel behind the level-1 synthesis (dwt_inv.cu).  The passes are NOT fused into the transform kernels
// (a warp of the transform owns one plane, the powers need all three); ipt.cuh holds the per-pixel
// arithmetic.
#include <algorithm>

#include "common.cuh"
#include "ipt.cuh"
#include "kernels.cuh"

namespace spihtb {

// tables of ipt_fast_pow into the current device's constant memory (once per context)
int ipt_upload_tables() { return ipt_upload_tables_tu(); }

// gridDim.y = images; the blocks of an image stride over its pixels (no division per pixel)
template <typename Tin>
__global__ void __launch_bounds__(256) rgb_to_ipt_kernel(const Tin *__restrict__ src, double *__restrict__ dst,
                                                         size_t plane)
{
    __shared__ PowTab s_tab;
    ipt_stage_table(&s_tab, false);
    const Tin *s = src + (size_t)blockIdx.y * 3 * plane;
    double *d = dst + (size_t)blockIdx.y * 3 * plane;
    // four pixels per thread and step, all loads first (the pass is latency-bound otherwise)
    constexpr int U = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o0 < plane; o0 += U * stride) {
        Tin r[U], g[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t o = o0 + u * stride;
            if (o < plane) {
                r[u] = s[o];
                g[u] = s[plane + o];
                b[u] = s[2 * plane + o];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t o = o0 + u * stride;
            if (o < plane) {
                double R = (double)r[u], G = (double)g[u], B = (double)b[u];
                if (sizeof(Tin) == 1) {  // uint8 pixels: imload's im / 255 (IEEE division, as numpy's)
                    R /= 255.0;
                    G /= 255.0;
                    B /= 255.0;
                }
                rgb_to_ipt_px(s_tab, R, G, B, d[o], d[plane + o], d[2 * plane + o]);
            }
        }
    }
}

template <typename Tout>
__global__ void __launch_bounds__(256) ipt_to_rgb_kernel(const double *__restrict__ src, Tout *__restrict__ dst,
                                                         size_t plane, const IptInv mi)
{
    __shared__ PowTab s_tab;
    ipt_stage_table(&s_tab, true);
    const double *s = src + (size_t)blockIdx.y * 3 * plane;
    Tout *d = dst + (size_t)blockIdx.y * 3 * plane;
    constexpr int U = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o0 < plane; o0 += U * stride) {
        double i[U], p[U], t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t o = o0 + u * stride;
            if (o < plane) {
                i[u] = s[o];
                p[u] = s[plane + o];
                t[u] = s[2 * plane + o];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t o = o0 + u * stride;
            if (o < plane) {
                double R, G, B;
                ipt_to_rgb_px(s_tab, mi, i[u], p[u], t[u], R, G, B);
                d[o] = (Tout)R;
                d[plane + o] = (Tout)G;
                d[2 * plane + o] = (Tout)B;
            }
        }
    }
}

static dim3 color_grid(spihtb_ctx *ctx, size_t plane, int B)
{
    // enough blocks per image to fill the device a few times over, at most 65535 images per launch (checked by callers)
    const size_t per_img = std::max<size_t>(1, std::min<size_t>((plane + 255) / 256, ((size_t)ctx->sm_count * 16 + B - 1) / B));
    return dim3((unsigned)per_img, (unsigned)B);
}

int launch_rgb_to_ipt(spihtb_ctx *ctx, const void *src, int src_dtype, double *dst, size_t plane, int B)
{
    if (B > 65535) {
        set_error("colour conversion: more than 65535 images in one call");
        return SPIHTB_EINVAL;
    }
    const dim3 nb = color_grid(ctx, plane, B);
    if (src_dtype == SPIHTB_F64)
        rgb_to_ipt_kernel<double><<<nb, 256, 0, ctx->stream>>>(static_cast<const double *>(src), dst, plane);
    else if (src_dtype == SPIHTB_U8)
        rgb_to_ipt_kernel<uint8_t><<<nb, 256, 0, ctx->stream>>>(static_cast<const uint8_t *>(src), dst, plane);
    else
        rgb_to_ipt_kernel<float><<<nb, 256, 0, ctx->stream>>>(static_cast<const float *>(src), dst, plane);
    ctx->launches++;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

int launch_ipt_to_rgb(spihtb_ctx *ctx, const double *src, void *dst, int dst_dtype, size_t plane, int B)
{
    const IptInv mi = make_ipt_inv();
    if (B > 65535) {
        set_error("colour conversion: more than 65535 images in one call");
        return SPIHTB_EINVAL;
    }
    const dim3 nb = color_grid(ctx, plane, B);
    if (dst_dtype == SPIHTB_F32)
        ipt_to_rgb_kernel<float><<<nb, 256, 0, ctx->stream>>>(src, static_cast<float *>(dst), plane, mi);
    else
        ipt_to_rgb_kernel<double><<<nb, 256, 0, ctx->stream>>>(src, static_cast<double *>(dst), plane, mi);
    ctx->launches++;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
