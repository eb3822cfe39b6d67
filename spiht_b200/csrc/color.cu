// RGB <-> IPT as stand-alone passes (spihtb_convert_color; replaces spiht/color_models.py:6-13 ->
// colour.convert(.., 'RGB', 'IPT') and back, colour-science 0.4.4).  The image path does not launch these:
// there the forward transform converts as it loads the level-1 rows and the inverse transform as it stores
// them (ipt.cuh holds the per-pixel arithmetic all of them share).
#include <algorithm>

#include "common.cuh"
#include "ipt.cuh"
#include "kernels.cuh"

namespace spihtb {

// tables of ipt_fast_pow into the current device's constant memory (once per context)
int ipt_upload_tables() { return ipt_upload_tables_tu(); }

template <typename Tin>
__global__ void __launch_bounds__(256) rgb_to_ipt_kernel(const Tin *__restrict__ src, double *__restrict__ dst,
                                                         size_t plane, size_t nimg)
{
    const size_t total = plane * nimg;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / plane, o = t - b * plane;
        const Tin *s = src + b * 3 * plane + o;
        double R = (double)s[0], G = (double)s[plane], B = (double)s[2 * plane];
        if (sizeof(Tin) == 1) {  // uint8 pixels: imload's im / 255 (IEEE division, as numpy's)
            R /= 255.0;
            G /= 255.0;
            B /= 255.0;
        }
        double *d = dst + b * 3 * plane + o;
        rgb_to_ipt_px(R, G, B, d[0], d[plane], d[2 * plane]);
    }
}

template <typename Tout>
__global__ void __launch_bounds__(256) ipt_to_rgb_kernel(const double *__restrict__ src, Tout *__restrict__ dst,
                                                         size_t plane, size_t nimg, const IptInv mi)
{
    const size_t total = plane * nimg;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / plane, o = t - b * plane;
        const double *s = src + b * 3 * plane + o;
        double R, G, B;
        ipt_to_rgb_px(mi, s[0], s[plane], s[2 * plane], R, G, B);
        Tout *d = dst + b * 3 * plane + o;
        d[0] = (Tout)R;
        d[plane] = (Tout)G;
        d[2 * plane] = (Tout)B;
    }
}

int launch_rgb_to_ipt(spihtb_ctx *ctx, const void *src, int src_dtype, double *dst, size_t plane, int B)
{
    const unsigned nb = (unsigned)std::min<size_t>((plane * B + 255) / 256, (size_t)ctx->sm_count * 32);
    if (src_dtype == SPIHTB_F64)
        rgb_to_ipt_kernel<double><<<nb, 256, 0, ctx->stream>>>(static_cast<const double *>(src), dst, plane, B);
    else if (src_dtype == SPIHTB_U8)
        rgb_to_ipt_kernel<uint8_t><<<nb, 256, 0, ctx->stream>>>(static_cast<const uint8_t *>(src), dst, plane, B);
    else
        rgb_to_ipt_kernel<float><<<nb, 256, 0, ctx->stream>>>(static_cast<const float *>(src), dst, plane, B);
    ctx->launches++;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

int launch_ipt_to_rgb(spihtb_ctx *ctx, const double *src, void *dst, int dst_dtype, size_t plane, int B)
{
    const IptInv mi = make_ipt_inv();
    const unsigned nb = (unsigned)std::min<size_t>((plane * B + 255) / 256, (size_t)ctx->sm_count * 32);
    if (dst_dtype == SPIHTB_F32)
        ipt_to_rgb_kernel<float><<<nb, 256, 0, ctx->stream>>>(src, static_cast<float *>(dst), plane, B, mi);
    else
        ipt_to_rgb_kernel<double><<<nb, 256, 0, ctx->stream>>>(src, static_cast<double *>(dst), plane, B, mi);
    ctx->launches++;
    SPIHTB_CUDA_CHECK(cudaGetLastError());
    return SPIHTB_OK;
}

}  // namespace spihtb
