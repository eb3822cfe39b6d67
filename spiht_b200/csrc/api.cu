// C ABI of libspiht_b200.so (see include/spiht_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "wavelets.cuh"

namespace spihtb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int dwt_max_level(int n, int f)
{
    // pywt.dwt_max_level: floor(log2(n / (f - 1))), 0 when n < f - 1
    if (f < 2 || n < f - 1) return 0;
    int q = n / (f - 1), l = 0;
    while ((q >> (l + 1)) > 0) ++l;
    return l;
}

static int coeff_len(int n, int f, int mode) { return mode == SPIHTB_MODE_PERIODIZATION ? (n + 1) / 2 : (n + f - 1) / 2; }

// every LL-root offspring (encoder_decoder.rs:50-62) must lie inside the array
static int check_coder_geom(int c, int h, int w, int ll_h, int ll_w)
{
    if (c <= 0 || h <= 0 || w <= 0) {
        set_error("coefficient array shape must be positive, got (%d,%d,%d)", c, h, w);
        return SPIHTB_EINVAL;
    }
    if (ll_h <= 1 || ll_w <= 1) {
        set_error("assertion failed: ll_h > 1 && ll_w > 1 (got %d, %d)", ll_h, ll_w);
        return SPIHTB_ELL;
    }
    // lowest offspring row of an LL root: 2 ll_h - 1 (ll_h even) or 2 ll_h - 2 (ll_h odd: the last odd root is
    // ll_h - 2, the even root ll_h - 1 has its block at rows ll_h - 1, ll_h)
    if (2LL * ll_h - (ll_h & 1) > h || 2LL * ll_w - (ll_w & 1) > w) {
        set_error("LL band %dx%d too large for a %dx%d array: root offspring out of bounds", ll_h, ll_w, h, w);
        return SPIHTB_EGEOM;
    }
    KeyFmt kf;
    if (!make_keyfmt(c, h, w, &kf)) {
        set_error("shape c=%d h=%d w=%d does not fit a 31-bit packed list entry", c, h, w);
        return SPIHTB_ESHAPE;
    }
    return SPIHTB_OK;
}

struct PyrBufs {
    uint8_t *dp, *lp, *dpll, *lpll;
    uint32_t *maxabs;
};

static int alloc_pyr(spihtb_ctx *ctx, int B, int C, int H, int W, int ll_h, int ll_w, PyrBufs *pb)
{
    const size_t nz = (size_t)B * C;
    const size_t nodes = ((nz * (H / 2) * (W / 2) + 255) / 256) * 256;
    const size_t roots = ((nz * ll_h * ll_w + 255) / 256) * 256;
    const size_t total = 2 * nodes + 2 * roots + (size_t)B * 4 + 256;
    int rc = ctx->ensure(ctx->pyr, total);
    if (rc) return rc;
    uint8_t *p = static_cast<uint8_t *>(ctx->pyr.p);
    pb->dp = p;
    pb->lp = p + nodes;
    pb->dpll = p + 2 * nodes;
    pb->lpll = p + 2 * nodes + roots;
    pb->maxabs = reinterpret_cast<uint32_t *>(p + 2 * nodes + 2 * roots);
    return SPIHTB_OK;
}

static uint64_t lis_shape_bound(int c, int h, int w, int ll_h, int ll_w)
{
    return (uint64_t)c * (h / 2 + 2) * (w / 2 + 2) * 5 / 4 + (uint64_t)c * ll_h * ll_w;
}

// bits a full encode can take when the top plane is `planes - 1`
static uint64_t stream_bits_bound(int c, int h, int w, int ll_h, int ll_w, int planes)
{
    const uint64_t chw = (uint64_t)c * h * w;
    return (uint64_t)planes * (chw + (uint64_t)c * ll_h * ll_w + lis_shape_bound(c, h, w, ll_h, ll_w)) + chw + 64;
}

static int fill_xform(XformArgs *x, int B, int C, const spihtb_geom *g, int color, const double *ch_scales, double q,
                      int pixel_dtype)
{
    if (!g || B <= 0 || C <= 0 || C > 8) {
        set_error("bad batch/channel count (B=%d, C=%d; C must be 1..8)", B, C);
        return SPIHTB_EINVAL;
    }
    if (color != SPIHTB_COLOR_NONE && color != SPIHTB_COLOR_IPT) {
        set_error("unknown colour model id %d", color);
        return SPIHTB_EINVAL;
    }
    if (pixel_dtype != SPIHTB_F32 && pixel_dtype != SPIHTB_F64 && pixel_dtype != SPIHTB_U8) {
        set_error("unknown pixel dtype %d", pixel_dtype);
        return SPIHTB_EINVAL;
    }
    x->B = B;
    x->C = C;
    x->g = *g;
    x->color = color;
    for (int c = 0; c < 8; ++c) x->scale[c] = (ch_scales && c < C) ? ch_scales[c] : 1.0;
    x->q = q;
    x->pixel_dtype = pixel_dtype;
    return SPIHTB_OK;
}

}  // namespace spihtb

using namespace spihtb;

int spihtb_ctx::ensure(DevBuf &b, size_t bytes)
{
    if (b.cap >= bytes) return SPIHTB_OK;
    if (b.p) {
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) e = cudaFree(b.p);
        if (e != cudaSuccess) {
            set_error("cudaFree failed: %s", cudaGetErrorString(e));
            return SPIHTB_ECUDA;
        }
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8;  // grow with headroom
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        return SPIHTB_ENOMEM;
    }
    b.cap = want;
    return SPIHTB_OK;
}

void spihtb_ctx::harvest(int s, bool all)
{
    StageProf &p = prof[s];
    // intervals [harvested, recorded) are pending; keep at most PROF_RING - 1 outstanding
    while (p.harvested < p.recorded && (all || p.recorded - p.harvested >= PROF_RING)) {
        const int slot = (int)(p.harvested % PROF_RING);
        float ms = 0.f;
        if (cudaEventSynchronize(p.b[slot]) == cudaSuccess && cudaEventElapsedTime(&ms, p.a[slot], p.b[slot]) == cudaSuccess)
            p.acc_ms += ms;
        p.harvested++;
    }
}

void spihtb_ctx::stage_begin(int s)
{
    if (!profiling) return;
    StageProf &p = prof[s];
    if (!p.made) {
        for (int i = 0; i < PROF_RING; ++i) {
            cudaEventCreate(&p.a[i]);
            cudaEventCreate(&p.b[i]);
        }
        p.made = true;
    }
    harvest(s, false);
    cudaEventRecord(p.a[p.recorded % PROF_RING], stream);
}

void spihtb_ctx::stage_end(int s)
{
    if (!profiling) return;
    StageProf &p = prof[s];
    cudaEventRecord(p.b[p.recorded % PROF_RING], stream);
    p.recorded++;
}

extern "C" {

int spihtb_profile_enable(spihtb_ctx *ctx, int enable)
{
    if (!ctx) {
        set_error("ctx is null");
        return SPIHTB_EINVAL;
    }
    ctx->profiling = enable != 0;
    for (int i = 0; i < ctx->nsubs; ++i) ctx->subs[i]->profiling = ctx->profiling;
    return SPIHTB_OK;
}

int spihtb_set_option(spihtb_ctx *ctx, int32_t option, int64_t value)
{
    if (!ctx) {
        set_error("ctx is null");
        return SPIHTB_EINVAL;
    }
    switch (option) {
        case SPIHTB_OPT_SCRATCH_COEFFS: ctx->scratch_coeffs = value != 0; return SPIHTB_OK;
        default: set_error("unknown option %d", (int)option); return SPIHTB_EINVAL;
    }
}

int spihtb_profile_read(spihtb_ctx *ctx, double *ms_out, int64_t *count_out, int reset)
{
    if (!ctx || !ms_out || !count_out) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    for (int s = 0; s < SPIHTB_NSTAGES; ++s) {
        ms_out[s] = 0.0;
        count_out[s] = 0;
    }
    // the context's own timers plus those of the group pipeline's sub-contexts (their intervals may overlap in
    // time: the sums are per-stage busy times, not a partition of the step)
    for (int i = -1; i < ctx->nsubs; ++i) {
        spihtb_ctx *c = i < 0 ? ctx : ctx->subs[i];
        for (int s = 0; s < SPIHTB_NSTAGES; ++s) {
            if (c->prof[s].made) c->harvest(s, true);
            ms_out[s] += c->prof[s].acc_ms;
            count_out[s] += c->prof[s].harvested;
            if (reset) {
                c->prof[s].acc_ms = 0.0;
                c->prof[s].recorded = 0;
                c->prof[s].harvested = 0;
            }
        }
    }
    return SPIHTB_OK;
}

int spihtb_version(void) { return SPIHTB_VERSION; }
const char *spihtb_last_error(void) { return g_err; }

int spihtb_create(int device, spihtb_ctx **out)
{
    if (!out) {
        set_error("ctx out pointer is null");
        return SPIHTB_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no usable CUDA device (%s); libspiht_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return SPIHTB_ECUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range (0..%d)", device, ndev - 1);
        return SPIHTB_EINVAL;
    }
    SPIHTB_CUDA_CHECK(cudaSetDevice(device));
    spihtb_ctx *c = new spihtb_ctx();
    c->device = device;
    int sms = 0;
    SPIHTB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    c->sm_count = sms > 0 ? sms : 148;
    c->stream = nullptr;
    SPIHTB_CUDA_CHECK(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
    {
        const int rc = ipt_upload_tables();
        if (rc) return rc;
    }
    *out = c;
    return SPIHTB_OK;
}

int spihtb_destroy(spihtb_ctx *ctx)
{
    if (!ctx) return SPIHTB_OK;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < ctx->nsubs; ++i) {
        spihtb_ctx *c = ctx->subs[i];
        cudaStreamSynchronize(c->coder_stream);
        cudaStreamDestroy(c->coder_stream);
        cudaEventDestroy(c->ev_pyr);
        cudaEventDestroy(c->ev_done);
        cudaStream_t st = c->stream;
        spihtb_destroy(c);     // synchronises c->stream, frees its workspaces
        cudaStreamDestroy(st);
    }
    ctx->nsubs = 0;
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    if (!ctx->is_sub) cudaStreamSynchronize(ctx->stream);
    else cudaStreamSynchronize(ctx->stream);
    if (ctx->aux) {
        cudaStreamSynchronize(ctx->aux);
        cudaStreamDestroy(ctx->aux);
        cudaEventDestroy(ctx->ev_fork);
        cudaEventDestroy(ctx->ev_join);
        cudaEventDestroy(ctx->ev_order);
    }
    DevBuf *bufs[] = {&ctx->pyr, &ctx->lists, &ctx->misc, &ctx->tmpa, &ctx->tmpb, &ctx->tail, &ctx->io, &ctx->io2, &ctx->fix, &ctx->u8lut, &ctx->blk};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (int s = 0; s < SPIHTB_NSTAGES; ++s)
        if (ctx->prof[s].made)
            for (int i = 0; i < PROF_RING; ++i) {
                cudaEventDestroy(ctx->prof[s].a[i]);
                cudaEventDestroy(ctx->prof[s].b[i]);
            }
    delete ctx;
    return SPIHTB_OK;
}

int spihtb_set_stream(spihtb_ctx *ctx, void *cuda_stream)
{
    if (!ctx) {
        set_error("ctx is null");
        return SPIHTB_EINVAL;
    }
    cudaStream_t ns = static_cast<cudaStream_t>(cuda_stream);
    if (ns != ctx->stream) {
        // The workspaces (pyramid planes, lists, scratch, work counters) are shared by every call on this context:
        // work already queued on the old stream must finish before anything queued on the new one touches them.
        SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
        SPIHTB_CUDA_CHECK(cudaEventRecord(ctx->ev_order, ctx->stream));
        SPIHTB_CUDA_CHECK(cudaStreamWaitEvent(ns, ctx->ev_order, 0));
        ctx->stream = ns;
    }
    return SPIHTB_OK;
}

int spihtb_sync(spihtb_ctx *ctx)
{
    if (!ctx) {
        set_error("ctx is null");
        return SPIHTB_EINVAL;
    }
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPIHTB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return SPIHTB_OK;
}

int64_t spihtb_launch_count(spihtb_ctx *ctx)
{
    if (!ctx) return 0;
    int64_t n = ctx->launches;
    for (int i = 0; i < ctx->nsubs; ++i) n += ctx->subs[i]->launches;
    return n;
}
int spihtb_forward_path(spihtb_ctx *ctx) { return ctx && ctx->last_forward_fused12 ? 12 : 1; }

int spihtb_plan(int32_t h, int32_t w, int32_t wavelet, int32_t mode, int32_t level, spihtb_geom *g)
{
    if (!g) {
        set_error("geom out pointer is null");
        return SPIHTB_EINVAL;
    }
    const int F = wavelet_flen(wavelet);
    if (F == 0) {
        set_error("unknown wavelet id %d", wavelet);
        return SPIHTB_EINVAL;
    }
    if (mode != SPIHTB_MODE_REFLECT && mode != SPIHTB_MODE_SYMMETRIC && mode != SPIHTB_MODE_PERIODIZATION) {
        set_error("unknown mode id %d", mode);
        return SPIHTB_EINVAL;
    }
    if (h <= 0 || w <= 0) {
        set_error("image size must be positive, got %dx%d", h, w);
        return SPIHTB_EINVAL;
    }
    const int maxlev = std::min(dwt_max_level(h, F), dwt_max_level(w, F));
    int L = level < 0 ? maxlev : level;
    if (L == 0) {
        set_error("decomposition level is 0 for a %dx%d image with filter length %d: no detail band to code", h, w, F);
        return SPIHTB_ELEVEL;
    }
    if (L > SPIHTB_MAX_LEVELS) {
        set_error("level %d exceeds the supported maximum %d", L, SPIHTB_MAX_LEVELS);
        return SPIHTB_ELEVEL;
    }
    memset(g, 0, sizeof(*g));
    g->h = h;
    g->w = w;
    g->wavelet = wavelet;
    g->mode = mode;
    g->levels = L;
    int ch = h, cw = w;
    for (int l = 0; l < L; ++l) {
        g->in_h[l] = ch;
        g->in_w[l] = cw;
        ch = coeff_len(ch, F, mode);
        cw = coeff_len(cw, F, mode);
        g->band_h[l] = ch;
        g->band_w[l] = cw;
    }
    g->ll_h = ch;
    g->ll_w = cw;
    int sh = ch, sw = cw;
    for (int l = L - 1; l >= 0; --l) {
        g->off_h[l] = sh;
        g->off_w[l] = sw;
        sh += g->band_h[l];
        sw += g->band_w[l];
    }
    g->enc_h = sh;
    g->enc_w = sw;
    const bool per = mode == SPIHTB_MODE_PERIODIZATION;
    g->rec_h = per ? 2 * g->band_h[0] : 2 * g->band_h[0] - F + 2;
    g->rec_w = per ? 2 * g->band_w[0] : 2 * g->band_w[0] - F + 2;
    return SPIHTB_OK;
}

int spihtb_wavelet_filters(int32_t wavelet, int32_t *flen, double *dec_lo, double *rec_lo)
{
    const int F = wavelet_flen(wavelet);
    if (!flen || !dec_lo || !rec_lo || F == 0) {
        set_error(F == 0 ? "unknown wavelet id %d" : "null pointer (wavelet id %d)", wavelet);
        return SPIHTB_EINVAL;
    }
    *flen = F;
    switch (wavelet) {
#define SPIHTB_COPY_FILTERS(WID)                     \
    for (int i = 0; i < Wav<WID>::F; ++i) {          \
        dec_lo[i] = Wav<WID>::dec_lo(i);             \
        rec_lo[i] = Wav<WID>::rec_lo(i);             \
    }
        case SPIHTB_WAVELET_BIOR22: SPIHTB_COPY_FILTERS(SPIHTB_WAVELET_BIOR22) break;
        case SPIHTB_WAVELET_BIOR44: SPIHTB_COPY_FILTERS(SPIHTB_WAVELET_BIOR44) break;
        case SPIHTB_WAVELET_BIOR68: SPIHTB_COPY_FILTERS(SPIHTB_WAVELET_BIOR68) break;
#undef SPIHTB_COPY_FILTERS
        default: {
            double dl[SPIHTB_GEN_MAXF], rl[SPIHTB_GEN_MAXF];
            int f2 = 0;
            if (!generic_wavelet_taps(wavelet, &f2, dl, rl) || f2 != F) {
                set_error("no filter bank for wavelet id %d", wavelet);
                return SPIHTB_EINVAL;
            }
            for (int i = 0; i < F; ++i) {
                dec_lo[i] = dl[i];
                rec_lo[i] = rl[i];
            }
        }
    }
    return SPIHTB_OK;
}

uint64_t spihtb_stream_bound(int32_t c, int32_t h, int32_t w, int32_t ll_h, int32_t ll_w)
{
    const uint64_t bits = stream_bits_bound(c, h, w, ll_h, ll_w, 31);
    return ((bits + 7) / 8 + 15) / 8 * 8;
}

}  // extern "C"

// pyramid (all of it, or the rings when the forward transform already wrote the base pass) + coder
static int encode_with_pyramid(spihtb_ctx *ctx, const int32_t *dev_coeffs, int B, int c, int h, int w, int ll_h, int ll_w,
                               const spihtb::PyrBufs &pb, bool base_done, uint64_t max_bits,
                               const uint64_t *dev_max_bits, uint8_t *dev_out, uint64_t out_stride,
                               uint64_t *dev_nbits, int32_t *dev_max_n, int32_t *dev_status)
{
    using namespace spihtb;
    int rc = launch_pyramid(ctx, dev_coeffs, B, c, h, w, ll_h, ll_w, pb.dp, pb.lp, pb.dpll, pb.lpll, pb.maxabs, base_done);
    if (rc) return rc;
    EncArgs a;
    a.coeffs = dev_coeffs;
    a.B = B; a.C = c; a.H = h; a.W = w; a.ll_h = ll_h; a.ll_w = ll_w;
    a.dp = pb.dp; a.lp = pb.lp; a.dpll = pb.dpll; a.lpll = pb.lpll;
    a.maxabs = pb.maxabs;
    a.max_bits = max_bits;
    a.dev_max_bits = dev_max_bits;
    a.out = dev_out;
    a.out_stride = out_stride;
    a.nbits = dev_nbits;
    a.max_n = dev_max_n;
    a.status = dev_status;
    if (ctx->is_sub && ctx->coder_stream) {
        // pipeline: the coder runs on the sub-context's high-priority stream, behind the pyramid
        SPIHTB_CUDA_CHECK(cudaEventRecord(ctx->ev_pyr, ctx->stream));
        SPIHTB_CUDA_CHECK(cudaStreamWaitEvent(ctx->coder_stream, ctx->ev_pyr, 0));
        cudaStream_t keep = ctx->stream;
        ctx->stream = ctx->coder_stream;
        rc = launch_encode(ctx, a);
        ctx->stream = keep;
        return rc;
    }
    return launch_encode(ctx, a);
}

// sub-contexts of the group pipeline (created on first use)
static int ensure_subs(spihtb_ctx *ctx, int n)
{
    using namespace spihtb;
    n = std::min(n, (int)spihtb_ctx::MAX_SUBS);
    int lo = 0, hi = 0;
    SPIHTB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo: least priority (largest number)
    for (int i = ctx->nsubs; i < n; ++i) {
        spihtb_ctx *c = new spihtb_ctx();
        c->device = ctx->device;
        c->sm_count = ctx->sm_count;
        c->is_sub = true;
        SPIHTB_CUDA_CHECK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, lo));
        SPIHTB_CUDA_CHECK(cudaStreamCreateWithPriority(&c->coder_stream, cudaStreamNonBlocking, hi));
        SPIHTB_CUDA_CHECK(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
        SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
        SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_pyr, cudaEventDisableTiming));
        SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
        c->profiling = ctx->profiling;
        ctx->subs[i] = c;
        ctx->nsubs = i + 1;
    }
    if (!ctx->ev_in) SPIHTB_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
    return SPIHTB_OK;
}

extern "C" {

int spihtb_encode_coeffs(spihtb_ctx *ctx, const int32_t *dev_coeffs, int32_t B, int32_t c, int32_t h, int32_t w,
                         int32_t ll_h, int32_t ll_w, uint64_t max_bits, const uint64_t *dev_max_bits,
                         uint8_t *dev_out, uint64_t out_stride, uint64_t *dev_nbits, int32_t *dev_max_n,
                         int32_t *dev_status)
{
    if (!ctx || !dev_coeffs || !dev_out || !dev_nbits || !dev_max_n || B <= 0) {
        set_error("null pointer or empty batch");
        return SPIHTB_EINVAL;
    }
    int rc = check_coder_geom(c, h, w, ll_h, ll_w);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    PyrBufs pb;
    rc = alloc_pyr(ctx, B, c, h, w, ll_h, ll_w, &pb);
    if (rc) return rc;
    return encode_with_pyramid(ctx, dev_coeffs, B, c, h, w, ll_h, ll_w, pb, false, max_bits, dev_max_bits, dev_out,
                               out_stride, dev_nbits, dev_max_n, dev_status);
}

int spihtb_max_abs(spihtb_ctx *ctx, const int32_t *dev_coeffs, int32_t B, uint64_t per_image, uint32_t *dev_out)
{
    if (!ctx || !dev_coeffs || !dev_out || B <= 0 || B > 65535 || per_image == 0) {
        set_error("null pointer, empty batch or more than 65535 images");
        return SPIHTB_EINVAL;
    }
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launch_max_abs(ctx, dev_coeffs, B, (size_t)per_image, dev_out);
}

int spihtb_decode_coeffs(spihtb_ctx *ctx, const uint8_t *dev_in, uint64_t in_stride, const uint64_t *dev_nbytes,
                         const int32_t *dev_n, int32_t B, int32_t c, int32_t h, int32_t w, int32_t ll_h,
                         int32_t ll_w, int32_t *dev_coeffs_out)
{
    if (!ctx || !dev_in || !dev_nbytes || !dev_n || !dev_coeffs_out || B <= 0) {
        set_error("null pointer or empty batch");
        return SPIHTB_EINVAL;
    }
    int rc = check_coder_geom(c, h, w, ll_h, ll_w);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    DecArgs a;
    a.in = dev_in;
    a.in_stride = in_stride;
    a.nbytes = dev_nbytes;
    a.n = dev_n;
    a.B = B; a.C = c; a.H = h; a.W = w; a.ll_h = ll_h; a.ll_w = ll_w;
    a.out = dev_coeffs_out;
    return launch_decode(ctx, a);
}

int spihtb_encode(spihtb_ctx *ctx, const int32_t *host_coeffs, int32_t c, int32_t h, int32_t w, int32_t ll_h,
                  int32_t ll_w, uint64_t max_bits, const uint8_t **out_bytes, uint64_t *out_nbits,
                  int32_t *out_max_n)
{
    if (!ctx || !host_coeffs || !out_bytes || !out_nbits || !out_max_n) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    int rc = check_coder_geom(c, h, w, ll_h, ll_w);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)c * h * w;
    rc = ctx->ensure(ctx->io, n * sizeof(int32_t) + 256);
    if (rc) return rc;
    int32_t *d_coeffs = static_cast<int32_t *>(ctx->io.p);
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(d_coeffs, host_coeffs, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));

    PyrBufs pb;
    rc = alloc_pyr(ctx, 1, c, h, w, ll_h, ll_w, &pb);
    if (rc) return rc;
    rc = launch_pyramid(ctx, d_coeffs, 1, c, h, w, ll_h, ll_w, pb.dp, pb.lp, pb.dpll, pb.lpll, pb.maxabs);
    if (rc) return rc;
    // size the stream: the budget, or the worst case for the top plane found
    uint32_t maxabs = 0;
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(&maxabs, pb.maxabs, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    int planes = 2;
    while (planes < 33 && (maxabs >> (planes - 1)) != 0) ++planes;  // floor(log2 max) + 2 >= max_n + 1
    uint64_t bits = stream_bits_bound(c, h, w, ll_h, ll_w, planes);
    if (max_bits != 0 && max_bits < bits) bits = max_bits;
    const uint64_t stride = ((bits + 7) / 8 + 15) / 8 * 8;
    rc = ctx->ensure(ctx->io2, stride + 64);
    if (rc) return rc;
    uint8_t *d_out = static_cast<uint8_t *>(ctx->io2.p);
    uint64_t *d_nbits = reinterpret_cast<uint64_t *>(d_out + stride);
    int32_t *d_max_n = reinterpret_cast<int32_t *>(d_out + stride + 8);
    int32_t *d_status = d_max_n + 1;

    EncArgs a;
    a.coeffs = d_coeffs;
    a.B = 1; a.C = c; a.H = h; a.W = w; a.ll_h = ll_h; a.ll_w = ll_w;
    a.dp = pb.dp; a.lp = pb.lp; a.dpll = pb.dpll; a.lpll = pb.lpll;
    a.maxabs = pb.maxabs;
    a.max_bits = max_bits;
    a.dev_max_bits = nullptr;
    a.out = d_out;
    a.out_stride = stride;
    a.nbits = d_nbits;
    a.max_n = d_max_n;
    a.status = d_status;
    rc = launch_encode(ctx, a);
    if (rc) return rc;
    struct { uint64_t nbits; int32_t max_n; int32_t status; } res;
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(&res, d_nbits, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (res.status & 1) {
        set_error("internal: stream bound %llu bits too small", (unsigned long long)bits);
        return SPIHTB_ECAP;
    }
    const size_t nbytes = (size_t)((res.nbits + 7) / 8);
    ctx->host_out.resize(nbytes ? nbytes : 1);
    if (nbytes) {
        SPIHTB_CUDA_CHECK(cudaMemcpyAsync(ctx->host_out.data(), d_out, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
        SPIHTB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    *out_bytes = ctx->host_out.data();
    *out_nbits = res.nbits;
    *out_max_n = res.max_n;
    return SPIHTB_OK;
}

int spihtb_decode(spihtb_ctx *ctx, const uint8_t *host_data, uint64_t nbytes, int32_t n, int32_t c, int32_t h,
                  int32_t w, int32_t ll_h, int32_t ll_w, int32_t *host_out)
{
    if (!ctx || (!host_data && nbytes) || !host_out) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    if (n < 0 || n > 255) {
        set_error("n must fit a u8, got %d", n);
        return SPIHTB_EINVAL;
    }
    int rc = check_coder_geom(c, h, w, ll_h, ll_w);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t ncoef = (size_t)c * h * w;
    const uint64_t stride = (nbytes + 15) / 8 * 8;
    rc = ctx->ensure(ctx->io, ncoef * sizeof(int32_t) + 256);
    if (rc) return rc;
    rc = ctx->ensure(ctx->io2, stride + 64);
    if (rc) return rc;
    uint8_t *d_in = static_cast<uint8_t *>(ctx->io2.p);
    uint64_t *d_nbytes = reinterpret_cast<uint64_t *>(d_in + stride);
    int32_t *d_n = reinterpret_cast<int32_t *>(d_in + stride + 8);
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(d_in, 0, stride + 64, ctx->stream));
    if (nbytes)
        SPIHTB_CUDA_CHECK(cudaMemcpyAsync(d_in, host_data, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    struct { uint64_t nb; int32_t n; } hdr = {nbytes, n};
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(d_nbytes, &hdr, 12, cudaMemcpyHostToDevice, ctx->stream));
    DecArgs a;
    a.in = d_in;
    a.in_stride = stride;
    a.nbytes = d_nbytes;
    a.n = d_n;
    a.B = 1; a.C = c; a.H = h; a.W = w; a.ll_h = ll_h; a.ll_w = ll_w;
    a.out = static_cast<int32_t *>(ctx->io.p);
    rc = launch_decode(ctx, a);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(host_out, ctx->io.p, ncoef * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return SPIHTB_OK;
}

int spihtb_decode_with_metadata(spihtb_ctx *ctx, const uint8_t *host_data, uint64_t nbytes, int32_t n, int32_t c,
                                int32_t h, int32_t w, int32_t ll_h, int32_t ll_w, const int32_t *top_slice,
                                const int32_t *other_slices, int32_t levels, int32_t *host_out, int32_t *host_meta)
{
    if (!ctx || (!host_data && nbytes) || !host_out || !host_meta || !top_slice || (!other_slices && levels > 0)) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    if (n < 0 || n > 255 || levels < 0 || levels > 255) {
        set_error("n and the number of levels must fit a u8 (got %d, %d)", n, levels);
        return SPIHTB_EINVAL;
    }
    int rc = check_coder_geom(c, h, w, ll_h, ll_w);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t ncoef = (size_t)c * h * w;
    const uint64_t stride = (nbytes + 15) / 8 * 8;
    const uint64_t rows = nbytes * 8 + 1;
    const size_t nsl = (size_t)levels * 12;
    // io: coefficients | metadata rows;  io2: stream | header | slices | error flag
    rc = ctx->ensure(ctx->io, (ncoef + rows * 8) * sizeof(int32_t) + 512);
    if (rc) return rc;
    rc = ctx->ensure(ctx->io2, stride + 64 + nsl * sizeof(int32_t) + 64);
    if (rc) return rc;
    int32_t *d_coef = static_cast<int32_t *>(ctx->io.p);
    int32_t *d_meta = d_coef + (ncoef + 63) / 64 * 64;   // 256-byte aligned rows
    uint8_t *d_in = static_cast<uint8_t *>(ctx->io2.p);
    uint64_t *d_nbytes = reinterpret_cast<uint64_t *>(d_in + stride);
    int32_t *d_n = reinterpret_cast<int32_t *>(d_in + stride + 8);
    int32_t *d_err = d_n + 1;
    int32_t *d_slices = reinterpret_cast<int32_t *>(d_in + stride + 64);
    rc = ctx->ensure(ctx->io, ((ncoef + 63) / 64 * 64 + rows * 8) * sizeof(int32_t) + 512);
    if (rc) return rc;
    d_coef = static_cast<int32_t *>(ctx->io.p);
    d_meta = d_coef + (ncoef + 63) / 64 * 64;
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(d_in, 0, stride + 64, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(d_meta, 0, rows * 8 * sizeof(int32_t), ctx->stream));
    if (nbytes) SPIHTB_CUDA_CHECK(cudaMemcpyAsync(d_in, host_data, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    struct { uint64_t nb; int32_t n; int32_t err; } hdr = {nbytes, n, 0};
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(d_nbytes, &hdr, 16, cudaMemcpyHostToDevice, ctx->stream));
    if (nsl)
        SPIHTB_CUDA_CHECK(cudaMemcpyAsync(d_slices, other_slices, nsl * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    DecArgs a;
    a.in = d_in;
    a.in_stride = stride;
    a.nbytes = d_nbytes;
    a.n = d_n;
    a.B = 1; a.C = c; a.H = h; a.W = w; a.ll_h = ll_h; a.ll_w = ll_w;
    a.out = d_coef;
    a.meta = d_meta;
    a.meta_rows = rows;
    a.level = levels;
    a.top_ei = top_slice[1];
    a.top_ej = top_slice[3];
    a.slices = d_slices;
    a.meta_err = d_err;
    rc = launch_decode(ctx, a);
    if (rc) return rc;
    int32_t err = 0;
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(host_out, d_coef, ncoef * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(host_meta, d_meta, rows * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, ctx->stream));
    SPIHTB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (err) {
        set_error("band rectangle index out of range: the tree is deeper than the %d levels of slices given "
                  "(the reference panics here)", levels);
        return SPIHTB_EGEOM;
    }
    return SPIHTB_OK;
}

int spihtb_forward(spihtb_ctx *ctx, const void *dev_pixels, int32_t pixel_dtype, int32_t B, int32_t C,
                   const spihtb_geom *geom, int32_t color_model, const double *ch_scales, double q,
                   int32_t *dev_coeffs)
{
    if (!ctx || !dev_pixels || !dev_coeffs) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    XformArgs x;
    int rc = fill_xform(&x, B, C, geom, color_model, ch_scales, q, pixel_dtype);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launch_forward(ctx, dev_pixels, x, dev_coeffs);
}

int spihtb_inverse(spihtb_ctx *ctx, const int32_t *dev_coeffs, int32_t B, int32_t C, const spihtb_geom *geom,
                   int32_t color_model, const double *ch_scales, double q, void *dev_pixels_out,
                   int32_t pixel_dtype)
{
    if (!ctx || !dev_pixels_out || !dev_coeffs) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    XformArgs x;
    int rc = fill_xform(&x, B, C, geom, color_model, ch_scales, q, pixel_dtype);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launch_inverse(ctx, dev_coeffs, x, dev_pixels_out);
}

int spihtb_convert_color(spihtb_ctx *ctx, const void *dev_in, int32_t in_dtype, int32_t B, uint64_t plane,
                         int32_t src_model, int32_t dst_model, void *dev_out, int32_t out_dtype)
{
    if (!ctx || !dev_in || !dev_out || B <= 0 || plane == 0) {
        set_error("null pointer or empty batch");
        return SPIHTB_EINVAL;
    }
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (src_model == SPIHTB_COLOR_NONE && dst_model == SPIHTB_COLOR_IPT) {
        if ((in_dtype != SPIHTB_F32 && in_dtype != SPIHTB_F64 && in_dtype != SPIHTB_U8) || out_dtype != SPIHTB_F64) {
            set_error("RGB -> IPT takes float32 / float64 / uint8 pixels and writes float64");
            return SPIHTB_EINVAL;
        }
        return launch_rgb_to_ipt(ctx, dev_in, in_dtype, static_cast<double *>(dev_out), (size_t)plane, B);
    }
    if (src_model == SPIHTB_COLOR_IPT && dst_model == SPIHTB_COLOR_NONE) {
        if (in_dtype != SPIHTB_F64 || (out_dtype != SPIHTB_F32 && out_dtype != SPIHTB_F64)) {
            set_error("IPT -> RGB takes float64 and writes float32 / float64");
            return SPIHTB_EINVAL;
        }
        return launch_ipt_to_rgb(ctx, static_cast<const double *>(dev_in), dev_out, out_dtype, (size_t)plane, B);
    }
    set_error("unsupported colour conversion %d -> %d (RGB <-> IPT only)", src_model, dst_model);
    return SPIHTB_EINVAL;
}

int spihtb_encode_images(spihtb_ctx *ctx, const void *dev_pixels, int32_t pixel_dtype, int32_t B, int32_t C,
                         const spihtb_geom *geom, int32_t color_model, const double *ch_scales, double q,
                         uint64_t max_bits, const uint64_t *dev_max_bits, int32_t *dev_coeffs_scratch,
                         uint8_t *dev_out, uint64_t out_stride, uint64_t *dev_nbits, int32_t *dev_max_n,
                         int32_t *dev_status)
{
    if (!geom || !dev_coeffs_scratch) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    int rc = check_coder_geom(C, geom->enc_h, geom->enc_w, geom->ll_h, geom->ll_w);
    if (rc) return rc;
    // forward transform with the pyramid base pass fused into its epilogue, then rings + coder
    XformArgs x;
    rc = fill_xform(&x, B, C, geom, color_model, ch_scales, q, pixel_dtype);
    if (rc) return rc;
    if (!ctx || !dev_pixels || !dev_out || !dev_nbits || !dev_max_n || B <= 0) {
        set_error("null pointer or empty batch");
        return SPIHTB_EINVAL;
    }
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    PyrBufs pb;
    rc = alloc_pyr(ctx, B, C, geom->enc_h, geom->enc_w, geom->ll_h, geom->ll_w, &pb);
    if (rc) return rc;
    // the pyramid's base pass rides on the transform's epilogue, except for the wavelets of dwt_gen.cu (stand-alone pass)
    const bool fuse_base = !wavelet_is_generic(geom->wavelet);
    // ---- group pipeline (see spihtb_ctx::subs): groups of G images on alternating sub-contexts
    int G = 0, NS = 3;
    if (const char *e = getenv("SPIHTB_GROUP")) G = atoi(e);
    if (const char *e = getenv("SPIHTB_GROUP_STREAMS")) NS = std::max(1, std::min(atoi(e), (int)spihtb_ctx::MAX_SUBS));
    if (G > 0 && G < B && !ctx->is_sub) {
        rc = ensure_subs(ctx, NS);
        if (rc) return rc;
        SPIHTB_CUDA_CHECK(cudaEventRecord(ctx->ev_in, ctx->stream));
        const size_t esz = pixel_dtype == SPIHTB_F64 ? 8 : (pixel_dtype == SPIHTB_U8 ? 1 : 4);
        const size_t px_img = (size_t)C * geom->h * geom->w * esz;
        const size_t co_img = (size_t)C * geom->enc_h * geom->enc_w;
        const size_t nodes_img = (size_t)C * (geom->enc_h / 2) * (geom->enc_w / 2);
        const size_t roots_img = (size_t)C * geom->ll_h * geom->ll_w;
        int g = 0;
        for (int b0 = 0; b0 < B; b0 += G, ++g) {
            spihtb_ctx *sc = ctx->subs[g % NS];
            const int nb = std::min(G, B - b0);
            SPIHTB_CUDA_CHECK(cudaStreamWaitEvent(sc->stream, ctx->ev_in, 0));
            XformArgs xg = x;
            xg.B = nb;
            PyrBufs pg;
            pg.dp = pb.dp + b0 * nodes_img;
            pg.lp = pb.lp + b0 * nodes_img;
            pg.dpll = pb.dpll + b0 * roots_img;
            pg.lpll = pb.lpll + b0 * roots_img;
            pg.maxabs = pb.maxabs + b0;
            const PyrFuse pfg = {pg.dp, pg.maxabs};
            int32_t *cg = dev_coeffs_scratch + b0 * co_img;
            rc = launch_forward(sc, static_cast<const char *>(dev_pixels) + b0 * px_img, xg, cg, fuse_base ? &pfg : nullptr);
            if (rc) return rc;
            rc = encode_with_pyramid(sc, cg, nb, C, geom->enc_h, geom->enc_w, geom->ll_h, geom->ll_w, pg, fuse_base, max_bits,
                                     dev_max_bits ? dev_max_bits + b0 : nullptr, dev_out + (size_t)b0 * out_stride,
                                     out_stride, dev_nbits + b0, dev_max_n + b0, dev_status ? dev_status + b0 : nullptr);
            if (rc) return rc;
            SPIHTB_CUDA_CHECK(cudaEventRecord(sc->ev_done, sc->coder_stream));
        }
        for (int i = 0; i < std::min(g, NS); ++i) SPIHTB_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->subs[i]->ev_done, 0));
        ctx->last_forward_fused12 = ctx->subs[0]->last_forward_fused12;
        return SPIHTB_OK;
    }
    const PyrFuse pf = {pb.dp, pb.maxabs};
    rc = launch_forward(ctx, dev_pixels, x, dev_coeffs_scratch, fuse_base ? &pf : nullptr);
    if (rc) return rc;
    return encode_with_pyramid(ctx, dev_coeffs_scratch, B, C, geom->enc_h, geom->enc_w, geom->ll_h, geom->ll_w, pb,
                               fuse_base, max_bits, dev_max_bits, dev_out, out_stride, dev_nbits, dev_max_n, dev_status);
}

int spihtb_decode_images(spihtb_ctx *ctx, const uint8_t *dev_in, uint64_t in_stride, const uint64_t *dev_nbytes,
                         const int32_t *dev_n, int32_t B, int32_t C, const spihtb_geom *geom, int32_t color_model,
                         const double *ch_scales, double q, int32_t *dev_coeffs_scratch, void *dev_pixels_out,
                         int32_t pixel_dtype)
{
    if (!geom || !dev_coeffs_scratch) {
        set_error("null pointer");
        return SPIHTB_EINVAL;
    }
    if (!ctx || !dev_in || !dev_nbytes || !dev_n || !dev_pixels_out || B <= 0) {
        set_error("null pointer or empty batch");
        return SPIHTB_EINVAL;
    }
    const int c = C, h = geom->enc_h, w = geom->enc_w;
    int rc = check_coder_geom(c, h, w, geom->ll_h, geom->ll_w);
    if (rc) return rc;
    XformArgs x;
    rc = fill_xform(&x, B, C, geom, color_model, ch_scales, q, pixel_dtype);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaSetDevice(ctx->device));
    // the decoder marks the 64x64 blocks of the array it writes into; the inverse transform skips the detail
    // bands of tasks with no marked block (at low rates the finest levels hold no coefficient at all)
    const size_t nblk = (size_t)B * C * ((h + 63) / 64) * ((w + 63) / 64);
    rc = ctx->ensure(ctx->blk, 2 * nblk + (size_t)B + 256);
    if (rc) return rc;
    SPIHTB_CUDA_CHECK(cudaMemsetAsync(ctx->blk.p, 0, 2 * nblk + (size_t)B, ctx->stream));
    DecArgs a;
    a.in = dev_in;
    a.in_stride = in_stride;
    a.nbytes = dev_nbytes;
    a.n = dev_n;
    a.B = B; a.C = c; a.H = h; a.W = w; a.ll_h = geom->ll_h; a.ll_w = geom->ll_w;
    a.out = dev_coeffs_scratch;
    a.blk = static_cast<uint8_t *>(ctx->blk.p);
    // marks of the finest level alone (its bands start at row off_h[0] / column off_w[0] of the array)
    a.blk1 = a.blk + nblk;
    a.l1_any = a.blk + 2 * nblk;
    a.fine_h0 = geom->off_h[0];
    a.fine_w0 = geom->off_w[0];
    // SPIHTB_OPT_SCRATCH_COEFFS: the caller does not read the coefficient array, so the finest detail bands (three
    // quarters of it) are zeroed only for the images whose streams reach them
    // Worth it only where streams rarely reach those bands: an image that does must zero them with one CTA, which is
    // far slower than the memset it replaces.  The row stride bounds the rate: lazy up to 0.75 bit per pixel.
    // (SPIHTB_LAZY_ZERO=1 / SPIHTB_NO_LAZY_ZERO=1 force it on / off: tests and A-B runs)
    a.lazy_zero = ctx->scratch_coeffs && !((geom->ll_h | geom->ll_w) & 1) && getenv("SPIHTB_NO_LAZY_ZERO") == nullptr &&
                  (getenv("SPIHTB_LAZY_ZERO") != nullptr ||
                   (double)in_stride * 8.0 <= 0.75 * (double)geom->h * (double)geom->w);
    rc = launch_decode(ctx, a);
    if (rc) return rc;
    x.blk = a.blk;
    x.blk1 = a.blk1;
    x.l1_any = a.l1_any;
    return launch_inverse(ctx, dev_coeffs_scratch, x, dev_pixels_out);
}

}  // extern "C"
