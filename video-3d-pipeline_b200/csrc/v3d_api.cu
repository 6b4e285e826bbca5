// C ABI of libv3d.so (see include/v3d.h): context, workspace, argument checking, stage sequencing.
#include "v3d_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>

namespace {
thread_local char g_err[512] = "";
const char* const kStageNames[ST_COUNT] = { "split_gray", "prefilter", "cost", "vertical", "lr", "wta", "select",
                                            "median", "speckle", "post", "guided_coeff", "guided_apply", "copy" };
}  // namespace

int v3d_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int v3d_cuda_check(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return V3D_OK;
    return v3d_fail(V3D_ECUDA, "CUDA error in %s: %s", what, cudaGetErrorString(e));
}

V3dScope::V3dScope(v3d_ctx* ctx, int stage, cudaStream_t st) : c(ctx), idx(-1), s(st)
{
    if (c->timing <= 0) return;
    V3dTimedSpan sp;
    sp.stage = stage;
    if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) return;
    cudaEventRecord(sp.a, s);
    c->spans.push_back(sp);
    idx = (int)c->spans.size() - 1;
}

V3dScope::~V3dScope()
{
    if (idx >= 0) cudaEventRecord(c->spans[idx].b, s);
}

static void drain_spans(v3d_ctx* ctx)
{
    for (auto& sp : ctx->spans) {
        cudaEventSynchronize(sp.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) ctx->stage_ms[sp.stage] += ms;
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    ctx->spans.clear();
}

extern "C" {

const char* v3d_version(void) { return "libv3d 0.1.0 (sm_100a)"; }
const char* v3d_last_error(void) { return g_err; }

void v3d_default_params(v3d_sgbm_params* p)
{
    if (!p) return;
    // depth.py:315-325
    p->minDisparity = 0;
    p->numDisparities = 64;
    p->blockSize = 5;
    p->P1 = 8 * 3 * 5 * 5;
    p->P2 = 32 * 3 * 5 * 5;
    p->disp12MaxDiff = 1;
    p->preFilterCap = 0;
    p->uniquenessRatio = 10;
    p->speckleWindowSize = 100;
    p->speckleRange = 32;
    p->mode = V3D_MODE_SGBM;
}

static void free_all(v3d_ctx* c)
{
    void* ptrs[] = { c->grayL, c->grayR, c->rexp, c->lexp, c->C, c->S, c->ckpt, c->rec, c->raw, c->med, c->disp, c->labels,
                     c->sizes, c->minmax, c->png_sums, c->f32_tmp, c->u16_tmp, c->ab, c->hs[0].in_dev, c->hs[0].guide_dev,
                     c->hs[0].out_dev, c->hs[1].in_dev, c->hs[1].guide_dev, c->hs[1].out_dev };
    for (void* p : ptrs) if (p) cudaFree(p);
}

int v3d_create(int device, const v3d_sgbm_params* params, int eye_w, int eye_h, int max_batch, v3d_ctx** out)
{
    if (!params || !out) return v3d_fail(V3D_EINVAL, "null argument");
    *out = nullptr;
    const v3d_sgbm_params& p = *params;
    // (minDisparity + numDisparities) * 16 and (minDisparity - 1) * 16 must fit the int16 output
    if (p.minDisparity < -1024 || p.minDisparity > 1024)
        return v3d_fail(V3D_EINVAL, "minDisparity %d unsupported (-1024 .. 1024; depth.py:316 uses 0)", p.minDisparity);
    if (p.numDisparities < 16 || p.numDisparities > 256 || (p.numDisparities % 16))     // cv2: positive multiple of 16
        return v3d_fail(V3D_EINVAL, "numDisparities %d unsupported (multiples of 16 up to 256)", p.numDisparities);
    if (p.blockSize < 1 || !(p.blockSize & 1) || p.blockSize > 7)
        return v3d_fail(V3D_EINVAL, "blockSize %d unsupported (1, 3, 5, 7)", p.blockSize);
    if (p.mode != V3D_MODE_SGBM && p.mode != V3D_MODE_HH)
        return v3d_fail(V3D_EINVAL, "mode %d unsupported (0 = SGBM, 1 = HH)", p.mode);
    if (eye_w <= 0 || eye_h <= 0 || max_batch <= 0) return v3d_fail(V3D_EINVAL, "bad size");
    if (eye_w > 32768) return v3d_fail(V3D_EINVAL, "eye width %d too large (k_select keeps a row in shared memory: 6 bytes per column)", eye_w);
    if ((long long)max_batch * eye_w * eye_h >= (1ll << 31))
        return v3d_fail(V3D_EINVAL, "max_batch * width * height must stay below 2^31 (32-bit pixel labels)");
    // the window of image columns that have every disparity: [max(minD + D, 0), W + min(minD, 0))  (OpenCV minX1, maxX1)
    const int win_x0 = std::max(p.minDisparity + p.numDisparities, 0), win_x1 = eye_w + std::min(p.minDisparity, 0);
    if (win_x1 - win_x0 <= p.blockSize / 2)            // cv2.error in stereosgbm.cpp (for minDisparity = 0: W - D <= blockSize / 2)
        return v3d_fail(V3D_EINVAL, "eye width %d too small for numDisparities %d, minDisparity %d (cv2 raises here)", eye_w,
                        p.numDisparities, p.minDisparity);
    if (p.preFilterCap < 0 || p.preFilterCap > 63) return v3d_fail(V3D_EINVAL, "preFilterCap out of range");
    if (p.uniquenessRatio > 100) return v3d_fail(V3D_EINVAL, "uniquenessRatio out of range");
    const int P1 = p.P1 > 0 ? p.P1 : 2;
    const int P2 = std::max(p.P2 > 0 ? p.P2 : 5, P1 + 1);
    const int ndirs = p.mode == V3D_MODE_HH ? 8 : 5;
    const int ftzero = std::max(p.preFilterCap, 15) | 1;
    // worst-case block cost: blockSize^2 * (2*ftzero + 255/4); packed int16 state needs headroom
    const int cmax = p.blockSize * p.blockSize * (2 * ftzero + 63);
    if (cmax + P2 + P1 > 32000 || (long)ndirs * (cmax + P2) > 65535)
        return v3d_fail(V3D_EINVAL, "P1/P2 too large for the packed 16-bit path state (cmax %d)", cmax);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return v3d_fail(V3D_ECUDA, "no CUDA device: libv3d has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return v3d_fail(V3D_EINVAL, "device %d out of range (%d devices)", device, ndev);
    V3D_CUDA(cudaSetDevice(device));

    v3d_ctx* c = new (std::nothrow) v3d_ctx();
    if (!c) return v3d_fail(V3D_ENOMEM, "host allocation failed");
    c->device = device; c->p = p;
    c->W = eye_w; c->H = eye_h; c->D = p.numDisparities; c->W1 = win_x1 - win_x0; c->R = p.blockSize / 2;
    c->minD = p.minDisparity; c->x0 = win_x0; c->inv = (p.minDisparity - 1) * 16;
    c->Dk = c->D <= 64 ? 64 : (c->D <= 128 ? 128 : 256);
    c->max_batch = max_batch; c->ndirs = ndirs;
    c->P1 = P1; c->P2 = P2;
    c->uniq = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
    c->maxdiff = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
    c->ftzero = ftzero;
    c->gpitch = ((size_t)eye_w + 127) / 128 * 128;
    c->last_batch = 0;
    c->rexp_wpw = v3d_rexp_words(eye_w);

    const size_t B = (size_t)max_batch, npx = (size_t)eye_w * eye_h;
    const size_t vol = B * (size_t)eye_h * c->W1 * c->Dk * sizeof(uint16_t);
    struct { void** p; size_t n; } allocs[] = {
        { (void**)&c->grayL, B * c->gpitch * eye_h }, { (void**)&c->grayR, B * c->gpitch * eye_h },
        { (void**)&c->rexp, B * eye_h * 4 * (size_t)c->rexp_wpw * sizeof(uint4) },
        { (void**)&c->lexp, B * eye_h * (size_t)v3d_lexp_cols(eye_w) * 2 * sizeof(uint4) },
        { (void**)&c->C, vol }, { (void**)&c->S, vol },
        { (void**)&c->ckpt, B * (size_t)eye_h * ((c->W1 + 7) / 8) * c->Dk * sizeof(uint16_t) },
        { (void**)&c->rec, B * (size_t)eye_h * c->W1 * sizeof(uint2) },
        { (void**)&c->raw, B * npx * 2 }, { (void**)&c->med, B * npx * 2 }, { (void**)&c->disp, B * npx * 2 },
        { (void**)&c->labels, B * npx * 4 }, { (void**)&c->sizes, B * npx * 4 },
        { (void**)&c->minmax, B * 2 * sizeof(int) }, { (void**)&c->png_sums, B * 2 * sizeof(unsigned long long) },
        { (void**)&c->f32_tmp, B * npx * 4 }, { (void**)&c->u16_tmp, B * npx * 2 },
    };
    for (auto& a : allocs) {
        cudaError_t e = cudaMalloc(a.p, a.n);
        if (e != cudaSuccess) {
            cudaGetLastError();
            free_all(c);
            delete c;
            return v3d_fail(V3D_ENOMEM, "workspace allocation of %zu bytes failed: %s", a.n, cudaGetErrorString(e));
        }
        c->bytes += a.n;
    }
    // side stream of the path chain (k_paths.cu); V3D_NO_SIDE_STREAM=1 keeps the chain on one stream
    const char* ns = getenv("V3D_NO_SIDE_STREAM");
    if (!(ns && ns[0] == '1')) {
        // highest priority: the pass it carries is HBM-bound, so its blocks should take the next free block slot on any SM
        // (next to other lanes' integer-bound kernels) instead of queueing behind them; V3D_SIDE_PRIORITY=0 = default priority
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const char* sp = getenv("V3D_SIDE_PRIORITY");
        const int prio = (sp && sp[0] == '0') ? prio_lo : prio_hi;
        if (cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            c->side_stream = nullptr;
        }
    }
    *out = c;
    return V3D_OK;
}

int v3d_destroy(v3d_ctx* ctx)
{
    if (!ctx) return V3D_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    drain_spans(ctx);
    if (ctx->side_stream) { cudaStreamDestroy(ctx->side_stream); cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); }
    if (ctx->up_stream) {
        cudaStreamDestroy(ctx->up_stream); cudaStreamDestroy(ctx->down_stream);
        cudaEvent_t evs[] = { ctx->ev_entry, ctx->ev_small, ctx->hs[0].ev_sbs, ctx->hs[0].ev_guide, ctx->hs[0].ev_compute,
                              ctx->hs[0].ev_done, ctx->hs[1].ev_sbs, ctx->hs[1].ev_guide, ctx->hs[1].ev_compute, ctx->hs[1].ev_done };
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    }
    free_all(ctx);
    delete ctx;
    return V3D_OK;
}

size_t v3d_workspace_bytes(const v3d_ctx* ctx) { return ctx ? ctx->bytes : 0; }
unsigned long long v3d_launch_count(const v3d_ctx* ctx) { return ctx ? ctx->launches : 0; }

int v3d_fused_sweep_clusters(const v3d_ctx* ctx) { return ctx ? ctx->max_clusters : 0; }

int v3d_set_debug_taps(v3d_ctx* ctx, int enabled)
{
    if (!ctx) return v3d_fail(V3D_EINVAL, "null context");
    ctx->debug_taps = enabled != 0;
    return V3D_OK;
}

size_t v3d_png16_payload_bytes(int w, int h)
{
    if (w <= 0 || h <= 0) return 0;
    const size_t n = v3d_png16_raw_bytes(w, h);
    return 2 + 5 * ((n + 65534) / 65535) + n + 4;       // zlib header, stored-block headers, scanlines, Adler-32
}

int v3d_png16_pack(v3d_ctx* ctx, const uint16_t* img_u16, int w, int h, int batch, uint8_t* payload,
                   size_t payload_stride, void* stream)
{
    if (!ctx || !img_u16 || !payload) return v3d_fail(V3D_EINVAL, "null argument");
    if (w <= 0 || h <= 0 || batch <= 0 || batch > ctx->max_batch) return v3d_fail(V3D_EINVAL, "bad size / batch");
    if (v3d_png16_raw_bytes(w, h) >= (1ull << 32)) return v3d_fail(V3D_EINVAL, "image too large for one IDAT chunk");
    if (payload_stride < v3d_png16_payload_bytes(w, h)) return v3d_fail(V3D_EINVAL, "payload_stride too small");
    V3D_CUDA(cudaSetDevice(ctx->device));
    return v3d_launch_png16_pack(ctx, img_u16, w, h, batch, payload, payload_stride, (cudaStream_t)stream);
}

int v3d_set_depth_scale(v3d_ctx* ctx, int fixed, float lo, float hi)
{
    if (!ctx) return v3d_fail(V3D_EINVAL, "null context");
    if (fixed && !(hi > lo)) return v3d_fail(V3D_EINVAL, "fixed depth scale needs hi > lo (got %g, %g)", lo, hi);
    ctx->fixed_scale = fixed ? 1 : 0;
    ctx->scale_lo = lo;
    ctx->scale_hi = hi;
    return V3D_OK;
}

int v3d_set_timing(v3d_ctx* ctx, int enabled)
{
    if (!ctx) return v3d_fail(V3D_EINVAL, "null context");
    ctx->timing = enabled != 0;
    return V3D_OK;
}

int v3d_reset_timing(v3d_ctx* ctx)
{
    if (!ctx) return v3d_fail(V3D_EINVAL, "null context");
    drain_spans(ctx);
    for (double& m : ctx->stage_ms) m = 0.0;
    return V3D_OK;
}

double v3d_stage_ms(v3d_ctx* ctx, int stage, const char** name)
{
    if (!ctx || stage < 0 || stage >= ST_COUNT) return -1.0;
    drain_spans(ctx);
    if (name) *name = kStageNames[stage];
    return ctx->stage_ms[stage];
}

static int check_batch(v3d_ctx* ctx, int batch)
{
    if (!ctx) return v3d_fail(V3D_EINVAL, "null context");
    if (batch <= 0 || batch > ctx->max_batch)
        return v3d_fail(V3D_EINVAL, "batch %d outside 1..%d", batch, ctx->max_batch);
    V3D_CUDA(cudaSetDevice(ctx->device));
    return V3D_OK;
}

int v3d_split_gray(v3d_ctx* ctx, const uint8_t* sbs_bgr, size_t sbs_pitch, size_t sbs_stride, int sbs_w, int h,
                   int batch, int unsqueeze, uint8_t* left_gray, uint8_t* right_gray, size_t gray_pitch,
                   size_t gray_stride, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!sbs_bgr || !left_gray || !right_gray) return v3d_fail(V3D_EINVAL, "null buffer");
    if (sbs_w <= 0 || (sbs_w & 1)) return v3d_fail(V3D_EINVAL, "SBS frame width must be even");   // depth.py:254-255
    const int half = sbs_w / 2;
    const int eye_w = unsqueeze ? 2 * half : half;
    if (h <= 0 || gray_pitch < (size_t)eye_w || sbs_pitch < (size_t)sbs_w * 3) return v3d_fail(V3D_EINVAL, "bad pitch");
    return v3d_launch_eyes_to_gray(ctx, sbs_bgr, sbs_bgr + (size_t)half * 3, sbs_pitch, sbs_stride, half, h, batch,
                                   unsqueeze, left_gray, right_gray, gray_pitch, gray_stride, (cudaStream_t)stream);
}

int v3d_bgr_to_gray(v3d_ctx* ctx, const uint8_t* bgr, size_t pitch, size_t stride, int w, int h, int batch,
                    uint8_t* gray, size_t gray_pitch, size_t gray_stride, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!bgr || !gray) return v3d_fail(V3D_EINVAL, "null buffer");
    if (w <= 0 || h <= 0 || gray_pitch < (size_t)w || pitch < (size_t)w * 3) return v3d_fail(V3D_EINVAL, "bad pitch");
    return v3d_launch_eyes_to_gray(ctx, bgr, nullptr, pitch, stride, w, h, batch, 0, gray, nullptr, gray_pitch,
                                   gray_stride, (cudaStream_t)stream);
}

int v3d_unsqueeze_bgr(int device, const uint8_t* bgr, size_t pitch, size_t stride, int w, int h, int batch,
                      uint8_t* out, size_t out_pitch, size_t out_stride, void* stream)
{
    if (!bgr || !out) return v3d_fail(V3D_EINVAL, "null buffer");
    if (w <= 0 || h <= 0 || batch <= 0 || pitch < (size_t)w * 3 || out_pitch < (size_t)w * 6)
        return v3d_fail(V3D_EINVAL, "bad size or pitch");
    V3D_CUDA(cudaSetDevice(device));
    return v3d_launch_unsqueeze_bgr(bgr, pitch, stride, w, h, batch, out, out_pitch, out_stride, (cudaStream_t)stream);
}

int v3d_normalize_u16(v3d_ctx* ctx, const float* depth_f32, size_t n, int batch, uint16_t* out_u16, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!depth_f32 || !out_u16 || n == 0) return v3d_fail(V3D_EINVAL, "null buffer");
    return v3d_launch_normalize_f32(ctx, depth_f32, n, batch, out_u16, (cudaStream_t)stream);
}

int v3d_sgbm_compute(v3d_ctx* ctx, const uint8_t* left_gray, const uint8_t* right_gray, size_t gray_pitch,
                     size_t gray_stride, int batch, int16_t* disp, size_t disp_pitch, size_t disp_stride,
                     void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!left_gray || !right_gray || !disp) return v3d_fail(V3D_EINVAL, "null buffer");
    if (gray_pitch < (size_t)ctx->W || disp_pitch < (size_t)ctx->W * 2 || (disp_pitch & 1) || (disp_stride & 1))
        return v3d_fail(V3D_EINVAL, "bad pitch");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = v3d_launch_prefilter(ctx, left_gray, right_gray, gray_pitch, gray_stride, batch, st))) return rc;
    if ((rc = v3d_launch_cost(ctx, batch, st))) return rc;
    if ((rc = v3d_launch_paths(ctx, batch, st))) return rc;
    if ((rc = v3d_launch_select(ctx, batch, st))) return rc;
    if (ctx->debug_taps) {
        // keep the pre-speckle median for tap 3
        if ((rc = v3d_launch_median(ctx, batch, ctx->med, (size_t)ctx->W * 2, (size_t)ctx->W * ctx->H * 2, st))) return rc;
    }
    if ((rc = v3d_launch_median(ctx, batch, disp, disp_pitch, disp_stride, st))) return rc;
    if ((rc = v3d_launch_speckle(ctx, batch, disp, disp_pitch, disp_stride, st))) return rc;
    ctx->last_batch = batch;
    return V3D_OK;
}

int v3d_debug_tap(v3d_ctx* ctx, int which, void** dev_ptr, size_t* bytes)
{
    if (!ctx || !dev_ptr || !bytes) return v3d_fail(V3D_EINVAL, "null argument");
    if (ctx->last_batch <= 0) return v3d_fail(V3D_ESTATE, "no v3d_sgbm_compute call yet");
    if (!ctx->debug_taps && (which == 1 || which == 3))
        return v3d_fail(V3D_ESTATE, "tap %d needs v3d_set_debug_taps(ctx, 1) before the compute call", which);
    const size_t B = (size_t)ctx->last_batch;
    const size_t vol = B * (size_t)ctx->H * ctx->W1 * ctx->Dk * sizeof(uint16_t);
    const size_t img = B * (size_t)ctx->W * ctx->H * sizeof(int16_t);
    switch (which) {
        case 0: *dev_ptr = ctx->C; *bytes = vol; return V3D_OK;
        case 1: *dev_ptr = ctx->S; *bytes = vol; return V3D_OK;
        case 2: *dev_ptr = ctx->raw; *bytes = img; return V3D_OK;
        case 3: *dev_ptr = ctx->med; *bytes = img; return V3D_OK;
    }
    return v3d_fail(V3D_EINVAL, "unknown tap %d", which);
}

int v3d_debug_tap_copy(v3d_ctx* ctx, int which, void* dst_dev, size_t dst_bytes, void* stream)
{
    void* src = nullptr;
    size_t n = 0;
    if (int rc = v3d_debug_tap(ctx, which, &src, &n)) return rc;
    if (!dst_dev || dst_bytes < n) return v3d_fail(V3D_EINVAL, "tap %d needs %zu bytes", which, n);
    V3D_CUDA(cudaMemcpyAsync(dst_dev, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return V3D_OK;
}

int v3d_postprocess(v3d_ctx* ctx, const int16_t* disp, size_t disp_pitch, size_t disp_stride, int batch,
                    float* depth_f32, uint16_t* depth_u16, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!disp) return v3d_fail(V3D_EINVAL, "null buffer");
    if (disp_pitch < (size_t)ctx->W * 2 || (disp_pitch & 1) || (disp_stride & 1)) return v3d_fail(V3D_EINVAL, "bad pitch");
    if (!depth_f32 && !depth_u16) return V3D_OK;
    return v3d_launch_post(ctx, disp, disp_pitch, disp_stride, batch, depth_f32, depth_u16, (cudaStream_t)stream);
}

int v3d_guided_upscale(v3d_ctx* ctx, const uint16_t* depth_u16, int w, int h, const uint8_t* guide_rgb, int gw,
                       int gh, int batch, int r, float eps, uint16_t* out_u16, float* q_f32, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!depth_u16 || !guide_rgb || !out_u16) return v3d_fail(V3D_EINVAL, "null buffer");
    if (w <= 0 || h <= 0 || gw <= 0 || gh <= 0 || gw > 32768 || gh > 32768 || w > 32768 || h > 32768)
        return v3d_fail(V3D_EINVAL, "bad size");
    if (!(eps > 0.f)) return v3d_fail(V3D_EINVAL, "eps must be positive");
    return v3d_launch_guided(ctx, depth_u16, w, h, guide_rgb, gw, gh, batch, r, eps, out_u16, q_f32,
                             (cudaStream_t)stream);
}

int v3d_depth_frames(v3d_ctx* ctx, const uint8_t* sbs_bgr, size_t sbs_pitch, size_t sbs_stride, int sbs_w, int h,
                     int batch, int unsqueeze, int16_t* disp, float* depth_f32, uint16_t* depth_u16,
                     const uint8_t* guide_rgb, int gw, int gh, int r, float eps, uint16_t* out_4k, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (sbs_w <= 0 || (sbs_w & 1)) return v3d_fail(V3D_EINVAL, "SBS frame width must be even");
    const int eye_w = unsqueeze ? sbs_w : sbs_w / 2;
    if (eye_w != ctx->W || h != ctx->H)
        return v3d_fail(V3D_EINVAL, "frame gives %dx%d eyes, context was created for %dx%d", eye_w, h, ctx->W, ctx->H);
    if (guide_rgb && !out_4k) return v3d_fail(V3D_EINVAL, "guide given without an output buffer");
    const size_t gstride = ctx->gpitch * ctx->H;
    const size_t dpitch = (size_t)ctx->W * 2, dstride = dpitch * ctx->H;
    int16_t* d = disp ? disp : ctx->disp;
    int rc;
    if ((rc = v3d_split_gray(ctx, sbs_bgr, sbs_pitch, sbs_stride, sbs_w, h, batch, unsqueeze, ctx->grayL, ctx->grayR,
                             ctx->gpitch, gstride, stream))) return rc;
    if ((rc = v3d_sgbm_compute(ctx, ctx->grayL, ctx->grayR, ctx->gpitch, gstride, batch, d, dpitch, dstride, stream)))
        return rc;
    uint16_t* u16 = depth_u16 ? depth_u16 : (guide_rgb ? ctx->u16_tmp : nullptr);
    if ((rc = v3d_postprocess(ctx, d, dpitch, dstride, batch, depth_f32, u16, stream))) return rc;
    if (guide_rgb)
        if ((rc = v3d_guided_upscale(ctx, u16, ctx->W, ctx->H, guide_rgb, gw, gh, batch, r, eps, out_4k, nullptr,
                                     stream))) return rc;
    return V3D_OK;
}

static int ensure(v3d_ctx* ctx, void** p, size_t* have, size_t need)
{
    if (*have >= need) return V3D_OK;
    if (*p) { V3D_CUDA(cudaDeviceSynchronize()); V3D_CUDA(cudaFree(*p)); ctx->bytes -= *have; *p = nullptr; *have = 0; }
    V3D_CUDA(cudaMalloc(p, need));
    *have = need; ctx->bytes += need;
    return V3D_OK;
}

}  // extern "C"

namespace {

// One asynchronous host call: which host buffers go in and come out.
struct HostJob {
    const uint8_t* sbs = nullptr; int sbs_w = 0, h = 0, unsqueeze = 0;      // SBS frames in (NULL: upscale only)
    const uint16_t* depth_in = nullptr;                                     // uint16 depth in (upscale only)
    const uint8_t* guide = nullptr; int gw = 0, gh = 0, r = 0; float eps = 0.f;
    int16_t* disp = nullptr; float* f32 = nullptr; uint16_t* u16 = nullptr; uint16_t* out4k = nullptr;
    int batch = 0;
    bool copy_only = false;                                                 // same copies, no kernels
};

int host_streams(v3d_ctx* ctx)
{
    if (ctx->up_stream) return V3D_OK;
    V3D_CUDA(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
    V3D_CUDA(cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    cudaEvent_t* evs[] = { &ctx->ev_entry, &ctx->ev_small, &ctx->hs[0].ev_sbs, &ctx->hs[0].ev_guide, &ctx->hs[0].ev_compute,
                           &ctx->hs[1].ev_sbs, &ctx->hs[1].ev_guide, &ctx->hs[1].ev_compute };
    for (cudaEvent_t* e : evs) V3D_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    // the host waits on these: blocking sync = the thread sleeps on an OS primitive instead of spinning, so
    // many lanes and ranks can share few host cores
    for (auto& sl : ctx->hs) V3D_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming | cudaEventBlockingSync));
    return V3D_OK;
}

// wait (sleeping) for the call that occupies a slot
int slot_wait(v3d_ctx* ctx, int s)
{
    auto& sl = ctx->hs[s];
    if (!sl.pending) return V3D_OK;
    sl.pending = 0;
    V3D_CUDA(cudaEventSynchronize(sl.ev_done));
    return V3D_OK;
}

// Enqueue one host call.  Three streams: uploads (frame by frame, SBS first: the SGBM chain starts as soon as the
// SBS frames are in and runs under the guide upload), kernels on the caller's stream, downloads (frame by frame).
// Two calls may be in flight; call k uses slot k & 1 (its own staging buffers and events), so
//   * the uploads of call k+1 only wait for the kernels of call k-1 (the previous readers of the slot's buffers) and
//     run under the kernels of call k;
//   * the caller's stream waits for the downloads of disp / depth_f32 / depth_u16 (they read workspace buffers the
//     next call's kernels overwrite) but not for the 4K download, which runs under the next call's kernels.
int host_submit(v3d_ctx* ctx, const HostJob& j, cudaStream_t st)
{
    int rc;
    const int B = j.batch;
    const size_t npx = (size_t)ctx->W * ctx->H;
    const size_t sbs_frame = (size_t)j.sbs_w * j.h * 3;
    const size_t guide_frame = (size_t)j.gw * j.gh * 3, out_frame = (size_t)j.gw * j.gh * 2;
    if ((rc = host_streams(ctx))) return rc;
    const int s = (int)(ctx->host_calls & 1);
    auto& sl = ctx->hs[s];
    if ((rc = slot_wait(ctx, s))) return rc;       // a third call in flight first waits for the oldest one
    if (j.sbs && (rc = ensure(ctx, (void**)&sl.in_dev, &sl.in_bytes, sbs_frame * B))) return rc;
    if (j.guide) {
        if ((rc = ensure(ctx, (void**)&sl.guide_dev, &sl.guide_bytes, guide_frame * B))) return rc;
        if ((rc = ensure(ctx, (void**)&sl.out_dev, &sl.out_bytes, out_frame * B))) return rc;
    }
    cudaStream_t up = ctx->up_stream, down = ctx->down_stream;
    if (sl.used) V3D_CUDA(cudaStreamWaitEvent(up, sl.ev_compute, 0));     // the slot's staged inputs have been consumed
    if (j.depth_in) {
        // the depth maps go straight into a workspace buffer that earlier work on the caller's stream may still read
        V3D_CUDA(cudaEventRecord(ctx->ev_entry, st));
        V3D_CUDA(cudaStreamWaitEvent(up, ctx->ev_entry, 0));
    }
    {
        V3dScope scope(ctx, ST_COPY, up);
        if (j.sbs)
            for (int f = 0; f < B; f++)
                V3D_CUDA(cudaMemcpyAsync(sl.in_dev + f * sbs_frame, j.sbs + f * sbs_frame, sbs_frame, cudaMemcpyHostToDevice, up));
        if (j.depth_in)
            V3D_CUDA(cudaMemcpyAsync(ctx->u16_tmp, j.depth_in, npx * 2 * B, cudaMemcpyHostToDevice, up));
        V3D_CUDA(cudaEventRecord(sl.ev_sbs, up));
        if (j.guide) {
            for (int f = 0; f < B; f++)
                V3D_CUDA(cudaMemcpyAsync(sl.guide_dev + f * guide_frame, j.guide + f * guide_frame, guide_frame, cudaMemcpyHostToDevice, up));
            V3D_CUDA(cudaEventRecord(sl.ev_guide, up));
        }
    }
    V3D_CUDA(cudaStreamWaitEvent(st, sl.ev_sbs, 0));
    if (j.sbs && !j.copy_only) {
        rc = v3d_depth_frames(ctx, sl.in_dev, (size_t)j.sbs_w * 3, sbs_frame, j.sbs_w, j.h, B, j.unsqueeze, ctx->disp,
                              j.f32 ? ctx->f32_tmp : nullptr, j.u16 || j.guide ? ctx->u16_tmp : nullptr,
                              nullptr, 0, 0, j.r, j.eps, nullptr, st);
        if (rc) return rc;
    }
    if (j.guide) {
        V3D_CUDA(cudaStreamWaitEvent(st, sl.ev_guide, 0));
        if (sl.used) V3D_CUDA(cudaStreamWaitEvent(st, sl.ev_done, 0));    // the slot's previous output has left the device
        if (!j.copy_only)
            if ((rc = v3d_guided_upscale(ctx, ctx->u16_tmp, ctx->W, ctx->H, sl.guide_dev, j.gw, j.gh, B, j.r, j.eps,
                                         sl.out_dev, nullptr, st))) return rc;
    }
    V3D_CUDA(cudaEventRecord(sl.ev_compute, st));
    V3D_CUDA(cudaStreamWaitEvent(down, sl.ev_compute, 0));
    {
        V3dScope scope(ctx, ST_COPY, down);
        if (j.disp) V3D_CUDA(cudaMemcpyAsync(j.disp, ctx->disp, npx * 2 * B, cudaMemcpyDeviceToHost, down));
        if (j.f32) V3D_CUDA(cudaMemcpyAsync(j.f32, ctx->f32_tmp, npx * 4 * B, cudaMemcpyDeviceToHost, down));
        if (j.u16) V3D_CUDA(cudaMemcpyAsync(j.u16, ctx->u16_tmp, npx * 2 * B, cudaMemcpyDeviceToHost, down));
        V3D_CUDA(cudaEventRecord(ctx->ev_small, down));
        if (j.out4k)
            for (int f = 0; f < B; f++)
                V3D_CUDA(cudaMemcpyAsync(j.out4k + f * (out_frame / 2), sl.out_dev + f * (out_frame / 2), out_frame,
                                         cudaMemcpyDeviceToHost, down));
    }
    V3D_CUDA(cudaEventRecord(sl.ev_done, down));
    V3D_CUDA(cudaStreamWaitEvent(st, ctx->ev_small, 0));
    sl.pending = 1;
    sl.used = 1;
    sl.call = ctx->host_calls++;
    return V3D_OK;
}

int check_frames_host_args(v3d_ctx* ctx, const uint8_t* sbs, int sbs_w, int h, int batch, int unsqueeze,
                           const uint8_t* guide, int gw, int gh, const uint16_t* out4k)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!sbs) return v3d_fail(V3D_EINVAL, "null buffer");
    if (sbs_w <= 0 || (sbs_w & 1)) return v3d_fail(V3D_EINVAL, "SBS frame width must be even");
    const int eye_w = unsqueeze ? sbs_w : sbs_w / 2;
    if (eye_w != ctx->W || h != ctx->H)
        return v3d_fail(V3D_EINVAL, "frame gives %dx%d eyes, context was created for %dx%d", eye_w, h, ctx->W, ctx->H);
    if (guide && !out4k) return v3d_fail(V3D_EINVAL, "guide given without an output buffer");
    if (guide && (gw <= 0 || gh <= 0 || gw > 32768 || gh > 32768)) return v3d_fail(V3D_EINVAL, "bad guide size");
    return V3D_OK;
}

}  // namespace

extern "C" {

int v3d_depth_frames_host_async(v3d_ctx* ctx, const uint8_t* sbs_bgr_host, int sbs_w, int h, int batch, int unsqueeze,
                                int16_t* disp_host, float* depth_f32_host, uint16_t* depth_u16_host,
                                const uint8_t* guide_rgb_host, int gw, int gh, int r, float eps, uint16_t* out_4k_host,
                                void* stream)
{
    if (int rc = check_frames_host_args(ctx, sbs_bgr_host, sbs_w, h, batch, unsqueeze, guide_rgb_host, gw, gh, out_4k_host))
        return rc;
    if (guide_rgb_host && !(eps > 0.f)) return v3d_fail(V3D_EINVAL, "eps must be positive");
    HostJob j;
    j.sbs = sbs_bgr_host; j.sbs_w = sbs_w; j.h = h; j.unsqueeze = unsqueeze; j.batch = batch;
    j.disp = disp_host; j.f32 = depth_f32_host; j.u16 = depth_u16_host;
    j.guide = guide_rgb_host; j.gw = gw; j.gh = gh; j.r = r; j.eps = eps; j.out4k = out_4k_host;
    return host_submit(ctx, j, (cudaStream_t)stream);
}

int v3d_guided_upscale_host_async(v3d_ctx* ctx, const uint16_t* depth_u16_host, const uint8_t* guide_rgb_host, int gw,
                                  int gh, int batch, int r, float eps, uint16_t* out_u16_host, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!depth_u16_host || !guide_rgb_host || !out_u16_host) return v3d_fail(V3D_EINVAL, "null buffer");
    if (gw <= 0 || gh <= 0 || gw > 32768 || gh > 32768) return v3d_fail(V3D_EINVAL, "bad guide size");
    if (!(eps > 0.f)) return v3d_fail(V3D_EINVAL, "eps must be positive");
    HostJob j;
    j.depth_in = depth_u16_host; j.batch = batch;
    j.guide = guide_rgb_host; j.gw = gw; j.gh = gh; j.r = r; j.eps = eps; j.out4k = out_u16_host;
    return host_submit(ctx, j, (cudaStream_t)stream);
}

int v3d_host_copy_only_async(v3d_ctx* ctx, const uint8_t* sbs_bgr_host, int sbs_w, int h, int batch, int16_t* disp_host,
                             const uint8_t* guide_rgb_host, int gw, int gh, uint16_t* out_4k_host, void* stream)
{
    if (int rc = check_batch(ctx, batch)) return rc;
    if (!sbs_bgr_host && !guide_rgb_host) return v3d_fail(V3D_EINVAL, "null buffer");
    if (guide_rgb_host && !out_4k_host) return v3d_fail(V3D_EINVAL, "guide given without an output buffer");
    HostJob j;
    j.sbs = sbs_bgr_host; j.sbs_w = sbs_w; j.h = h; j.batch = batch; j.disp = disp_host;
    j.guide = guide_rgb_host; j.gw = gw; j.gh = gh; j.out4k = out_4k_host;
    j.copy_only = true;
    return host_submit(ctx, j, (cudaStream_t)stream);
}

int v3d_host_wait(v3d_ctx* ctx)
{
    if (!ctx) return v3d_fail(V3D_EINVAL, "null context");
    if (!ctx->hs[0].pending && !ctx->hs[1].pending) return V3D_OK;
    V3D_CUDA(cudaSetDevice(ctx->device));
    const int first = (ctx->hs[0].pending && ctx->hs[1].pending && ctx->hs[1].call < ctx->hs[0].call) ? 1 : 0;
    if (int rc = slot_wait(ctx, first)) return rc;
    return slot_wait(ctx, first ^ 1);
}

int v3d_host_wait_oldest(v3d_ctx* ctx)
{
    if (!ctx) return v3d_fail(V3D_EINVAL, "null context");
    const bool p0 = ctx->hs[0].pending != 0, p1 = ctx->hs[1].pending != 0;
    if (!p0 && !p1) return V3D_OK;
    V3D_CUDA(cudaSetDevice(ctx->device));
    const int s = (p0 && p1) ? (ctx->hs[1].call < ctx->hs[0].call ? 1 : 0) : (p0 ? 0 : 1);
    return slot_wait(ctx, s);
}

int v3d_host_pending(const v3d_ctx* ctx) { return ctx ? (ctx->hs[0].pending != 0) + (ctx->hs[1].pending != 0) : 0; }

int v3d_depth_frames_host(v3d_ctx* ctx, const uint8_t* sbs_bgr_host, int sbs_w, int h, int batch, int unsqueeze,
                          int16_t* disp_host, float* depth_f32_host, uint16_t* depth_u16_host,
                          const uint8_t* guide_rgb_host, int gw, int gh, int r, float eps, uint16_t* out_4k_host,
                          void* stream)
{
    if (int rc = v3d_depth_frames_host_async(ctx, sbs_bgr_host, sbs_w, h, batch, unsqueeze, disp_host, depth_f32_host,
                                             depth_u16_host, guide_rgb_host, gw, gh, r, eps, out_4k_host, stream))
        return rc;
    return v3d_host_wait(ctx);
}

}  // extern "C"
