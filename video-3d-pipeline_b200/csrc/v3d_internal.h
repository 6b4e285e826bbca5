// Internal declarations shared by the libv3d translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../include/v3d.h"

#define V3D_FULL_MASK 0xffffffffu

// Debug build (-DV3D_DEBUG_ASSERTS=1, `V3D_NVCC_EXTRA=-DV3D_DEBUG_ASSERTS=1 python build.py`): index and protocol
// checks in the code that synchronises without locks -- the neighbour inboxes / mbarrier ring of the fused vertical
// sweep and the union-find forest of the speckle filter -- trap instead of corrupting memory (compute-sanitizer is
// not usable on every GPU pool).  tools/gpu_debug_asserts.sh runs the parity tests against such a build.
#ifdef V3D_DEBUG_ASSERTS
#include <stdio.h>
#define V3D_DASSERT(c) do { if (!(c)) { printf("V3D_DASSERT failed: %s (%s:%d) block %d thread %d\n", #c, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define V3D_DASSERT(c) ((void)0)
#endif

enum V3dStage {
    ST_SPLIT_GRAY = 0, ST_PREFILTER, ST_COST, ST_PATHS, ST_LR, ST_WTA, ST_SELECT, ST_MEDIAN,
    ST_SPECKLE, ST_POST, ST_GUIDED, ST_GUIDED_APPLY, ST_COPY, ST_COUNT
};

struct V3dTimedSpan { int stage; cudaEvent_t a, b; };

struct v3d_ctx {
    int device;
    v3d_sgbm_params p;
    int W, H, D, W1, R, max_batch, ndirs;
    int minD, x0, inv;           // minDisparity; first image column of the window, max(minD + D, 0); invalid value (minD - 1) * 16
    int Dk;                      // disparities the kernels are instantiated for (64/128/256 >= D); d in [D, Dk) is padding
    int P1, P2, uniq, maxdiff, ftzero;

    // workspace (device)
    uint8_t *grayL, *grayR;      // [B][H][gpitch]
    size_t gpitch;
    uint4 *rexp, *lexp;          // expanded BT operands (k_cost.cu): right [B][H][2][2][rexp_wpw], left [B][H][W+16][2]
    int rexp_wpw;
    uint16_t *C, *S;             // [B][H][W1][D]
    uint16_t* ckpt;              // [B][H][ceil(W1/8)][D] left-to-right path state entering every 8-pixel chunk (k_paths_h.cu)
    cudaStream_t side_stream; cudaEvent_t ev_fork, ev_join;   // the checkpoint pass runs next to the vertical sweep
    uint2* rec;                  // WTA records [B][H][W1]
    int16_t *raw, *med, *disp;   // [B][H][W]
    int *labels, *sizes;         // [B][H*W]
    int* minmax;                 // [B][2]
    unsigned long long* png_sums; // [B][2] Adler-32 partial sums of v3d_png16_pack
    float* f32_tmp;              // [B][H][W]
    uint16_t* u16_tmp;           // [B][H][W]
    // lazily sized buffers
    float4* ab; size_t ab_bytes;           // guided coefficients [B][gh][gw]
    // host entry points (v3d_depth_frames_host_async): uploads and downloads run frame by frame on their own streams
    // so that both DMA engines stay busy next to the kernels; completion is a blocking-sync event (the waiting host
    // thread sleeps instead of spinning).  TWO calls may be in flight: each has its own staging buffers and events
    // (slot = call number & 1), so the uploads of call k+1 run under the kernels of call k and the 4K download of
    // call k under the kernels of call k+1.
    struct HostSlot {
        uint8_t* in_dev; size_t in_bytes;          // staged SBS frames
        uint8_t* guide_dev; size_t guide_bytes;    // staged guide frames
        uint16_t* out_dev; size_t out_bytes;       // upscaled output before its download
        cudaEvent_t ev_sbs, ev_guide, ev_compute, ev_done;
        int pending;                               // submitted and not waited for yet
        int used;                                  // ev_compute / ev_done have been recorded at least once
        unsigned long long call;                   // number of the call that occupies the slot
    } hs[2];
    cudaStream_t up_stream, down_stream;
    cudaEvent_t ev_entry, ev_small;                // caller's stream at entry; the small downloads (disp / f32 / u16) are done
    unsigned long long host_calls;                 // asynchronous host calls issued so far
    size_t bytes;
    int last_batch;

    unsigned long long launches;
    int timing;
    int debug_taps;          // keep S_total and the pre-speckle median for v3d_debug_tap
    int guided_attr_set;
    int fixed_scale; float scale_lo, scale_hi;     // v3d_set_depth_scale (0 = per-frame min-max, the reference)
    int max_clusters;        // co-resident frame clusters of the fused vertical sweep (0 = not queried)
    int v3_cl;               // CTAs per cluster the fused sweep runs with on this device (0 = not chosen yet)
    int no_fused_vertical;
    int h_attr_set;
    int select_attr_set;
    unsigned long long cost_attr_set;   // test hook: force the one-direction-per-launch path kernels
    std::vector<V3dTimedSpan> spans;
    double stage_ms[ST_COUNT];
};

// error plumbing (v3d_api.cu)
int v3d_fail(int code, const char* fmt, ...);
int v3d_cuda_check(cudaError_t e, const char* what);
#define V3D_CUDA(x) do { int _rc = v3d_cuda_check((x), #x); if (_rc) return _rc; } while (0)
#define V3D_LAUNCHED(ctx, n) do { (ctx)->launches += (n); int _rc = v3d_cuda_check(cudaGetLastError(), "kernel launch"); if (_rc) return _rc; } while (0)

struct V3dScope {   // per-stage CUDA-event timing when ctx->timing is on
    v3d_ctx* c; int idx; cudaStream_t s;
    V3dScope(v3d_ctx* ctx, int stage, cudaStream_t st);
    ~V3dScope();
};

// k_gray.cu
int v3d_launch_eyes_to_gray(v3d_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, size_t pitch,
                            size_t stride, int src_w, int h, int batch, int unsqueeze,
                            uint8_t* left_gray, uint8_t* right_gray, size_t gpitch, size_t gstride,
                            cudaStream_t st);
int v3d_launch_unsqueeze_bgr(const uint8_t* src, size_t pitch, size_t stride, int w, int h, int batch,
                             uint8_t* dst, size_t dpitch, size_t dstride, cudaStream_t st);
// k_cost.cu
int v3d_rexp_words(int W);
int v3d_lexp_cols(int W);
int v3d_launch_prefilter(v3d_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t gpitch,
                         size_t gstride, int batch, cudaStream_t st);
int v3d_launch_cost(v3d_ctx* ctx, int batch, cudaStream_t st);
// k_paths.cu
int v3d_launch_paths(v3d_ctx* ctx, int batch, cudaStream_t st);
int v3d_launch_path_lr(v3d_ctx* ctx, int batch, cudaStream_t st);
int v3d_launch_path_rl_wta(v3d_ctx* ctx, int batch, cudaStream_t st);
// k_post.cu
int v3d_launch_select(v3d_ctx* ctx, int batch, cudaStream_t st);
int v3d_launch_median(v3d_ctx* ctx, int batch, int16_t* dst, size_t dpitch, size_t dstride, cudaStream_t st);
int v3d_launch_speckle(v3d_ctx* ctx, int batch, int16_t* disp, size_t dpitch, size_t dstride, cudaStream_t st);
int v3d_launch_post(v3d_ctx* ctx, const int16_t* disp, size_t dpitch, size_t dstride, int batch,
                    float* f32, uint16_t* u16, cudaStream_t st);
int v3d_launch_normalize_f32(v3d_ctx* ctx, const float* in, size_t n, int batch, uint16_t* out, cudaStream_t st);
// k_png.cu
int v3d_launch_png16_pack(v3d_ctx* ctx, const uint16_t* img, int w, int h, int batch, uint8_t* payload,
                           size_t payload_stride, cudaStream_t st);
size_t v3d_png16_raw_bytes(int w, int h);

// k_guided.cu
int v3d_launch_guided(v3d_ctx* ctx, const uint16_t* depth, int w, int h, const uint8_t* guide, int gw, int gh,
                      int batch, int r, float eps, uint16_t* out, float* q, cudaStream_t st);
