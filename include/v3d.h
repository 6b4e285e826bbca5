/*
 * v3d.h -- C ABI of libv3d.so, the B200 (sm_100a) implementation of the
 * video-3d-pipeline per-frame depth hot path.
 *
 * The reference (jabberjabberjabber/video-3d-pipeline) has NO native / FFI layer:
 * its hot path is Python calling cv2 (depth.py) and ffmpeg (upscale.py).  The
 * entry points below are therefore what a ctypes binding placed at the
 * reference's own call sites needs; each one cites the reference lines whose
 * arithmetic it replaces (paths under /root/reference/src/video_3d_pipeline/).
 * INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - extern "C", plain C types only.  Every function returns 0 on success and a
 *     negative V3D_E* code on failure; v3d_last_error() gives a thread-local
 *     message.  Nothing throws across the boundary.
 *   - All image pointers are CALLER-OWNED DEVICE pointers (e.g. torch CUDA
 *     tensors' data_ptr()) unless the function name ends in _host.  The library
 *     owns only the opaque context and its workspace.
 *   - Every call takes the cudaStream_t to run on (as void*; 0 = legacy default
 *     stream; pass torch.cuda.current_stream().cuda_stream) and is stream-ordered
 *     with no internal synchronisation, except v3d_depth_frames_host,
 *     v3d_host_wait and v3d_host_wait_oldest, which wait (sleeping, not spinning)
 *     for the outputs (so does a third *_host_async call while two are in flight).
 *   - A context belongs to one device and one host thread at a time.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with V3D_ECUDA.
 */
#ifndef V3D_H
#define V3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V3D_OK        0
#define V3D_EINVAL   (-1)  /* bad argument; mirrors the ValueError / cv2.error sites */
#define V3D_ENOMEM   (-2)
#define V3D_ECUDA    (-3)  /* CUDA runtime error or no device */
#define V3D_ESTATE   (-4)  /* call sequence error (e.g. tap before compute) */

#define V3D_INVALID_DISP (-16)   /* cv2: (minDisparity - 1) * 16; this is the value for minDisparity = 0 */

#define V3D_MODE_SGBM 0          /* 5 path directions (cv2.STEREO_SGBM_MODE_SGBM) */
#define V3D_MODE_HH   1          /* 8 path directions (cv2.STEREO_SGBM_MODE_HH)   */

/* cv2.StereoSGBM_create arguments, depth.py:315-325.  Field order is ABI. */
typedef struct v3d_sgbm_params {
    int32_t minDisparity;      /* 0 in the reference (depth.py:316); -1024 .. 1024 accepted, cv2 semantics */
    int32_t numDisparities;    /* positive multiple of 16, <= 256 (cv2's rule); kernels run at 64/128/256 */
    int32_t blockSize;         /* 5 (depth.py:318); 1,3,5,7 accepted */
    int32_t P1, P2;            /* depth.py:319-320 */
    int32_t disp12MaxDiff;     /* depth.py:321; <= 0 means 1, as in cv2 */
    int32_t preFilterCap;      /* cv2 default 0 -> ftzero 15 */
    int32_t uniquenessRatio;   /* depth.py:322; < 0 means 10, as in cv2 */
    int32_t speckleWindowSize; /* depth.py:323; 0 disables the speckle filter */
    int32_t speckleRange;      /* depth.py:324 */
    int32_t mode;              /* V3D_MODE_SGBM | V3D_MODE_HH */
} v3d_sgbm_params;

typedef struct v3d_ctx v3d_ctx;

/* Library identity / errors. */
const char* v3d_version(void);
const char* v3d_last_error(void);

/* Fill *p with the reference's literals (depth.py:315-325; mode = MODE_SGBM). */
void v3d_default_params(v3d_sgbm_params* p);

/* Create a context for eyes of eye_w x eye_h and up to max_batch frames per
 * call.  Fails with V3D_EINVAL where cv2 raises (eye_w - numDisparities <=
 * blockSize/2) and for unsupported parameters.  Allocates the workspace. */
int v3d_create(int device, const v3d_sgbm_params* params, int eye_w, int eye_h,
               int max_batch, v3d_ctx** out);
int v3d_destroy(v3d_ctx* ctx);
/* Device bytes held by the context (cost volumes, path state, labels ...). */
size_t v3d_workspace_bytes(const v3d_ctx* ctx);

/* depth.py:250-268 (split_sbs_frame, optional INTER_LANCZOS4 x2 unsqueeze) +
 * depth.py:274-275 (BGR2RGB) + depth.py:337-338 (RGB2GRAY), fused.
 * sbs_bgr: [batch][h][sbs_w][3] uint8, row pitch sbs_pitch bytes, frame stride
 * sbs_stride bytes.  left/right: [batch][h][eye_w] uint8 with gray_pitch /
 * gray_stride.  eye_w = sbs_w/2, or sbs_w when unsqueeze.  V3D_EINVAL on odd
 * sbs_w (depth.py:254-255). */
int v3d_split_gray(v3d_ctx* ctx, const uint8_t* sbs_bgr, size_t sbs_pitch, size_t sbs_stride,
                   int sbs_w, int h, int batch, int unsqueeze,
                   uint8_t* left_gray, uint8_t* right_gray, size_t gray_pitch, size_t gray_stride,
                   void* stream);

/* depth.py:274-275 + 337-338 for one already-split BGR eye (the
 * process_frame_batch entry, depth.py:297-341).  bgr: [batch][h][w][3]. */
int v3d_bgr_to_gray(v3d_ctx* ctx, const uint8_t* bgr, size_t pitch, size_t stride,
                    int w, int h, int batch, uint8_t* gray, size_t gray_pitch, size_t gray_stride,
                    void* stream);

/* The colour half of split_sbs_frame's unsqueeze, depth.py:263-266:
 * cv2.resize(eye, (2w, h), INTER_LANCZOS4) on a BGR eye [batch][h][w][3] ->
 * [batch][h][2w][3].  Context-free (no workspace needed). */
int v3d_unsqueeze_bgr(int device, const uint8_t* bgr, size_t pitch, size_t stride, int w, int h, int batch,
                      uint8_t* out, size_t out_pitch, size_t out_stride, void* stream);

/* depth.py:341 stereo.compute(left_gray, right_gray): prefilter, BT cost, box
 * sum, 5/8-path aggregation, WTA + uniqueness + sub-pixel + LR check, 3x3
 * median, speckle filter.  disp: [batch][eye_h][eye_w] int16, x16 fixed point,
 * invalid = V3D_INVALID_DISP; disp_pitch in BYTES. */
int v3d_sgbm_compute(v3d_ctx* ctx, const uint8_t* left_gray, const uint8_t* right_gray,
                     size_t gray_pitch, size_t gray_stride, int batch,
                     int16_t* disp, size_t disp_pitch, size_t disp_stride, void* stream);

/* Ask subsequent v3d_sgbm_compute calls to keep S_total and the pre-speckle
 * median in the workspace (taps 1 and 3 below).  Off by default: the product
 * path never writes S_total back to memory. */
int v3d_set_debug_taps(v3d_ctx* ctx, int enabled);

/* Parity-test taps into the workspace of the LAST v3d_sgbm_compute call.
 * which: 0 = block cost C  [batch][H][W1][Dk] uint16, Dk = numDisparities rounded up to 64/128/256,
 *            W1 = (W + min(minDisparity, 0)) - max(minDisparity + numDisparities, 0)
 *        1 = aggregated S  [batch][H][W1][Dk] uint16 (unsaturated sum); d >= numDisparities is padding
 *        2 = raw disparity (pre-median)  [batch][H][W] int16
 *        3 = post-median, pre-speckle    [batch][H][W] int16
 * Returns a device pointer valid until the next compute/destroy. */
int v3d_debug_tap(v3d_ctx* ctx, int which, void** dev_ptr, size_t* bytes);
/* Stream-ordered device-to-device copy of a tap into a caller-owned buffer. */
int v3d_debug_tap_copy(v3d_ctx* ctx, int which, void* dst_dev, size_t dst_bytes, void* stream);

/* depth.py:341 (.astype(float32)/16.0) + depth.py:374 (<=0 -> 0) into depth_f32
 * (may be NULL), and depth.py:400-403 (per-frame min-max -> uint16) into
 * depth_u16 (may be NULL).  Dense [batch][h][w] outputs. */
int v3d_postprocess(v3d_ctx* ctx, const int16_t* disp, size_t disp_pitch, size_t disp_stride,
                    int batch, float* depth_f32, uint16_t* depth_u16, void* stream);

/* save_depth_map's normalisation (depth.py:400-403) for an arbitrary float map:
 * depth_f32 [batch][n] -> out_u16 [batch][n], per-frame min-max, all-equal -> 0. */
int v3d_normalize_u16(v3d_ctx* ctx, const float* depth_f32, size_t n, int batch, uint16_t* out_u16, void* stream);

/* Guided upscale that replaces upscale.py:47-59 (ffmpeg scale): depth_u16
 * [batch][h][w] (dense) is bilinearly upsampled to gw x gh and colour-guided-
 * filtered (radius r, eps) with guide_rgb [batch][gh][gw][3] uint8 (dense).
 * out_u16: [batch][gh][gw].  q_f32 (may be NULL) receives the unquantised
 * result for tolerance tests.  Definition: oracle/guided.py. */
int v3d_guided_upscale(v3d_ctx* ctx, const uint16_t* depth_u16, int w, int h,
                       const uint8_t* guide_rgb, int gw, int gh, int batch,
                       int r, float eps, uint16_t* out_u16, float* q_f32, void* stream);

/* The whole device-resident frame path: v3d_split_gray -> v3d_sgbm_compute ->
 * v3d_postprocess [-> v3d_guided_upscale when guide_rgb != NULL].
 * Any of disp / depth_f32 / depth_u16 / out_4k may be NULL. */
int v3d_depth_frames(v3d_ctx* ctx, const uint8_t* sbs_bgr, size_t sbs_pitch, size_t sbs_stride,
                     int sbs_w, int h, int batch, int unsqueeze,
                     int16_t* disp, float* depth_f32, uint16_t* depth_u16,
                     const uint8_t* guide_rgb, int gw, int gh, int r, float eps,
                     uint16_t* out_4k, void* stream);

/* Same path from HOST buffers (dense arrays; pinned memory recommended): copies
 * the inputs host->device, runs, copies the requested outputs device->host and
 * waits for them (v3d_depth_frames_host_async + v3d_host_wait below).  This pair is
 * the call the end-to-end numbers time. */
int v3d_depth_frames_host(v3d_ctx* ctx, const uint8_t* sbs_bgr_host, int sbs_w, int h, int batch,
                          int unsqueeze, int16_t* disp_host, float* depth_f32_host,
                          uint16_t* depth_u16_host, const uint8_t* guide_rgb_host, int gw, int gh,
                          int r, float eps, uint16_t* out_4k_host, void* stream);

/* Asynchronous form of v3d_depth_frames_host (replaces the batch loop body of depth.py:448-461 for a
 * caller that keeps several batches in flight): enqueues the uploads -- one copy per frame on the
 * context's upload stream, SBS frames first so that the SGBM chain runs under the guide upload -- the
 * kernels on `stream`, and the downloads (one copy per frame, on the context's download stream), then
 * returns without waiting.
 * TWO calls may be in flight per context (each has its own staging buffers): the uploads of call k+1 run
 * under the kernels of call k, the 4K download of call k under the kernels of call k+1.  A third call
 * first waits (sleeping) for the oldest one.  The uploads start at once: the host input buffers must be
 * complete when the call is made.  `stream` waits for the kernels and for the downloads of disp /
 * depth_f32 / depth_u16; the download of out_4k completes on its own -- the host buffers of a call must
 * stay valid, and its outputs must not be read, until v3d_host_wait_oldest / v3d_host_wait has covered
 * it.  v3d_depth_frames_host = this + v3d_host_wait. */
int v3d_depth_frames_host_async(v3d_ctx* ctx, const uint8_t* sbs_bgr_host, int sbs_w, int h, int batch,
                                int unsqueeze, int16_t* disp_host, float* depth_f32_host,
                                uint16_t* depth_u16_host, const uint8_t* guide_rgb_host, int gw, int gh,
                                int r, float eps, uint16_t* out_4k_host, void* stream);
/* The upscale step alone from HOST buffers (replaces upscale.py:47-59 for one batch): depth_u16_host
 * [batch][eye_h][eye_w] (the context's eye size) + guide_rgb_host [batch][gh][gw][3] -> out_u16_host
 * [batch][gh][gw].  Asynchronous like the call above. */
int v3d_guided_upscale_host_async(v3d_ctx* ctx, const uint16_t* depth_u16_host, const uint8_t* guide_rgb_host,
                                  int gw, int gh, int batch, int r, float eps, uint16_t* out_u16_host,
                                  void* stream);
/* Block the calling host thread until EVERY *_host_async call of the context has delivered its outputs.
 * The wait sleeps on a blocking-sync CUDA event (no spinning), so one submit thread can drive many
 * contexts and many ranks can share few host cores. */
int v3d_host_wait(v3d_ctx* ctx);
/* The same for the OLDEST call in flight only (the pipelined loop: submit k+1, wait for k); no-op when
 * nothing is pending.  v3d_host_pending: calls in flight (0, 1 or 2). */
int v3d_host_wait_oldest(v3d_ctx* ctx);
int v3d_host_pending(const v3d_ctx* ctx);
/* Measurement aid: exactly the host<->device copies of v3d_depth_frames_host_async (same streams, same
 * per-frame granularity, same event dependencies) with NO kernel in between -- the ceiling the host side
 * of a box puts on the end-to-end number.  disp_host / out_4k_host receive whatever the workspace holds. */
int v3d_host_copy_only_async(v3d_ctx* ctx, const uint8_t* sbs_bgr_host, int sbs_w, int h, int batch,
                             int16_t* disp_host, const uint8_t* guide_rgb_host, int gw, int gh,
                             uint16_t* out_4k_host, void* stream);

/* How many frames the fused vertical sweep keeps co-resident on this device (one thread-block
 * cluster per frame); batches that are a multiple of it leave no partial wave.  0 until the first
 * v3d_sgbm_compute call, or when the shape uses the unfused path kernels. */
int v3d_fused_sweep_clusters(const v3d_ctx* ctx);
/* Number of kernel launches issued through this context so far. */
unsigned long long v3d_launch_count(const v3d_ctx* ctx);
/* GPU-side 16-bit PNG writer (cv2.imwrite of depth.py:406 costs ~50 ms per frame and core).
 * Packs uint16 gray images into complete IDAT payloads: a zlib stream of STORED deflate blocks (no
 * compression) holding the filter-0 scanlines with big-endian samples, Adler-32 included.  The host only
 * adds the fixed chunks and the IDAT CRC (video_3d_pipeline._native.png16_file_chunks) -- the decoded
 * pixels are identical to what cv2.imwrite stores.
 * payload: [batch][payload_stride] bytes, payload_stride >= v3d_png16_payload_bytes(w, h). */
size_t v3d_png16_payload_bytes(int w, int h);
int v3d_png16_pack(v3d_ctx* ctx, const uint16_t* img_u16, int w, int h, int batch, uint8_t* payload,
                   size_t payload_stride, void* stream);

/* OPT-IN behaviour change (the default reproduces the reference).
 * save_depth_map (depth.py:400-401) stretches every frame to its own min/max, so
 * the 16-bit depth scale flickers from frame to frame.  With fixed = 1 the uint16
 * maps of subsequent v3d_postprocess / v3d_depth_frames* calls use one scale for
 * the whole clip instead:  u16 = trunc(clip((d - lo) / (hi - lo), 0, 1) * 65535)
 * in fp32 with d in pixels.  fixed = 0 restores the per-frame min-max.
 * Returns V3D_EINVAL unless hi > lo. */
int v3d_set_depth_scale(v3d_ctx* ctx, int fixed, float lo, float hi);

/* Record per-stage CUDA-event timings for subsequent calls (0 = off).  With
 * timing on, v3d_stage_ms(ctx, i, &name) returns the accumulated milliseconds
 * of stage i (synchronises) or a negative value when i is out of range. */
int v3d_set_timing(v3d_ctx* ctx, int enabled);
double v3d_stage_ms(v3d_ctx* ctx, int stage, const char** name);
int v3d_reset_timing(v3d_ctx* ctx);

/* Measurement aid, not on the data path (the ALU roofline's denominator is the MEASURED integer
 * min/add rate of the device).  Runs a register-only kernel at full occupancy on `device` (synchronous,
 * a few milliseconds) and reports the sustained rate of one instruction mix of the path recurrence:
 *   kind 0  int32 add + min (one fused instruction);   kind 1  packed u16x2 add, then min (two);
 *   kind 2  fused packed u16x2 add-min (one DPX instruction).
 * lane_instr_per_s = thread-level instructions per second, algorithmic_ops_per_s = add/min operations
 * on cost cells per second (a packed instruction handles two cells). */
int v3d_probe_int_throughput(int device, int kind, double* lane_instr_per_s, double* algorithmic_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* V3D_H */
