// SBS split + (optional Lanczos4 x2 unsqueeze) + BGR->gray, fused.
// Replaces depth.py:250-268 (split_sbs_frame), :274-275 (BGR2RGB), :337-338 (RGB2GRAY).
#include "v3d_internal.h"

namespace {

__device__ __forceinline__ uint32_t gray15(uint32_t b, uint32_t g, uint32_t r)
{
    // cv2 RGB2GRAY, 15-bit fixed point (exact over all 2^24 colours)
    return (9798u * r + 19235u * g + 3735u * b + 16384u) >> 15;
}

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[12], int i)
{
    return (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
}

// No unsqueeze: one thread converts 16 pixels = three 128-bit loads -> one 128-bit store.
// grid: (ceil(w/16/128), h, batch*eyes); blockIdx.z % eyes selects the eye.
__global__ void __launch_bounds__(128)
k_bgr2gray16(const uint8_t* __restrict__ srcL, const uint8_t* __restrict__ srcR, size_t pitch, size_t stride,
             int w, uint8_t* __restrict__ dstL, uint8_t* __restrict__ dstR, size_t gpitch, size_t gstride, int eyes)
{
    const int eye = blockIdx.z % eyes, b = blockIdx.z / eyes, y = blockIdx.y;
    const int px0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (px0 >= w) return;
    const uint8_t* s = (eye ? srcR : srcL) + (size_t)b * stride + (size_t)y * pitch + (size_t)px0 * 3;
    uint8_t* d = (eye ? dstR : dstL) + (size_t)b * gstride + (size_t)y * gpitch + px0;
    const bool vec = (px0 + 16 <= w) && ((reinterpret_cast<uintptr_t>(s) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(d) & 15) == 0);
    if (vec) {
        uint32_t wv[12];
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            uint4 v = __ldg(s4 + i);
            wv[4 * i] = v.x; wv[4 * i + 1] = v.y; wv[4 * i + 2] = v.z; wv[4 * i + 3] = v.w;
        }
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = (q * 4 + j) * 3;
                acc |= gray15(byte_of(wv, i), byte_of(wv, i + 1), byte_of(wv, i + 2)) << (8 * j);
            }
            o[q] = acc;
        }
        *reinterpret_cast<uint4*>(d) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        const int n = min(16, w - px0);
        for (int i = 0; i < n; i++) d[i] = (uint8_t)gray15(s[3 * i], s[3 * i + 1], s[3 * i + 2]);
    }
}

// Unsqueeze: cv2.resize(eye, (2w, h), INTER_LANCZOS4) horizontally, then gray.
// One thread per SOURCE column k produces destination columns 2k and 2k+1:
//   dst 2k   : taps T[0..7]   over src[k-4 .. k+3]
//   dst 2k+1 : taps T[7..0]   over src[k-3 .. k+4]        (replicate border)
__global__ void __launch_bounds__(128)
k_unsqueeze_gray(const uint8_t* __restrict__ srcL, const uint8_t* __restrict__ srcR, size_t pitch, size_t stride,
                 int w, uint8_t* __restrict__ dstL, uint8_t* __restrict__ dstR, size_t gpitch, size_t gstride, int eyes)
{
    const int eye = blockIdx.z % eyes, b = blockIdx.z / eyes, y = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= w) return;
    const uint8_t* s = (eye ? srcR : srcL) + (size_t)b * stride + (size_t)y * pitch;
    uint8_t* d = (eye ? dstR : dstL) + (size_t)b * gstride + (size_t)y * gpitch;
    const int T[8] = { -8, 64, -188, 579, 1830, -312, 114, -31 };
    uint32_t ev[3], od[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int v[9];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int xs = min(max(k - 4 + i, 0), w - 1);
            v[i] = __ldg(s + (size_t)xs * 3 + c);
        }
        int ae = 0, ao = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { ae += T[i] * v[i]; ao += T[7 - i] * v[i + 1]; }
        ev[c] = (uint32_t)min(max((ae + 1024) >> 11, 0), 255);
        od[c] = (uint32_t)min(max((ao + 1024) >> 11, 0), 255);
    }
    d[2 * k] = (uint8_t)gray15(ev[0], ev[1], ev[2]);
    d[2 * k + 1] = (uint8_t)gray15(od[0], od[1], od[2]);
}

// The same resize kept in colour: split_sbs_frame's own return value (depth.py:263-266).
// One thread per source column and channel triple; dst is [h][2w][3].
__global__ void __launch_bounds__(128)
k_unsqueeze_bgr(const uint8_t* __restrict__ src, size_t pitch, size_t stride, int w,
                uint8_t* __restrict__ dst, size_t dpitch, size_t dstride)
{
    const int b = blockIdx.z, y = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= w) return;
    const uint8_t* s = src + (size_t)b * stride + (size_t)y * pitch;
    uint8_t* d = dst + (size_t)b * dstride + (size_t)y * dpitch;
    const int T[8] = { -8, 64, -188, 579, 1830, -312, 114, -31 };
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int v[9];
#pragma unroll
        for (int i = 0; i < 9; i++) v[i] = __ldg(s + (size_t)min(max(k - 4 + i, 0), w - 1) * 3 + c);
        int ae = 0, ao = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { ae += T[i] * v[i]; ao += T[7 - i] * v[i + 1]; }
        d[(size_t)(2 * k) * 3 + c] = (uint8_t)min(max((ae + 1024) >> 11, 0), 255);
        d[(size_t)(2 * k + 1) * 3 + c] = (uint8_t)min(max((ao + 1024) >> 11, 0), 255);
    }
}

}  // namespace

int v3d_launch_unsqueeze_bgr(const uint8_t* src, size_t pitch, size_t stride, int w, int h, int batch,
                             uint8_t* dst, size_t dpitch, size_t dstride, cudaStream_t st)
{
    dim3 grid((w + 127) / 128, h, batch);
    k_unsqueeze_bgr<<<grid, 128, 0, st>>>(src, pitch, stride, w, dst, dpitch, dstride);
    return v3d_cuda_check(cudaGetLastError(), "k_unsqueeze_bgr");
}

int v3d_launch_eyes_to_gray(v3d_ctx* ctx, const uint8_t* left_bgr, const uint8_t* right_bgr, size_t pitch,
                            size_t stride, int src_w, int h, int batch, int unsqueeze,
                            uint8_t* left_gray, uint8_t* right_gray, size_t gpitch, size_t gstride,
                            cudaStream_t st)
{
    V3dScope scope(ctx, ST_SPLIT_GRAY, st);
    const int eyes = right_bgr ? 2 : 1;
    if (unsqueeze) {
        dim3 grid((src_w + 127) / 128, h, batch * eyes);
        k_unsqueeze_gray<<<grid, 128, 0, st>>>(left_bgr, right_bgr, pitch, stride, src_w, left_gray, right_gray,
                                               gpitch, gstride, eyes);
    } else {
        dim3 grid(((src_w + 15) / 16 + 127) / 128, h, batch * eyes);
        k_bgr2gray16<<<grid, 128, 0, st>>>(left_bgr, right_bgr, pitch, stride, src_w, left_gray, right_gray,
                                           gpitch, gstride, eyes);
    }
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}
