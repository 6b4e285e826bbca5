// 16-bit gray PNG payloads on the GPU.  Replaces the encoding half of cv2.imwrite
// (depth.py:406): the IDAT chunk's content is produced here as a zlib stream of STORED deflate blocks
// -- pure byte placement plus the Adler-32 checksum -- so the host is left with the 8 + 25 + 12 fixed
// bytes, one CRC-32 over the payload and the file write.
//   raw stream   : per row one filter byte (0 = None) and W big-endian samples,  n = H * (1 + 2W) bytes
//   zlib stream  : 78 01 | blocks of <= 65535 raw bytes, each behind {BFINAL, LEN, ~LEN} | Adler-32 (big endian)
//   Adler-32     : a = 1 + sum d_j,  b = n + sum (n - j) d_j   (mod 65521), accumulated in 64 bits
#include "v3d_internal.h"

namespace {

constexpr unsigned BLK = 65535u;        // raw bytes per stored block (the format's maximum)

__device__ __forceinline__ size_t zpos(unsigned j) { return 2 + 5 * (size_t)(j / BLK + 1) + j; }

__global__ void __launch_bounds__(256)
k_png16_pack(const uint16_t* __restrict__ img, int W, int H, uint8_t* __restrict__ payload, size_t pstride,
             unsigned long long* __restrict__ sums)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    const unsigned rowb = 1u + 2u * (unsigned)W, n = rowb * (unsigned)H;
    uint8_t* out = payload + (size_t)b * pstride;
    unsigned long long s1 = 0, s2 = 0;
    if (x < W) {
        const unsigned v = img[((size_t)b * H + y) * W + x];
        const unsigned j = (unsigned)y * rowb + 1u + 2u * (unsigned)x;
        const unsigned hi = v >> 8, lo = v & 0xffu;
        if (x == 0) out[zpos(j - 1)] = 0;                  // filter type None
        out[zpos(j)] = (uint8_t)hi;
        out[zpos(j + 1)] = (uint8_t)lo;
        s1 = hi + lo;
        s2 = (unsigned long long)(n - j) * hi + (unsigned long long)(n - j - 1) * lo;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(V3D_FULL_MASK, s1, o);
        s2 += __shfl_xor_sync(V3D_FULL_MASK, s2, o);
    }
    __shared__ unsigned long long w1[8], w2[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { w1[wid] = s1; w2[wid] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { s1 += w1[i]; s2 += w2[i]; }
        atomicAdd(&sums[2 * b], s1);
        atomicAdd(&sums[2 * b + 1], s2);
    }
}

// zlib header, the stored-block headers and the checksum; one CTA per frame
__global__ void __launch_bounds__(256)
k_png16_frame(int W, int H, uint8_t* __restrict__ payload, size_t pstride, const unsigned long long* __restrict__ sums)
{
    const int b = blockIdx.x;
    const unsigned n = (1u + 2u * (unsigned)W) * (unsigned)H;
    const unsigned nblk = (n + BLK - 1) / BLK;
    uint8_t* out = payload + (size_t)b * pstride;
    for (unsigned k = threadIdx.x; k < nblk; k += blockDim.x) {
        const unsigned len = min(BLK, n - k * BLK);
        uint8_t* hp = out + 2 + (size_t)k * (BLK + 5);
        hp[0] = k + 1 == nblk ? 1 : 0;                     // BFINAL, BTYPE = 00 (stored)
        hp[1] = (uint8_t)(len & 0xff); hp[2] = (uint8_t)(len >> 8);
        hp[3] = (uint8_t)(~len & 0xff); hp[4] = (uint8_t)((~len >> 8) & 0xff);
    }
    if (threadIdx.x == 0) {
        out[0] = 0x78; out[1] = 0x01;                      // deflate, 32K window, no preset dictionary, fastest
        const unsigned a = (unsigned)((1ull + sums[2 * b]) % 65521ull);
        const unsigned bb = (unsigned)(((unsigned long long)n + sums[2 * b + 1]) % 65521ull);
        uint8_t* t = out + 2 + 5 * (size_t)nblk + n;
        t[0] = (uint8_t)(bb >> 8); t[1] = (uint8_t)bb; t[2] = (uint8_t)(a >> 8); t[3] = (uint8_t)a;
    }
}

}  // namespace

size_t v3d_png16_raw_bytes(int w, int h) { return (size_t)h * (1 + 2 * (size_t)w); }

int v3d_launch_png16_pack(v3d_ctx* ctx, const uint16_t* img, int w, int h, int batch, uint8_t* payload,
                          size_t payload_stride, cudaStream_t st)
{
    V3dScope scope(ctx, ST_POST, st);
    V3D_CUDA(cudaMemsetAsync(ctx->png_sums, 0, (size_t)batch * 2 * sizeof(unsigned long long), st));
    dim3 grid((w + 255) / 256, h, batch);
    k_png16_pack<<<grid, 256, 0, st>>>(img, w, h, payload, payload_stride, ctx->png_sums);
    k_png16_frame<<<batch, 256, 0, st>>>(w, h, payload, payload_stride, ctx->png_sums);
    V3D_LAUNCHED(ctx, 2);
    return V3D_OK;
}
