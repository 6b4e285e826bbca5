// Guided-filter depth upscale (colour guide = the 4K frame).
// Replaces upscale.py:47-59, where the reference merely lets ffmpeg `scale` the depth PNGs; the
// guided filter its readme promises (readme.md:97,119) does not exist upstream, so the normative
// definition is oracle/guided.py (He/Sun/Tang colour guided filter on a bilinearly upsampled depth).
//
// Two kernels per frame, both tiled 32x32 with an r-pixel halo and separable DIRECT box sums in
// shared memory (no running sums: fp32 prefix/sliding sums drift far beyond the 0.5-LSB16 budget):
//   k_guided_coeff : moments of (I, p) -> 3x3 solve -> per-pixel (a0, a1, a2, b)        [float4 plane]
//   k_guided_apply : box mean of (a, b) -> q = a.I + b -> uint16
// Moments are taken about a per-tile centre (box sums are shift-covariant) so the covariance
// subtraction does not cancel.
#include "v3d_internal.h"

namespace {

constexpr int GT = 32;      // tile edge
constexpr int GRMAX = 8;    // largest supported radius

__device__ __forceinline__ int reflect_idx(int i, int n)
{
    // fedcba|abcdef|fedcba  (cv2.BORDER_REFLECT / numpy 'symmetric')
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - i - 1;
    return min(max(i, 0), n - 1);
}

// Bilinear sample of depth/65535 at guide pixel (X, Y): half-pixel centres, clamped taps.  The source
// coordinate is formed in integers ((2X+1)*w - gw over 2*gw) so that only the final weights are rounded.
__device__ __forceinline__ float sample_depth(const uint16_t* __restrict__ d, int w, int h, int gw, int gh, int X, int Y)
{
    const int nx = (2 * X + 1) * w - gw, ny = (2 * Y + 1) * h - gh;
    const int dx2 = 2 * gw, dy2 = 2 * gh;
    int x0 = nx >= 0 ? nx / dx2 : -((-nx + dx2 - 1) / dx2);
    int y0 = ny >= 0 ? ny / dy2 : -((-ny + dy2 - 1) / dy2);
    const float fx = __fdiv_rn((float)(nx - x0 * dx2), (float)dx2);
    const float fy = __fdiv_rn((float)(ny - y0 * dy2), (float)dy2);
    const int x1 = min(max(x0 + 1, 0), w - 1), y1 = min(max(y0 + 1, 0), h - 1);
    x0 = min(max(x0, 0), w - 1); y0 = min(max(y0, 0), h - 1);
    const float s = 1.0f / 65535.0f;
    const float p00 = __ldg(d + (size_t)y0 * w + x0) * s, p01 = __ldg(d + (size_t)y0 * w + x1) * s;
    const float p10 = __ldg(d + (size_t)y1 * w + x0) * s, p11 = __ldg(d + (size_t)y1 * w + x1) * s;
    const float top = p00 * (1.0f - fx) + p01 * fx;
    const float bot = p10 * (1.0f - fx) + p11 * fx;
    return top * (1.0f - fy) + bot * fy;
}

__device__ __forceinline__ float3 load_guide(const uint8_t* __restrict__ g, int gw, int X, int Y)
{
    const uint8_t* p = g + ((size_t)Y * gw + X) * 3;
    const float s = 1.0f / 255.0f;
    return make_float3(__ldg(p) * s, __ldg(p + 1) * s, __ldg(p + 2) * s);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_guided_coeff(const uint16_t* __restrict__ depth, int w, int h, const uint8_t* __restrict__ guide, int gw, int gh,
               int r, float eps, float4* __restrict__ ab)
{
    extern __shared__ float4 gsm[];
    const int RW = GT + 2 * r, RH = GT + 2 * r, BP = RW + 1;          // base pitch (float4), odd -> conflict free
    float4* base = gsm;                                                 // [RH][BP]
    float* hs = reinterpret_cast<float*>(gsm + (size_t)RH * BP);       // [13][RH][33]
    const int HP = GT + 1;
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * GT, Y0 = blockIdx.y * GT, b = blockIdx.z;
    depth += (size_t)b * w * h;
    guide += (size_t)b * gw * gh * 3;
    ab += (size_t)b * gw * gh;

    // per-tile centre
    const int Xc = min(X0 + GT / 2, gw - 1), Yc = min(Y0 + GT / 2, gh - 1);
    const float3 cI = load_guide(guide, gw, Xc, Yc);
    const float cp = sample_depth(depth, w, h, gw, gh, Xc, Yc);

    for (int i = tid; i < RW * RH; i += blockDim.x) {
        const int j = i / RW, t = i - j * RW;
        const int X = reflect_idx(X0 - r + t, gw), Y = reflect_idx(Y0 - r + j, gh);
        const float3 I = load_guide(guide, gw, X, Y);
        const float p = sample_depth(depth, w, h, gw, gh, X, Y);
        base[j * BP + t] = make_float4(I.x - cI.x, I.y - cI.y, I.z - cI.z, p - cp);
    }
    __syncthreads();

    // horizontal sums: item = (row j, group of 4 output columns)
    for (int it = tid; it < RH * (GT / 4); it += blockDim.x) {
        const int j = it % RH, g = it / RH;
        float acc[4][13];
#pragma unroll
        for (int o = 0; o < 4; o++)
#pragma unroll
            for (int m = 0; m < 13; m++) acc[o][m] = 0.0f;
        const float4* row = base + j * BP + g * 4;
        for (int t = 0; t < 4 + 2 * r; t++) {
            const float4 v = row[t];
            const float m[13] = { v.x, v.y, v.z, v.w, v.x * v.w, v.y * v.w, v.z * v.w,
                                  v.x * v.x, v.x * v.y, v.x * v.z, v.y * v.y, v.y * v.z, v.z * v.z };
#pragma unroll
            for (int o = 0; o < 4; o++)
                if (t >= o && t <= o + 2 * r) {
#pragma unroll
                    for (int q = 0; q < 13; q++) acc[o][q] += m[q];
                }
        }
#pragma unroll
        for (int m = 0; m < 13; m++)
#pragma unroll
            for (int o = 0; o < 4; o++) hs[((size_t)m * RH + j) * HP + g * 4 + o] = acc[o][m];
    }
    __syncthreads();

    // vertical sums + solve: thread = (column, group of 4 output rows)
    {
        const int col = tid & 31, rg = tid >> 5;
        float acc[4][13];
#pragma unroll
        for (int o = 0; o < 4; o++)
#pragma unroll
            for (int m = 0; m < 13; m++) acc[o][m] = 0.0f;
        for (int t = 0; t < 4 + 2 * r; t++) {
            float m[13];
#pragma unroll
            for (int q = 0; q < 13; q++) m[q] = hs[((size_t)q * RH + rg * 4 + t) * HP + col];
#pragma unroll
            for (int o = 0; o < 4; o++)
                if (t >= o && t <= o + 2 * r) {
#pragma unroll
                    for (int q = 0; q < 13; q++) acc[o][q] += m[q];
                }
        }
        const float inv_n = 1.0f / (float)((2 * r + 1) * (2 * r + 1));
        const int X = X0 + col;
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int Y = Y0 + rg * 4 + o;
            if (X >= gw || Y >= gh) continue;
            float m[13];
#pragma unroll
            for (int q = 0; q < 13; q++) m[q] = acc[o][q] * inv_n;
            const float mI0 = m[0], mI1 = m[1], mI2 = m[2], mp = m[3];
            const float c0 = m[4] - mI0 * mp, c1 = m[5] - mI1 * mp, c2 = m[6] - mI2 * mp;
            const float s00 = m[7] - mI0 * mI0 + eps, s01 = m[8] - mI0 * mI1, s02 = m[9] - mI0 * mI2;
            const float s11 = m[10] - mI1 * mI1 + eps, s12 = m[11] - mI1 * mI2, s22 = m[12] - mI2 * mI2 + eps;
            const float k00 = s11 * s22 - s12 * s12, k01 = s02 * s12 - s01 * s22, k02 = s01 * s12 - s02 * s11;
            const float k11 = s00 * s22 - s02 * s02, k12 = s01 * s02 - s00 * s12, k22 = s00 * s11 - s01 * s01;
            const float det = s00 * k00 + s01 * k01 + s02 * k02;
            const float idet = __fdiv_rn(1.0f, det);
            const float a0 = (k00 * c0 + k01 * c1 + k02 * c2) * idet;
            const float a1 = (k01 * c0 + k11 * c1 + k12 * c2) * idet;
            const float a2 = (k02 * c0 + k12 * c1 + k22 * c2) * idet;
            // b referred to a guide centred at 0.5:  q = a.(I - 0.5) + b
            const float bb = (mp + cp) - a0 * (mI0 + cI.x - 0.5f) - a1 * (mI1 + cI.y - 0.5f) - a2 * (mI2 + cI.z - 0.5f);
            ab[(size_t)Y * gw + X] = make_float4(a0, a1, a2, bb);
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_guided_apply(const float4* __restrict__ ab, const uint8_t* __restrict__ guide, int gw, int gh, int r,
               uint16_t* __restrict__ out, float* __restrict__ qout)
{
    extern __shared__ float4 gsm[];
    const int RW = GT + 2 * r, RH = GT + 2 * r, BP = RW + 1, HP = GT + 1;
    float4* base = gsm;                               // [RH][BP]
    float4* hs = gsm + (size_t)RH * BP;               // [RH][HP]
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * GT, Y0 = blockIdx.y * GT, b = blockIdx.z;
    ab += (size_t)b * gw * gh;
    guide += (size_t)b * gw * gh * 3;
    out += (size_t)b * gw * gh;
    if (qout) qout += (size_t)b * gw * gh;

    for (int i = tid; i < RW * RH; i += blockDim.x) {
        const int j = i / RW, t = i - j * RW;
        const int X = reflect_idx(X0 - r + t, gw), Y = reflect_idx(Y0 - r + j, gh);
        base[j * BP + t] = __ldg(ab + (size_t)Y * gw + X);
    }
    __syncthreads();
    for (int it = tid; it < RH * (GT / 4); it += blockDim.x) {
        const int j = it % RH, g = it / RH;
        float4 acc[4];
#pragma unroll
        for (int o = 0; o < 4; o++) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* row = base + j * BP + g * 4;
        for (int t = 0; t < 4 + 2 * r; t++) {
            const float4 v = row[t];
#pragma unroll
            for (int o = 0; o < 4; o++)
                if (t >= o && t <= o + 2 * r) { acc[o].x += v.x; acc[o].y += v.y; acc[o].z += v.z; acc[o].w += v.w; }
        }
#pragma unroll
        for (int o = 0; o < 4; o++) hs[j * HP + g * 4 + o] = acc[o];
    }
    __syncthreads();
    {
        const int col = tid & 31, rg = tid >> 5;
        float4 acc[4];
#pragma unroll
        for (int o = 0; o < 4; o++) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < 4 + 2 * r; t++) {
            const float4 v = hs[(rg * 4 + t) * HP + col];
#pragma unroll
            for (int o = 0; o < 4; o++)
                if (t >= o && t <= o + 2 * r) { acc[o].x += v.x; acc[o].y += v.y; acc[o].z += v.z; acc[o].w += v.w; }
        }
        const float inv_n = 1.0f / (float)((2 * r + 1) * (2 * r + 1));
        const int X = X0 + col;
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int Y = Y0 + rg * 4 + o;
            if (X >= gw || Y >= gh) continue;
            const float3 I = load_guide(guide, gw, X, Y);
            const float q = (acc[o].x * (I.x - 0.5f) + acc[o].y * (I.y - 0.5f) + acc[o].z * (I.z - 0.5f) + acc[o].w) * inv_n;
            const float qc = fminf(fmaxf(q, 0.0f), 1.0f);
            out[(size_t)Y * gw + X] = (uint16_t)floorf(qc * 65535.0f + 0.5f);
            if (qout) qout[(size_t)Y * gw + X] = q;
        }
    }
}

}  // namespace

int v3d_launch_guided(v3d_ctx* ctx, const uint16_t* depth, int w, int h, const uint8_t* guide, int gw, int gh,
                      int batch, int r, float eps, uint16_t* out, float* q, cudaStream_t st)
{
    if (r < 1 || r > GRMAX) return v3d_fail(V3D_EINVAL, "guided radius %d unsupported (1..%d)", r, GRMAX);
    if (gw < 2 * r + 1 || gh < 2 * r + 1) return v3d_fail(V3D_EINVAL, "guide smaller than the filter window");
    const size_t need = (size_t)batch * gw * gh * sizeof(float4);
    if (ctx->ab_bytes < need) {
        if (ctx->ab) { V3D_CUDA(cudaStreamSynchronize(st)); V3D_CUDA(cudaFree(ctx->ab)); ctx->bytes -= ctx->ab_bytes; ctx->ab = nullptr; ctx->ab_bytes = 0; }
        V3D_CUDA(cudaMalloc(&ctx->ab, need));
        ctx->ab_bytes = need; ctx->bytes += need;
    }
    V3dScope scope(ctx, ST_GUIDED, st);
    const int RW = GT + 2 * r;
    const size_t sm_coeff = (size_t)RW * (RW + 1) * sizeof(float4) + (size_t)13 * RW * (GT + 1) * sizeof(float);
    const size_t sm_apply = (size_t)RW * (RW + 1) * sizeof(float4) + (size_t)RW * (GT + 1) * sizeof(float4);
    if (!ctx->guided_attr_set) {
        const int RWm = GT + 2 * GRMAX;
        V3D_CUDA(cudaFuncSetAttribute(k_guided_coeff, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)RWm * (RWm + 1) * 16 + (size_t)13 * RWm * (GT + 1) * 4)));
        V3D_CUDA(cudaFuncSetAttribute(k_guided_apply, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)RWm * (RWm + 1) * 16 + (size_t)RWm * (GT + 1) * 16)));
        ctx->guided_attr_set = 1;
    }
    dim3 grid((gw + GT - 1) / GT, (gh + GT - 1) / GT, batch);
    k_guided_coeff<<<grid, 256, sm_coeff, st>>>(depth, w, h, guide, gw, gh, r, eps, ctx->ab);
    k_guided_apply<<<grid, 256, sm_apply, st>>>(ctx->ab, guide, gw, gh, r, out, q);
    V3D_LAUNCHED(ctx, 2);
    return V3D_OK;
}
