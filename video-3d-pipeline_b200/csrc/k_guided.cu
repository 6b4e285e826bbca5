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
constexpr int GRUN = 8;     // outputs per thread along the filtered axis (first direct, the rest slid)
constexpr int GRUNV = 4;    // run length of the vertical pass: shorter runs keep all 256 threads of a tile busy
constexpr int GRMAX = 8;    // largest supported radius

__device__ __forceinline__ int reflect_idx(int i, int n)
{
    // fedcba|abcdef|fedcba  (cv2.BORDER_REFLECT / numpy 'symmetric')
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - i - 1;
    return min(max(i, 0), n - 1);
}

// Bilinear tap of one axis at guide coordinate X (already reflected): half-pixel centres, clamped taps.
// The source coordinate is formed in integers ((2X+1)*w - gw over 2*gw) so that only the weight is rounded.
__device__ __forceinline__ void axis_tap(int X, int w, int gw, int& i0, int& i1, float& f)
{
    const int n = (2 * X + 1) * w - gw, d2 = 2 * gw;
    int q = n >= 0 ? n / d2 : -((-n + d2 - 1) / d2);
    f = __fdiv_rn((float)(n - q * d2), (float)d2);
    i1 = min(max(q + 1, 0), w - 1);
    i0 = min(max(q, 0), w - 1);
}

__device__ __forceinline__ float3 load_guide(const uint8_t* __restrict__ g, int gw, int X, int Y)
{
    const uint8_t* p = g + ((size_t)Y * gw + X) * 3;
    const float s = 1.0f / 255.0f;
    return make_float3(__ldg(p) * s, __ldg(p + 1) * s, __ldg(p + 2) * s);
}

__device__ __forceinline__ void moments13(const float4& v, float (&m)[13])
{
    m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
    m[4] = v.x * v.w; m[5] = v.y * v.w; m[6] = v.z * v.w;
    m[7] = v.x * v.x; m[8] = v.x * v.y; m[9] = v.x * v.z;
    m[10] = v.y * v.y; m[11] = v.y * v.z; m[12] = v.z * v.z;
}

struct TileTaps {           // per-tile sampling tables (one entry per region column / row)
    int gx[GT + 2 * GRMAX], gy[GT + 2 * GRMAX];              // reflected guide coordinates
    int x0[GT + 2 * GRMAX], x1[GT + 2 * GRMAX], y0[GT + 2 * GRMAX], y1[GT + 2 * GRMAX];
    float fx[GT + 2 * GRMAX], fy[GT + 2 * GRMAX];
};

// ------------------------------------------------------------------------------------------------
// RT > 0: radius known at compile time (the loops unroll); RT == 0: runtime radius.
template <int RT>
__global__ void __launch_bounds__(256)
k_guided_coeff(const uint16_t* __restrict__ depth, int w, int h, const uint8_t* __restrict__ guide, int gw, int gh,
               int r_arg, float eps, float4* __restrict__ ab)
{
    extern __shared__ float4 gsm[];
    const int r = RT > 0 ? RT : r_arg;
    const int RW = GT + 2 * r, RH = GT + 2 * r, BP = RW + 1;          // base pitch (8-byte entries), odd
    const int HP = GT + 1;
    uint2* base = reinterpret_cast<uint2*>(gsm);                       // [RH][BP] {rgb bytes, p - centre}
    float* hs = reinterpret_cast<float*>(base + (size_t)RH * BP);      // [13][RH][HP]
    __shared__ TileTaps tp;
    __shared__ float4 centre;
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * GT, Y0 = blockIdx.y * GT, b = blockIdx.z;
    depth += (size_t)b * w * h;
    guide += (size_t)b * gw * gh * 3;
    ab += (size_t)b * gw * gh;

    if (tid < RW) {
        const int X = reflect_idx(X0 - r + tid, gw);
        tp.gx[tid] = X;
        axis_tap(X, w, gw, tp.x0[tid], tp.x1[tid], tp.fx[tid]);
    } else if (tid >= 64 && tid < 64 + RH) {
        const int j = tid - 64;
        const int Y = reflect_idx(Y0 - r + j, gh);
        tp.gy[j] = Y;
        axis_tap(Y, h, gh, tp.y0[j], tp.y1[j], tp.fy[j]);
    }
    __syncthreads();
    auto sample = [&](int t, int j, uint32_t& rgb) -> float {
        const uint8_t* gp = guide + ((size_t)tp.gy[j] * gw + tp.gx[t]) * 3;
        rgb = (uint32_t)__ldg(gp) | ((uint32_t)__ldg(gp + 1) << 8) | ((uint32_t)__ldg(gp + 2) << 16);
        const float s = 1.0f / 65535.0f;
        const uint16_t* r0 = depth + (size_t)tp.y0[j] * w;
        const uint16_t* r1 = depth + (size_t)tp.y1[j] * w;
        const float fx = tp.fx[t], fy = tp.fy[j];
        const float p00 = __ldg(r0 + tp.x0[t]) * s, p01 = __ldg(r0 + tp.x1[t]) * s;
        const float p10 = __ldg(r1 + tp.x0[t]) * s, p11 = __ldg(r1 + tp.x1[t]) * s;
        const float top = p00 * (1.0f - fx) + p01 * fx;
        const float bot = p10 * (1.0f - fx) + p11 * fx;
        return top * (1.0f - fy) + bot * fy;
    };
    const float k255 = 1.0f / 255.0f;
    // per-tile centre: moments are taken about it (box sums are shift covariant)
    if (tid == 0) {
        uint32_t rgb;
        const float p = sample(min(r + GT / 2, RW - 1), min(r + GT / 2, RH - 1), rgb);
        centre = make_float4((rgb & 0xff) * k255, ((rgb >> 8) & 0xff) * k255, ((rgb >> 16) & 0xff) * k255, p);
    }
    __syncthreads();
    const float4 cc = centre;
#pragma unroll 3
    for (int i = tid; i < RW * RH; i += 256) {
        const int j = i / RW, t = i - j * RW;
        uint32_t rgb;
        const float p = sample(t, j, rgb);
        base[j * BP + t] = make_uint2(rgb, __float_as_uint(p - cc.w));
    }
    __syncthreads();
    auto tap = [&](const uint2 e) -> float4 {     // centred (I, p) of one region pixel
        return make_float4(fmaf((float)(e.x & 0xff), k255, -cc.x), fmaf((float)((e.x >> 8) & 0xff), k255, -cc.y),
                           fmaf((float)((e.x >> 16) & 0xff), k255, -cc.z), __uint_as_float(e.y));
    };

    // horizontal box sums: item = (row j, run of GRUN output columns); first output direct, rest slid
    for (int it = tid; it < RH * (GT / GRUN); it += 256) {
        const int j = it % RH, g = it / RH;
        const uint2* row = base + j * BP + g * GRUN;
        float acc[13], m[13];
#pragma unroll
        for (int q = 0; q < 13; q++) acc[q] = 0.0f;
#pragma unroll 1
        for (int t = 0; t <= 2 * r; t++) {
            moments13(tap(row[t]), m);
#pragma unroll
            for (int q = 0; q < 13; q++) acc[q] += m[q];
        }
        float* out = hs + (size_t)j * HP + g * GRUN;
#pragma unroll
        for (int q = 0; q < 13; q++) out[(size_t)q * RH * HP] = acc[q];
#pragma unroll 1
        for (int o = 1; o < GRUN; o++) {
            moments13(tap(row[o + 2 * r]), m);
#pragma unroll
            for (int q = 0; q < 13; q++) acc[q] += m[q];
            moments13(tap(row[o - 1]), m);
#pragma unroll
            for (int q = 0; q < 13; q++) acc[q] -= m[q];
#pragma unroll
            for (int q = 0; q < 13; q++) out[(size_t)q * RH * HP + o] = acc[q];
        }
    }
    __syncthreads();

    // vertical box sums + 3x3 solve: item = (column, run of GRUN output rows)
    if (tid < GT * (GT / GRUNV)) {
        const int col = tid & 31, rg = tid >> 5;
        const float* hc = hs + (size_t)(rg * GRUNV) * HP + col;
        float acc[13];
#pragma unroll
        for (int q = 0; q < 13; q++) acc[q] = 0.0f;
#pragma unroll 1
        for (int t = 0; t <= 2 * r; t++) {
#pragma unroll
            for (int q = 0; q < 13; q++) acc[q] += hc[((size_t)q * RH + t) * HP];
        }
        const float inv_n = 1.0f / (float)((2 * r + 1) * (2 * r + 1));
        const int X = X0 + col;
#pragma unroll 1
        for (int o = 0; o < GRUNV; o++) {
            if (o > 0) {
#pragma unroll
                for (int q = 0; q < 13; q++)
                    acc[q] += hc[((size_t)q * RH + o + 2 * r) * HP] - hc[((size_t)q * RH + o - 1) * HP];
            }
            const int Y = Y0 + rg * GRUNV + o;
            if (X >= gw || Y >= gh) continue;
            float m[13];
#pragma unroll
            for (int q = 0; q < 13; q++) m[q] = acc[q] * inv_n;
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
            const float bb = (mp + cc.w) - a0 * (mI0 + cc.x - 0.5f) - a1 * (mI1 + cc.y - 0.5f) - a2 * (mI2 + cc.z - 0.5f);
            ab[(size_t)Y * gw + X] = make_float4(a0, a1, a2, bb);
        }
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void sub4(float4& a, const float4& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }

template <int RT>
__global__ void __launch_bounds__(256)
k_guided_apply(const float4* __restrict__ ab, const uint8_t* __restrict__ guide, int gw, int gh, int r_arg,
               uint16_t* __restrict__ out, float* __restrict__ qout)
{
    extern __shared__ float4 gsm[];
    const int r = RT > 0 ? RT : r_arg;
    const int RW = GT + 2 * r, RH = GT + 2 * r, BP = RW + 1, HP = GT + 1;
    float4* base = gsm;                               // [RH][BP]
    float4* hs = gsm + (size_t)RH * BP;               // [RH][HP]
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * GT, Y0 = blockIdx.y * GT, b = blockIdx.z;
    ab += (size_t)b * gw * gh;
    guide += (size_t)b * gw * gh * 3;
    out += (size_t)b * gw * gh;
    if (qout) qout += (size_t)b * gw * gh;

    // region load: 16-byte cp.async (LDGSTS) straight into shared memory, all of a thread's copies in flight
    for (int i = tid; i < RW * RH; i += 256) {
        const int j = i / RW, t = i - j * RW;
        const float4* src = ab + (size_t)reflect_idx(Y0 - r + j, gh) * gw + reflect_idx(X0 - r + t, gw);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(base + j * BP + t);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int it = tid; it < RH * (GT / GRUN); it += 256) {
        const int j = it % RH, g = it / RH;
        const float4* row = base + j * BP + g * GRUN;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int t = 0; t <= 2 * r; t++) add4(acc, row[t]);
        float4* o4 = hs + j * HP + g * GRUN;
        o4[0] = acc;
#pragma unroll
        for (int o = 1; o < GRUN; o++) { add4(acc, row[o + 2 * r]); sub4(acc, row[o - 1]); o4[o] = acc; }
    }
    __syncthreads();
    if (tid < GT * (GT / GRUNV)) {
        const int col = tid & 31, rg = tid >> 5;
        const float4* hc = hs + (rg * GRUNV) * HP + col;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int t = 0; t <= 2 * r; t++) add4(acc, hc[t * HP]);
        const float inv_n = 1.0f / (float)((2 * r + 1) * (2 * r + 1));
        const int X = X0 + col;
#pragma unroll
        for (int o = 0; o < GRUNV; o++) {
            if (o > 0) { add4(acc, hc[(o + 2 * r) * HP]); sub4(acc, hc[(o - 1) * HP]); }
            const int Y = Y0 + rg * GRUNV + o;
            if (X >= gw || Y >= gh) continue;
            const float3 I = load_guide(guide, gw, X, Y);
            const float q = (acc.x * (I.x - 0.5f) + acc.y * (I.y - 0.5f) + acc.z * (I.z - 0.5f) + acc.w) * inv_n;
            const float qc = fminf(fmaxf(q, 0.0f), 1.0f);
            out[(size_t)Y * gw + X] = (uint16_t)floorf(qc * 65535.0f + 0.5f);
            if (qout) qout[(size_t)Y * gw + X] = q;
        }
    }
}

constexpr size_t sm_coeff_max()
{
    return (size_t)(GT + 2 * GRMAX) * (GT + 2 * GRMAX + 1) * 8 + (size_t)13 * (GT + 2 * GRMAX) * (GT + 1) * 4;
}
constexpr size_t sm_apply_max()
{
    return (size_t)(GT + 2 * GRMAX) * (GT + 2 * GRMAX + 1) * 16 + (size_t)(GT + 2 * GRMAX) * (GT + 1) * 16;
}

template <int RT>
int launch_guided_rt(v3d_ctx* ctx, const uint16_t* depth, int w, int h, const uint8_t* guide, int gw, int gh,
                     int batch, int r, float eps, uint16_t* out, float* q, cudaStream_t st)
{
    const int RW = GT + 2 * r;
    const size_t sm_coeff = (size_t)RW * (RW + 1) * sizeof(uint2) + (size_t)13 * RW * (GT + 1) * sizeof(float);
    const size_t sm_apply = (size_t)RW * (RW + 1) * sizeof(float4) + (size_t)RW * (GT + 1) * sizeof(float4);
    if (!(ctx->guided_attr_set & (1 << RT))) {
        V3D_CUDA(cudaFuncSetAttribute(k_guided_coeff<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_coeff_max()));
        V3D_CUDA(cudaFuncSetAttribute(k_guided_apply<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_apply_max()));
        ctx->guided_attr_set |= (1 << RT);
    }
    dim3 grid((gw + GT - 1) / GT, (gh + GT - 1) / GT, batch);
    k_guided_coeff<RT><<<grid, 256, sm_coeff, st>>>(depth, w, h, guide, gw, gh, r, eps, ctx->ab);
    k_guided_apply<RT><<<grid, 256, sm_apply, st>>>(ctx->ab, guide, gw, gh, r, out, q);
    V3D_LAUNCHED(ctx, 2);
    return V3D_OK;
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
    switch (r) {
        case 8: return launch_guided_rt<8>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
        case 4: return launch_guided_rt<4>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
        default: return launch_guided_rt<0>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
    }
}
