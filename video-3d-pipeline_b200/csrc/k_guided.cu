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
#include <cstdlib>

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

// ================================================================================================
// Streaming variant: a CTA owns a strip of TW output columns (NT = TW + 2r region columns, one thread
// each) and walks down a segment of rows.  Vertical box sums live in registers as running sums
// (enter the new row, leave the row 2r+1 above, which a (2r+1)-row ring in shared memory remembers);
// every R rows the column sums go through shared memory once for the horizontal pass.
//   * guide moments are kept in BYTE units about an integer centre: I, I.I sums are integers below
//     2^24, so fp32 adds them exactly -- no drift however long the strip;
//   * the four planes that involve the depth (p, I.p) are compensated (Kahan) running sums, so their
//     error stays that of one direct 17-term sum; the horizontal pass slides over runs of GR only.
struct RowTap { int gy, y0, y1; float fy; };

__device__ __forceinline__ float u16f(uint32_t v) { return __uint_as_float(0x4b000000u | v) - 8388608.0f; }
// byte k of `w` as 8388608 + byte (one PRMT); subtract (8388608 + centre) to get the centred value exactly
template <int K> __device__ __forceinline__ float byte_magic(uint32_t w)
{
    return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7540 | K));
}
__device__ __forceinline__ void kahan(float& v, float& c, float d)
{
    const float y = __fsub_rn(d, c), t = __fadd_rn(v, y);
    c = __fsub_rn(__fsub_rn(t, v), y);
    v = t;
}

template <int RT, int NT, int R, int GR>
__global__ void __launch_bounds__(NT, 512 / NT)
k_guided_coeff_s(const uint16_t* __restrict__ depth, int w, int h, const uint8_t* __restrict__ guide, int gw, int gh,
                 int r_arg, float eps, int seg, float4* __restrict__ ab)
{
    extern __shared__ float4 gsm[];
    const int r = RT > 0 ? RT : r_arg, win = 2 * r + 1;
    const int TW = (NT - 2 * r) & ~7;
    constexpr int VP = NT + NT / GR + 1;                            // float4 slots per (plane group, row)
    float4* vbuf = gsm;                                             // [4 plane groups][R][VP], column x at x + x/GR
    uint2* ring = reinterpret_cast<uint2*>(gsm + 4 * R * VP);       // [win][NT] {rgb bytes, p - centre}
    __shared__ RowTap taps[2][R];
    __shared__ float centre_p;
    __shared__ uint32_t centre_rgb;
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * seg, b = blockIdx.z;
    const int seg_h = min(seg, gh - Y0), nrows = seg_h + 2 * r;
    depth += (size_t)b * w * h;
    guide += (size_t)b * gw * gh * 3;
    ab += (size_t)b * gw * gh;
    const float s16 = 1.0f / 65535.0f, k255 = 1.0f / 255.0f;

    const int gx = reflect_idx(X0 - r + tid, gw);
    int x0, x1;
    float fx;
    axis_tap(gx, w, gw, x0, x1, fx);
    auto rowlerp = [&](int y) -> float {
        const uint16_t* rp = depth + (size_t)y * w;
        return u16f(__ldg(rp + x0)) * s16 * (1.0f - fx) + u16f(__ldg(rp + x1)) * s16 * fx;
    };
    auto make_tap = [&](int j) -> RowTap {
        RowTap t;
        t.gy = reflect_idx(Y0 - r + j, gh);
        axis_tap(t.gy, h, gh, t.y0, t.y1, t.fy);
        return t;
    };
    if (tid < R) taps[0][tid] = make_tap(tid);
    if (tid == NT / 2) {                                            // strip centre: column X0 + TW/2, middle row
        const int Yc = min(Y0 + seg_h / 2, gh - 1);
        int y0, y1;
        float fy;
        axis_tap(Yc, h, gh, y0, y1, fy);
        const uint8_t* gp = guide + ((size_t)Yc * gw + gx) * 3;
        centre_rgb = (uint32_t)__ldg(gp) | ((uint32_t)__ldg(gp + 1) << 8) | ((uint32_t)__ldg(gp + 2) << 16);
        centre_p = rowlerp(y0) * (1.0f - fy) + rowlerp(y1) * fy;
    }
    __syncthreads();
    const uint32_t crgb = centre_rgb;
    const float cp = centre_p;
    const float cb0 = 8388608.0f + (float)(crgb & 0xff), cb1 = 8388608.0f + (float)((crgb >> 8) & 0xff),
                cb2 = 8388608.0f + (float)((crgb >> 16) & 0xff);
    for (int k = 0; k < win; k++) ring[k * NT + tid] = make_uint2(crgb, 0u);   // centre pixel: all moments 0

    float V[13], C[4];
#pragma unroll
    for (int q = 0; q < 13; q++) V[q] = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) C[q] = 0.0f;
    int cy0 = -1, cy1 = -1, rslot = 0;
    float ctop = 0.0f, cbot = 0.0f;
    const int wslot = tid + tid / GR;
    const int runs = TW / GR;
    const float inv_n = 1.0f / (float)(win * win);
    const uint8_t* gcol = guide + (size_t)gx * 3;

    for (int g = 0; g * R < nrows; g++) {
        const RowTap* tp = taps[g & 1];
        uint32_t rgbv[R];
#pragma unroll
        for (int jj = 0; jj < R; jj++) {                            // all guide loads of the group in flight first
            if (g * R + jj < nrows) {
                const uint8_t* gp = gcol + (size_t)tp[jj].gy * gw * 3;
                rgbv[jj] = (uint32_t)__ldg(gp) | ((uint32_t)__ldg(gp + 1) << 8) | ((uint32_t)__ldg(gp + 2) << 16);
            }
        }
#pragma unroll
        for (int jj = 0; jj < R; jj++) {
            const int j = g * R + jj;
            if (j >= nrows) break;
            const RowTap rt = tp[jj];
            const uint32_t rgb = rgbv[jj];
            float nt, nb;
            if (rt.y0 == cy0) nt = ctop; else if (rt.y0 == cy1) nt = cbot; else nt = rowlerp(rt.y0);
            if (rt.y1 == cy1) nb = cbot; else if (rt.y1 == cy0) nb = ctop; else nb = rowlerp(rt.y1);
            cy0 = rt.y0; cy1 = rt.y1; ctop = nt; cbot = nb;
            const float ap = nt * (1.0f - rt.fy) + nb * rt.fy - cp;
            uint2* slot = ring + rslot * NT + tid;
            rslot = (rslot + 1 == win) ? 0 : rslot + 1;
            const uint2 old = *slot;
            *slot = make_uint2(rgb, __float_as_uint(ap));
            const float a0 = byte_magic<0>(rgb) - cb0, a1 = byte_magic<1>(rgb) - cb1, a2 = byte_magic<2>(rgb) - cb2;
            const float b0 = byte_magic<0>(old.x) - cb0, b1 = byte_magic<1>(old.x) - cb1, b2 = byte_magic<2>(old.x) - cb2;
            const float bp = __uint_as_float(old.y);
            // exact planes (integers): I, I.I
            V[0] += a0 - b0; V[1] += a1 - b1; V[2] += a2 - b2;
            V[7] += fmaf(a0, a0, -(b0 * b0));  V[8] += fmaf(a0, a1, -(b0 * b1));  V[9] += fmaf(a0, a2, -(b0 * b2));
            V[10] += fmaf(a1, a1, -(b1 * b1)); V[11] += fmaf(a1, a2, -(b1 * b2)); V[12] += fmaf(a2, a2, -(b2 * b2));
            // depth planes: compensated
            kahan(V[3], C[0], ap - bp);
            kahan(V[4], C[1], fmaf(a0, ap, -(b0 * bp)));
            kahan(V[5], C[2], fmaf(a1, ap, -(b1 * bp)));
            kahan(V[6], C[3], fmaf(a2, ap, -(b2 * bp)));
            if (j >= 2 * r) {
                float4* vr = vbuf + jj * VP + wslot;
                vr[0] = make_float4(V[0], V[1], V[2], V[3]);
                vr[R * VP] = make_float4(V[4], V[5], V[6], V[7]);
                vr[2 * R * VP] = make_float4(V[8], V[9], V[10], V[11]);
                reinterpret_cast<float*>(vr + 3 * R * VP)[0] = V[12];
            }
        }
        if (tid < R) taps[(g + 1) & 1][tid] = make_tap((g + 1) * R + tid);
        __syncthreads();

        // horizontal pass + 3x3 solve: item = (row of the group, run of GR output columns)
        for (int it = tid; it < R * runs; it += NT) {
            const int jj = it / runs, run = it - jj * runs;
            const int o = g * R + jj - 2 * r;
            const int xb = run * GR;
            if (o < 0 || o >= seg_h || X0 + xb >= gw) continue;
            const float4* vr = vbuf + jj * VP + xb + run;          // window start; xb is a multiple of GR
            float acc[13], m[13];
            auto ld13 = [&](int dx) {                               // dx = offset from the window start
                const float4* e = vr + (RT > 0 ? dx + dx / GR : (xb + dx) + (xb + dx) / GR - xb - run);
                const float4 g0 = e[0], g1 = e[R * VP], g2 = e[2 * R * VP];
                m[0] = g0.x; m[1] = g0.y; m[2] = g0.z; m[3] = g0.w;
                m[4] = g1.x; m[5] = g1.y; m[6] = g1.z; m[7] = g1.w;
                m[8] = g2.x; m[9] = g2.y; m[10] = g2.z; m[11] = g2.w;
                m[12] = reinterpret_cast<const float*>(e + 3 * R * VP)[0];
            };
#pragma unroll
            for (int q = 0; q < 13; q++) acc[q] = 0.0f;
            if (RT > 0) {
                const float4* v0 = vr;
#pragma unroll 1
                for (int t0 = 0; t0 + GR <= 2 * RT + 1; t0 += GR, vr += GR + 1) {   // whole runs: pad advances with them
#pragma unroll
                    for (int t = 0; t < GR; t++) {
                        ld13(t);
#pragma unroll
                        for (int q = 0; q < 13; q++) acc[q] += m[q];
                    }
                }
                vr = v0;
#pragma unroll
                for (int t = (2 * RT + 1) / GR * GR; t <= 2 * RT; t++) {
                    ld13(t);
#pragma unroll
                    for (int q = 0; q < 13; q++) acc[q] += m[q];
                }
            } else {
#pragma unroll 1
                for (int t = 0; t <= 2 * r; t++) {
                    ld13(t);
#pragma unroll
                    for (int q = 0; q < 13; q++) acc[q] += m[q];
                }
            }
            const size_t orow = (size_t)(Y0 + o) * gw;
            const float kn = k255 * inv_n, kkn = k255 * k255 * inv_n;
            const float cI0 = (float)(crgb & 0xff) * k255 - 0.5f, cI1 = (float)((crgb >> 8) & 0xff) * k255 - 0.5f,
                        cI2 = (float)((crgb >> 16) & 0xff) * k255 - 0.5f;
#pragma unroll
            for (int o2 = 0; o2 < GR; o2++) {
                if (o2 > 0) {
                    ld13(o2 + 2 * r);
#pragma unroll
                    for (int q = 0; q < 13; q++) acc[q] += m[q];
                    ld13(o2 - 1);
#pragma unroll
                    for (int q = 0; q < 13; q++) acc[q] -= m[q];
                }
                const int X = X0 + xb + o2;
                if (X >= gw) break;
                const float mI0 = acc[0] * kn, mI1 = acc[1] * kn, mI2 = acc[2] * kn, mp = acc[3] * inv_n;
                const float c0 = acc[4] * kn - mI0 * mp, c1 = acc[5] * kn - mI1 * mp, c2 = acc[6] * kn - mI2 * mp;
                const float s00 = acc[7] * kkn - mI0 * mI0 + eps, s01 = acc[8] * kkn - mI0 * mI1, s02 = acc[9] * kkn - mI0 * mI2;
                const float s11 = acc[10] * kkn - mI1 * mI1 + eps, s12 = acc[11] * kkn - mI1 * mI2, s22 = acc[12] * kkn - mI2 * mI2 + eps;
                const float k00 = s11 * s22 - s12 * s12, k01 = s02 * s12 - s01 * s22, k02 = s01 * s12 - s02 * s11;
                const float k11 = s00 * s22 - s02 * s02, k12 = s01 * s02 - s00 * s12, k22 = s00 * s11 - s01 * s01;
                const float det = s00 * k00 + s01 * k01 + s02 * k02;
                const float idet = __fdiv_rn(1.0f, det);
                const float a0 = (k00 * c0 + k01 * c1 + k02 * c2) * idet;
                const float a1 = (k01 * c0 + k11 * c1 + k12 * c2) * idet;
                const float a2 = (k02 * c0 + k12 * c1 + k22 * c2) * idet;
                // b referred to a guide centred at 0.5:  q = a.(I - 0.5) + b
                const float bb = (mp + cp) - a0 * (mI0 + cI0) - a1 * (mI1 + cI1) - a2 * (mI2 + cI2);
                ab[orow + X] = make_float4(a0, a1, a2, bb);
            }
        }
        __syncthreads();
    }
}

template <int RT, int NT, int R, int GR>
__global__ void __launch_bounds__(NT)
k_guided_apply_s(const float4* __restrict__ ab, const uint8_t* __restrict__ guide, int gw, int gh, int r_arg, int seg,
                 int vec_ok, uint16_t* __restrict__ out, float* __restrict__ qout)
{
    extern __shared__ float4 gsm[];
    const int r = RT > 0 ? RT : r_arg, win = 2 * r + 1;
    const int TW = (NT - 2 * r) & ~7;
    constexpr int VP = NT + NT / GR + 1;
    float4* vbuf = gsm;                  // [R][VP]
    float4* ring = gsm + R * VP;         // [win][NT]
    __shared__ int rows[2][R];
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * seg, b = blockIdx.z;
    const int seg_h = min(seg, gh - Y0), nrows = seg_h + 2 * r;
    ab += (size_t)b * gw * gh;
    guide += (size_t)b * gw * gh * 3;
    out += (size_t)b * gw * gh;
    if (qout) qout += (size_t)b * gw * gh;
    const int gx = reflect_idx(X0 - r + tid, gw);
    for (int k = 0; k < win; k++) ring[k * NT + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < R) rows[0][tid] = reflect_idx(Y0 - r + tid, gh);
    __syncthreads();
    float4 V = make_float4(0.f, 0.f, 0.f, 0.f), C = V;
    int rslot = 0;
    const int runs = TW / GR;
    const float inv_n = 1.0f / (float)(win * win), k255 = 1.0f / 255.0f;
    const int wslot = tid + tid / GR;

    for (int g = 0; g * R < nrows; g++) {
        const int* rw = rows[g & 1];
        float4 nv[R];
#pragma unroll
        for (int jj = 0; jj < R; jj++)
            if (g * R + jj < nrows) nv[jj] = __ldg(ab + (size_t)rw[jj] * gw + gx);
#pragma unroll
        for (int jj = 0; jj < R; jj++) {
            const int j = g * R + jj;
            if (j >= nrows) break;
            float4* slot = ring + rslot * NT + tid;
            rslot = (rslot + 1 == win) ? 0 : rslot + 1;
            const float4 old = *slot;
            *slot = nv[jj];
            kahan(V.x, C.x, nv[jj].x - old.x); kahan(V.y, C.y, nv[jj].y - old.y);
            kahan(V.z, C.z, nv[jj].z - old.z); kahan(V.w, C.w, nv[jj].w - old.w);
            if (j >= 2 * r) vbuf[jj * VP + wslot] = V;
        }
        if (tid < R) rows[(g + 1) & 1][tid] = reflect_idx(Y0 - r + (g + 1) * R + tid, gh);
        __syncthreads();

        for (int it = tid; it < R * runs; it += NT) {
            const int jj = it / runs, run = it - jj * runs;
            const int o = g * R + jj - 2 * r;
            const int xb = run * GR, X = X0 + xb;
            if (o < 0 || o >= seg_h || X >= gw) continue;
            const int Y = Y0 + o;
            const float4* vr = vbuf + jj * VP + xb + run;          // window start; xb is a multiple of GR
            auto at = [&](int dx) -> float4 { return vr[RT > 0 ? dx + dx / GR : (xb + dx) + (xb + dx) / GR - xb - run]; };
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (RT > 0) {
#pragma unroll
                for (int t = 0; t <= 2 * RT; t++) add4(acc, at(t));
            } else {
#pragma unroll 1
                for (int t = 0; t <= 2 * r; t++) add4(acc, at(t));
            }
            const size_t base = (size_t)Y * gw + X;
            const bool vec = vec_ok && GR == 8 && X + GR <= gw;
            uint32_t gbytes[6];
            if (vec) {
                const uint2* gp = reinterpret_cast<const uint2*>(guide + base * 3);
                const uint2 u0 = __ldg(gp), u1 = __ldg(gp + 1), u2 = __ldg(gp + 2);
                gbytes[0] = u0.x; gbytes[1] = u0.y; gbytes[2] = u1.x; gbytes[3] = u1.y; gbytes[4] = u2.x; gbytes[5] = u2.y;
            }
            uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int o2 = 0; o2 < GR; o2++) {
                if (o2 > 0) { add4(acc, at(o2 + 2 * r)); sub4(acc, at(o2 - 1)); }
                if (X + o2 >= gw) break;
                float I0, I1, I2;
                if (vec) {
                    auto byte_at = [&](int k) -> float {
                        const uint32_t wd = gbytes[(k >> 2) % 6];
                        return __uint_as_float(__byte_perm(wd, 0x4b000000u, 0x7540 | (k & 3))) - 8388608.0f;
                    };
                    I0 = byte_at(3 * o2); I1 = byte_at(3 * o2 + 1); I2 = byte_at(3 * o2 + 2);
                } else {
                    const uint8_t* gp = guide + (base + o2) * 3;
                    I0 = (float)__ldg(gp); I1 = (float)__ldg(gp + 1); I2 = (float)__ldg(gp + 2);
                }
                const float q = (acc.x * fmaf(I0, k255, -0.5f) + acc.y * fmaf(I1, k255, -0.5f) + acc.z * fmaf(I2, k255, -0.5f) + acc.w) * inv_n;
                const float qc = fminf(fmaxf(q, 0.0f), 1.0f);
                const uint32_t u = (uint32_t)floorf(qc * 65535.0f + 0.5f);
                if (vec) packed[(o2 >> 1) & 3] |= u << ((o2 & 1) * 16);
                else out[base + o2] = (uint16_t)u;
                if (qout) qout[base + o2] = q;
            }
            if (vec) *reinterpret_cast<uint4*>(out + base) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        __syncthreads();
    }
}

constexpr int SNT = 128, SR = 4, SGR = 8;

template <int RT>
int launch_guided_stream(v3d_ctx* ctx, const uint16_t* depth, int w, int h, const uint8_t* guide, int gw, int gh,
                         int batch, int r, float eps, uint16_t* out, float* q, cudaStream_t st)
{
    const int win = 2 * r + 1, TW = (SNT - 2 * r) & ~7;
    const int strips = (gw + TW - 1) / TW;
    // segments: enough CTAs for ~3 waves of a full machine, never shorter than 64 rows
    int segs = (3 * 148 * 4 + strips * batch - 1) / (strips * batch);
    segs = max(1, min(segs, gh / 64));
    int seg = (gh + segs - 1) / segs;
    seg = (seg + SR - 1) / SR * SR;
    segs = (gh + seg - 1) / seg;
    const size_t sm_c = (size_t)4 * SR * (SNT + SNT / SGR + 1) * 16 + (size_t)win * SNT * 8;
    const size_t sm_a = (size_t)SR * (SNT + SNT / SGR + 1) * 16 + (size_t)win * SNT * 16;
    if (!(ctx->guided_attr_set & (0x100 << RT))) {
        V3D_CUDA(cudaFuncSetAttribute(k_guided_coeff_s<RT, SNT, SR, SGR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)4 * SR * (SNT + SNT / SGR + 1) * 16 + (size_t)(2 * GRMAX + 1) * SNT * 8)));
        V3D_CUDA(cudaFuncSetAttribute(k_guided_apply_s<RT, SNT, SR, SGR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)SR * (SNT + SNT / SGR + 1) * 16 + (size_t)(2 * GRMAX + 1) * SNT * 16)));
        ctx->guided_attr_set |= (0x100 << RT);
    }
    const int vec_ok = (gw % 8 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)guide % 8 == 0);
    dim3 grid(strips, segs, batch);
    k_guided_coeff_s<RT, SNT, SR, SGR><<<grid, SNT, sm_c, st>>>(depth, w, h, guide, gw, gh, r, eps, seg, ctx->ab);
    k_guided_apply_s<RT, SNT, SR, SGR><<<grid, SNT, sm_a, st>>>(ctx->ab, guide, gw, gh, r, seg, vec_ok, out, q);
    V3D_LAUNCHED(ctx, 2);
    return V3D_OK;
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
    static const bool use_tiles = getenv("V3D_GUIDED_TILES") != nullptr;     // development A/B switch
    if (!use_tiles) {
        switch (r) {
            case 8: return launch_guided_stream<8>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
            case 4: return launch_guided_stream<4>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
            default: return launch_guided_stream<0>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
        }
    }
    switch (r) {
        case 8: return launch_guided_rt<8>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
        case 4: return launch_guided_rt<4>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
        default: return launch_guided_rt<0>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
    }
}
