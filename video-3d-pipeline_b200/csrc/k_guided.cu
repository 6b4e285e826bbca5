// Guided-filter depth upscale (colour guide = the 4K frame).
// Replaces upscale.py:47-59, where the reference merely lets ffmpeg `scale` the depth PNGs; the
// guided filter its readme promises (readme.md:97,119) does not exist upstream, so the normative
// definition is oracle/guided.py (He/Sun/Tang colour guided filter on a bilinearly upsampled depth).
//
// Two kernels per frame, both streaming: a CTA owns a strip of TW output columns (NT = TW + 2r region
// columns, one thread each) and walks down a segment of rows.
//   k_guided_coeff_s : moments of (I, p) -> 3x3 solve -> per-pixel (a0, a1, a2, b)        [float4 plane]
//   k_guided_apply_s : box mean of (a, b) -> q = a.I + b -> uint16
// Vertical box sums live in registers as running sums (enter the new row, leave the row 2r+1 above,
// which a ring of rows in shared memory remembers); every R rows the column sums go through shared
// memory once for the horizontal pass (first output of a run summed directly, the rest slid).
//   * guide moments are kept in BYTE units about an integer centre: the I and I.I sums are integers
//     below 2^24, so fp32 adds them exactly -- no drift however long the strip;
//   * the planes that involve the depth (p, I.p, and a, b in the second kernel) are compensated
//     (Kahan) running sums, so their error stays that of one direct (2r+1)-term sum.
// Global loads are asynchronous: guide rows / coefficient rows of the NEXT row group are cp.async'ed
// into shared memory while the horizontal pass of the current group runs.
#include "v3d_internal.h"

namespace {

constexpr int GRMAX = 16;   // largest supported radius (8 and 4 have compile-time instantiations, the rest take the run-time one)
#ifndef V3D_GUIDED_DEPTH_PREFETCH
#define V3D_GUIDED_DEPTH_PREFETCH 0
#endif
#ifndef V3D_GUIDED_APPLY_HOIST
#define V3D_GUIDED_APPLY_HOIST 1
#endif

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

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void sub4(float4& a, const float4& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }

struct RowTap { int gy, o0, o1; float fy; };   // guide row; element offsets of the two depth rows; weight

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

// Pitch (slots) of one row of column sums: column x sits at x + x/GR, so a warp whose lanes take consecutive runs
// reads consecutive slots mod 8; making the pitch congruent to the number of runs keeps that true across rows.
template <int RT, int NT, int GR>
__host__ __device__ constexpr int row_pitch()
{
    const int runs = ((NT - 2 * (RT > 0 ? RT : GRMAX)) & ~7) / GR;
    int vp = NT + NT / GR + 1;
    while ((vp - (GR + 1) * runs) % 8 != 0) vp++;      // a run occupies GR + 1 slots
    return vp;
}

// The coefficient kernel's horizontal pass splits every run of 8 outputs over a lane pair when it can (radius 8,
// one row of the group per warp).
template <int RT, int NT, int R, int GR>
__host__ __device__ constexpr bool paired_pass() { return RT == 8 && GR == 8 && NT % (32 * R) == 0; }

// shared-memory bytes of k_guided_coeff_s
template <int RT, int NT, int R, int GR>
constexpr size_t coeff_s_smem(int win)
{
    return (size_t)3 * R * row_pitch<RT, NT, GR>() * 16 + (size_t)R * row_pitch<RT, NT, GR>() * 4 + (size_t)win * NT * 8 +
           (size_t)2 * R * (NT * 3 + 16);
}

template <int RT, int NT, int R, int GR>
__global__ void __launch_bounds__(NT, (R >= 8 ? 256 : 512) / NT)
k_guided_coeff_s(const uint16_t* __restrict__ depth, int w, int h, const uint8_t* __restrict__ guide, int gw, int gh,
                 int r_arg, float eps, int seg, float4* __restrict__ ab)
{
    extern __shared__ float4 gsm[];
    const int r = RT > 0 ? RT : r_arg, win = 2 * r + 1;
    const int TW = (NT - 2 * r) & ~7;
    constexpr int VP = row_pitch<RT, NT, GR>();                     // slots per (plane group, row): column x at x + x/GR
    constexpr int GSB = NT * 3 + 16;                                // staged guide row: NT pixels of rgb (+ pad), bytes
    float4* vbuf = gsm;                                             // [3 plane groups][R][VP] float4 (sums 0..11)
    float* v12 = reinterpret_cast<float*>(gsm + 3 * R * VP);        // [R][VP] float (sum 12)
    uint2* ring = reinterpret_cast<uint2*>(v12 + R * VP);           // [win][NT] {rgb bytes, p - centre}
    uint8_t* gst = reinterpret_cast<uint8_t*>(ring + win * NT);     // [2][R][GSB] guide rows of the next group
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
    // interior strips stage whole guide rows with 4-byte cp.async; strips that touch the reflected border load bytes
    const bool staged = (gw & 3) == 0 && ((X0 - r) & 3) == 0 && X0 - r >= 0 && X0 - r + NT <= gw &&
                        (reinterpret_cast<uintptr_t>(guide) & 3) == 0;

    const int gx = reflect_idx(X0 - r + tid, gw);
    int x0, x1;
    float fx;
    axis_tap(gx, w, gw, x0, x1, fx);
    const float ax0 = (1.0f - fx) * s16, ax1 = fx * s16;
    const uint16_t* const dx0 = depth + x0;
    const uint16_t* const dx1 = depth + x1;
    auto depth_at = [&](const RowTap& rt) -> float {                // bilinear depth tap, branch free
        const float top = (float)__ldg(dx0 + rt.o0) * ax0 + (float)__ldg(dx1 + rt.o0) * ax1;
        const float bot = (float)__ldg(dx0 + rt.o1) * ax0 + (float)__ldg(dx1 + rt.o1) * ax1;
        return top * (1.0f - rt.fy) + bot * rt.fy;
    };
    auto make_tap = [&](int j) -> RowTap {
        RowTap t;
        int y0, y1;
        t.gy = reflect_idx(Y0 - r + j, gh);
        axis_tap(t.gy, h, gh, y0, y1, t.fy);
        t.o0 = y0 * w;
        t.o1 = y1 * w;
        return t;
    };
    auto stage_rows = [&](int grp) {                                // guide rows of group `grp` -> gst[grp & 1]
        if (staged && tid < NT * 3 / 4) {
            const RowTap* tp = taps[grp & 1];
#pragma unroll
            for (int jj = 0; jj < R; jj++) {
                const uint8_t* src = guide + ((size_t)tp[jj].gy * gw + (X0 - r)) * 3 + tid * 4;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(gst + ((grp & 1) * R + jj) * GSB + tid * 4);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            }
        }
    };
    if (tid < R) taps[0][tid] = make_tap(tid);
    if (tid == NT / 2) {                                            // strip centre: column X0 + TW/2, middle row
        RowTap c;
        int y0, y1;
        c.gy = min(Y0 + seg_h / 2, gh - 1);
        axis_tap(c.gy, h, gh, y0, y1, c.fy);
        c.o0 = y0 * w;
        c.o1 = y1 * w;
        const uint8_t* gp = guide + ((size_t)c.gy * gw + gx) * 3;
        centre_rgb = (uint32_t)__ldg(gp) | ((uint32_t)__ldg(gp + 1) << 8) | ((uint32_t)__ldg(gp + 2) << 16);
        centre_p = depth_at(c);
    }
    __syncthreads();
    stage_rows(0);
    const uint32_t crgb = centre_rgb;
    const float cp = centre_p;
    const float cb0 = 8388608.0f + (float)(crgb & 0xff), cb1 = 8388608.0f + (float)((crgb >> 8) & 0xff),
                cb2 = 8388608.0f + (float)((crgb >> 16) & 0xff);
    for (int k = 0; k < win; k++) ring[k * NT + tid] = make_uint2(crgb, 0u);   // centre pixel: all moments 0
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    float V[13], C[4];
#pragma unroll
    for (int q = 0; q < 13; q++) V[q] = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) C[q] = 0.0f;
    uint2* rp = ring + tid;
    uint2* const rend = ring + win * NT;
    const int wslot = tid + tid / GR;
    const int runs = TW / GR;
    const float inv_n = 1.0f / (float)(win * win);
    const uint8_t* gcol = guide + (size_t)gx * 3;
    const int sword = (tid * 3) >> 2;                                // staged row: word holding this column's first byte
    const uint32_t ssel = 0x3210u + 0x1111u * ((tid * 3) & 3);       // PRMT selector picking its 3 bytes

    for (int g = 0; g * R < nrows; g++) {
        const RowTap* tp = taps[g & 1];
        const bool full = g * R + R <= nrows;
        // one row of the column: enter it into the running sums, leave the row 2r+1 above
        auto row_step = [&](int jj, uint32_t rgb, float p) {
            const float ap = p - cp;
            const uint2 old = *rp;
            *rp = make_uint2(rgb, __float_as_uint(ap));
            rp += NT;
            if (rp >= rend) rp = ring + tid;
            const float a0 = byte_magic<0>(rgb) - cb0, a1 = byte_magic<1>(rgb) - cb1, a2 = byte_magic<2>(rgb) - cb2;
            const float b0 = byte_magic<0>(old.x) - cb0, b1 = byte_magic<1>(old.x) - cb1, b2 = byte_magic<2>(old.x) - cb2;
            const float bp = __uint_as_float(old.y);
            // exact planes (integers): I, I.I
            V[0] += a0 - b0; V[1] += a1 - b1; V[2] += a2 - b2;
            // (integers below 2^24 at every intermediate step, so two fused multiply-adds give the same bits as
            // product difference + add, with one instruction less per plane)
            V[7] = fmaf(-b0, b0, fmaf(a0, a0, V[7]));   V[8] = fmaf(-b0, b1, fmaf(a0, a1, V[8]));   V[9] = fmaf(-b0, b2, fmaf(a0, a2, V[9]));
            V[10] = fmaf(-b1, b1, fmaf(a1, a1, V[10])); V[11] = fmaf(-b1, b2, fmaf(a1, a2, V[11])); V[12] = fmaf(-b2, b2, fmaf(a2, a2, V[12]));
            // depth planes: compensated
            kahan(V[3], C[0], ap - bp);
            kahan(V[4], C[1], fmaf(a0, ap, -(b0 * bp)));
            kahan(V[5], C[2], fmaf(a1, ap, -(b1 * bp)));
            kahan(V[6], C[3], fmaf(a2, ap, -(b2 * bp)));
            if (g * R + jj >= 2 * r) {
                float4* vr = vbuf + jj * VP + wslot;
                vr[0] = make_float4(V[0], V[1], V[2], V[3]);
                vr[R * VP] = make_float4(V[4], V[5], V[6], V[7]);
                vr[2 * R * VP] = make_float4(V[8], V[9], V[10], V[11]);
                v12[jj * VP + wslot] = V[12];
            }
        };
        auto guide_at = [&](int jj) -> uint32_t {
            if (staged) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(gst + ((g & 1) * R + jj) * GSB) + sword;
                return __byte_perm(wp[0], wp[1], ssel);
            }
            const uint8_t* gp = gcol + (size_t)tp[jj].gy * gw * 3;
            return (uint32_t)__ldg(gp) | ((uint32_t)__ldg(gp + 1) << 8) | ((uint32_t)__ldg(gp + 2) << 16);
        };
        if (full) {
            float pv[R];
            uint32_t gv[R];
#pragma unroll
            for (int jj = 0; jj < R; jj++) { pv[jj] = depth_at(tp[jj]); gv[jj] = guide_at(jj); }
#pragma unroll
            for (int jj = 0; jj < R; jj++) row_step(jj, gv[jj], pv[jj]);
        } else {
            for (int jj = 0; g * R + jj < nrows; jj++) row_step(jj, guide_at(jj), depth_at(tp[jj]));
        }
        if (tid < R) taps[(g + 1) & 1][tid] = make_tap((g + 1) * R + tid);
        __syncthreads();
        stage_rows(g + 1);                                           // lands while the horizontal pass runs
#if V3D_GUIDED_DEPTH_PREFETCH
        if ((g + 1) * R < nrows) {                                   // the next group's depth taps: into L1 behind the horizontal pass
            const RowTap* tn = taps[(g + 1) & 1];
#pragma unroll
            for (int jj = 0; jj < R; jj++) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(dx0 + tn[jj].o0));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(dx0 + tn[jj].o1));
            }
        }
#endif

        // horizontal pass + 3x3 solve
        const float kn = k255 * inv_n, kkn = k255 * k255 * inv_n;
        const float cI0 = (float)(crgb & 0xff) * k255 - 0.5f, cI1 = (float)((crgb >> 8) & 0xff) * k255 - 0.5f,
                    cI2 = (float)((crgb >> 16) & 0xff) * k255 - 0.5f;
        // window sums acc[13] of one output pixel -> (a0, a1, a2, b)
        auto solve = [&](const float (&acc)[13]) -> float4 {
            const float mI0 = acc[0] * kn, mI1 = acc[1] * kn, mI2 = acc[2] * kn, mp = acc[3] * inv_n;
            const float c0 = acc[4] * kn - mI0 * mp, c1 = acc[5] * kn - mI1 * mp, c2 = acc[6] * kn - mI2 * mp;
            const float s00 = acc[7] * kkn - mI0 * mI0 + eps, s01 = acc[8] * kkn - mI0 * mI1, s02 = acc[9] * kkn - mI0 * mI2;
            const float s11 = acc[10] * kkn - mI1 * mI1 + eps, s12 = acc[11] * kkn - mI1 * mI2, s22 = acc[12] * kkn - mI2 * mI2 + eps;
            // (Sigma + eps I) a = cov(I, p) by LDL^T: the matrix is symmetric positive definite (pivots >= eps)
            // and for nearly collinear colour channels -- the usual case -- this stays accurate where the
            // adjugate / determinant form loses digits like (lambda_max / eps)^2.  Reciprocals: approx + one
            // Newton step (~1 ulp).
            auto rcp = [](float x) -> float {
                float r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
                return r * fmaf(-x, r, 2.0f);
            };
            const float i0 = rcp(s00);
            const float l1 = s01 * i0, l2 = s02 * i0;
            const float d1 = fmaf(-l1, s01, s11), e1 = fmaf(-l1, s02, s12);
            const float i1 = rcp(d1);
            const float l21 = e1 * i1;
            const float d2 = fmaf(-l21, e1, fmaf(-l2, s02, s22));
            const float i2 = rcp(d2);
            const float y1 = fmaf(-l1, c0, c1);
            const float y2 = fmaf(-l21, y1, fmaf(-l2, c0, c2));
            const float a2 = y2 * i2;
            const float a1 = fmaf(-l21, a2, y1 * i1);
            const float a0 = fmaf(-l2, a2, fmaf(-l1, a1, c0 * i0));
            // b referred to a guide centred at 0.5:  q = a.(I - 0.5) + b
            const float bb = (mp + cp) - a0 * (mI0 + cI0) - a1 * (mI1 + cI1) - a2 * (mI2 + cI2);
            return make_float4(a0, a1, a2, bb);
        };
        if constexpr (paired_pass<RT, NT, R, GR>()) {
            // One row of the group per warp; lanes l and l + 16 share a run of 8 outputs (4 each).  Their two first windows
            // (17 columns, 4 apart) overlap in 13 columns: each lane sums half of the overlap plus its own 4 columns, the
            // halves are exchanged with one shuffle per plane, then each lane slides its window over its 4 outputs.
            // 28 of 32 lanes work, against 56 of 128 threads when one thread owns a whole run.
            constexpr int WPR = NT / (32 * R);                      // warps per row of the group (1 at 128 threads)
            const int lane = tid & 31, jj = (tid >> 5) / WPR;
            const int rpw = (runs + WPR - 1) / WPR;                 // runs per warp (<= 16: one per lane of a half-warp)
            const int o = g * R + jj - 2 * RT;                      // output row of this warp in this group (warp-uniform)
            if (o >= 0 && o < seg_h) {
                const int half = lane >> 4;
                const int run_raw = (lane & 15) < rpw ? ((tid >> 5) % WPR) * rpw + (lane & 15) : runs;
                const int run = min(run_raw, runs - 1), xb = run * GR;      // lanes past the last run repeat it, store nothing
                const float4* vr = vbuf + jj * VP + xb + run;           // slot of the run's first column
                const float* vs = v12 + jj * VP + xb + run;
                float acc[13], part[13], m[13];
                auto ld13 = [&](int da, int db) {                       // column offset from the run start: da (lanes 0..15), db (16..31)
                    const int e = half ? db + db / GR : da + da / GR;
                    const float4 g0 = vr[e], g1 = vr[e + R * VP], g2 = vr[e + 2 * R * VP];
                    m[0] = g0.x; m[1] = g0.y; m[2] = g0.z; m[3] = g0.w;
                    m[4] = g1.x; m[5] = g1.y; m[6] = g1.z; m[7] = g1.w;
                    m[8] = g2.x; m[9] = g2.y; m[10] = g2.z; m[11] = g2.w;
                    m[12] = vs[e];
                };
#pragma unroll
                for (int q = 0; q < 13; q++) { acc[q] = 0.0f; part[q] = 0.0f; }
#pragma unroll
                for (int t = 0; t < 4; t++) {                           // own columns: 0..3 / 17..20
                    ld13(t, 17 + t);
#pragma unroll
                    for (int q = 0; q < 13; q++) acc[q] += m[q];
                }
                const float w6 = half ? 0.0f : 1.0f;                    // the upper half has only 6 shared columns
#pragma unroll
                for (int t = 0; t < 7; t++) {                           // half of the shared columns: 4..10 / 11..16
                    ld13(4 + t, t < 6 ? 11 + t : 11);
#pragma unroll
                    for (int q = 0; q < 13; q++) part[q] = t < 6 ? part[q] + m[q] : fmaf(m[q], w6, part[q]);   // x * 1 + y rounds like x + y
                }
#pragma unroll
                for (int q = 0; q < 13; q++) acc[q] += part[q] + __shfl_xor_sync(0xffffffffu, part[q], 16);
                const size_t orow = (size_t)(Y0 + o) * gw;
#pragma unroll
                for (int o2 = 0; o2 < 4; o2++) {
                    if (o2 > 0) {
                        ld13(o2 + 2 * RT, 4 + o2 + 2 * RT);
#pragma unroll
                        for (int q = 0; q < 13; q++) acc[q] += m[q];
                        ld13(o2 - 1, 4 + o2 - 1);
#pragma unroll
                        for (int q = 0; q < 13; q++) acc[q] -= m[q];
                    }
                    const int X = X0 + xb + 4 * half + o2;
                    const float4 c4 = solve(acc);
                    if (run_raw < runs && X < gw) ab[orow + X] = c4;
                }
            }
        } else {
        // item = (row of the group, run of GR output columns)
        for (int it = tid; it < R * runs; it += NT) {
            const int jj = it / runs, run = it - jj * runs;
            const int o = g * R + jj - 2 * r;
            const int xb = run * GR;
            if (o < 0 || o >= seg_h || X0 + xb >= gw) continue;
            const float4* vr = vbuf + jj * VP + xb + run;          // window start; xb is a multiple of GR
            const float* vs = v12 + jj * VP + xb + run;
            float acc[13], m[13];
            auto ld13 = [&](int dx) {                               // dx = offset from the window start
                const int e = RT > 0 ? dx + dx / GR : (xb + dx) + (xb + dx) / GR - xb - run;
                const float4 g0 = vr[e], g1 = vr[e + R * VP], g2 = vr[e + 2 * R * VP];
                m[0] = g0.x; m[1] = g0.y; m[2] = g0.z; m[3] = g0.w;
                m[4] = g1.x; m[5] = g1.y; m[6] = g1.z; m[7] = g1.w;
                m[8] = g2.x; m[9] = g2.y; m[10] = g2.z; m[11] = g2.w;
                m[12] = vs[e];
            };
#pragma unroll
            for (int q = 0; q < 13; q++) acc[q] = 0.0f;
            if (RT > 0) {
                const float4* v0 = vr;
                const float* s0 = vs;
#pragma unroll 1
                for (int t0 = 0; t0 + GR <= 2 * RT + 1; t0 += GR, vr += GR + 1, vs += GR + 1) {   // whole runs: the pad advances with them
#pragma unroll
                    for (int t = 0; t < GR; t++) {
                        ld13(t);
#pragma unroll
                        for (int q = 0; q < 13; q++) acc[q] += m[q];
                    }
                }
                vr = v0;
                vs = s0;
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
                ab[orow + X] = solve(acc);
            }
        }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
    }
}

// VEC: gw % 8 == 0 and 16-byte aligned rows -- 8 outputs leave as one 16-byte store, their guide bytes arrive as 3x8 bytes
template <int RT, int NT, int R, int GR, bool VEC>
__global__ void __launch_bounds__(NT)
k_guided_apply_s(const float4* __restrict__ ab, const uint8_t* __restrict__ guide, int gw, int gh, int r_arg, int seg,
                 uint16_t* __restrict__ out, float* __restrict__ qout)
{
    extern __shared__ float4 gsm[];
    const int r = RT > 0 ? RT : r_arg, win = 2 * r + 1;
    const int TW = (NT - 2 * r) & ~7;
    constexpr int VP = row_pitch<RT, NT, GR>();
    const int RING = win + R;            // rows of (a, b) kept: the window plus the group being prefetched
    float4* vbuf = gsm;                  // [R][VP] column sums of the group
    float4* ring = gsm + R * VP;         // [RING][NT], row j at slot j % RING, filled by cp.async
    __shared__ int rows[2][R];
    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * seg, b = blockIdx.z;
    const int seg_h = min(seg, gh - Y0), nrows = seg_h + 2 * r;
    ab += (size_t)b * gw * gh;
    guide += (size_t)b * gw * gh * 3;
    out += (size_t)b * gw * gh;
    if (!VEC && qout) qout += (size_t)b * gw * gh;
    const int gx = reflect_idx(X0 - r + tid, gw);
    int fslot = 0;                       // ring slot of the first row of the next group to fetch
    auto fetch_rows = [&](int grp) {     // this column's entries of group `grp` -> ring (own column only: no barrier needed)
        const int* rw = rows[grp & 1];
#pragma unroll
        for (int jj = 0; jj < R; jj++) {
            const int j = grp * R + jj;
            if (j < nrows) {
                const float4* src = ab + (size_t)rw[jj] * gw + gx;
                const int slot = fslot + jj >= RING ? fslot + jj - RING : fslot + jj;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + slot * NT + tid);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        fslot = fslot + R >= RING ? fslot + R - RING : fslot + R;
    };
    if (tid < R) rows[0][tid] = reflect_idx(Y0 - r + tid, gh);
    for (int k = R; k < RING; k++) ring[k * NT + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    fetch_rows(0);
    float4 V = make_float4(0.f, 0.f, 0.f, 0.f), C = V;
    const int runs = TW / GR;
    const float inv_n = 1.0f / (float)(win * win), k255 = 1.0f / 255.0f;
    const int wslot = tid + tid / GR;
    int nslot = 0, oslot = R;            // slot of the entering row j and of the leaving row j - win (= j + R mod RING)
    // Horizontal pass: item = (row of the group, run of GR output columns); R * runs <= TW < NT, so a thread owns at most
    // one item per group, the same (row, run) in every group.
    static_assert(R <= GR, "one horizontal item per thread and group");
    const int h_jj = tid / runs, h_run = tid - h_jj * runs, h_xb = h_run * GR;
    const bool h_mine = tid < R * runs && X0 + h_xb < gw;

    for (int g = 0; g * R < nrows; g++) {
#if V3D_GUIDED_APPLY_HOIST
        // the item's 24 guide bytes are requested HERE, a whole vertical pass and a block barrier ahead of their use
        // (they were 36 % of this kernel's stall samples when requested next to the window sums)
        uint2 u0 = make_uint2(0u, 0u), u1 = u0, u2 = u0;
        {
            const int o = g * R + h_jj - 2 * r;
            if (VEC && h_mine && o >= 0 && o < seg_h) {
                const uint2* gp = reinterpret_cast<const uint2*>(guide + ((size_t)(Y0 + o) * gw + X0 + h_xb) * 3);
                u0 = __ldg(gp); u1 = __ldg(gp + 1); u2 = __ldg(gp + 2);
            }
        }
#endif
        asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < R; jj++) {
            const int j = g * R + jj;
            if (j >= nrows) break;
            const float4 nv = ring[nslot * NT + tid];
            const float4 old = ring[oslot * NT + tid];              // zero while j < win (slots R.. start zeroed)
            nslot = nslot + 1 == RING ? 0 : nslot + 1;
            oslot = oslot + 1 == RING ? 0 : oslot + 1;
            kahan(V.x, C.x, nv.x - old.x); kahan(V.y, C.y, nv.y - old.y);
            kahan(V.z, C.z, nv.z - old.z); kahan(V.w, C.w, nv.w - old.w);
            if (j >= 2 * r) vbuf[jj * VP + wslot] = V;
        }
        if (tid < R) rows[(g + 1) & 1][tid] = reflect_idx(Y0 - r + (g + 1) * R + tid, gh);
        __syncthreads();
        fetch_rows(g + 1);               // lands while the horizontal pass runs

        const int o = g * R + h_jj - 2 * r;
        if (h_mine && o >= 0 && o < seg_h) {
            const int jj = h_jj, run = h_run;
            const int xb = h_xb, X = X0 + xb;
            const int Y = Y0 + o;
            const float4* vr = vbuf + jj * VP + xb + run;          // window start; xb is a multiple of GR
            auto at = [&](int dx) -> float4 { return vr[RT > 0 ? dx + dx / GR : (xb + dx) + (xb + dx) / GR - xb - run]; };
            const size_t base = (size_t)Y * gw + X;
#if !V3D_GUIDED_APPLY_HOIST
            // the run's 24 guide bytes are requested before the window sums so that their latency hides behind them
            uint2 u0 = make_uint2(0u, 0u), u1 = u0, u2 = u0;
            if (VEC) {
                const uint2* gp = reinterpret_cast<const uint2*>(guide + base * 3);
                u0 = __ldg(gp); u1 = __ldg(gp + 1); u2 = __ldg(gp + 2);
            }
#endif
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (RT > 0) {
#pragma unroll
                for (int t = 0; t <= 2 * RT; t++) add4(acc, at(t));
            } else {
#pragma unroll 1
                for (int t = 0; t <= 2 * r; t++) add4(acc, at(t));
            }
            if (VEC) {
                static_assert(!VEC || GR == 8, "vector path writes runs of 8");
                const uint32_t gb[6] = {u0.x, u0.y, u1.x, u1.y, u2.x, u2.y};
                uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int o2 = 0; o2 < GR; o2++) {
                    if (o2 > 0) { add4(acc, at(o2 + 2 * r)); sub4(acc, at(o2 - 1)); }
                    auto byte_at = [&](int k) -> float {
                        return __uint_as_float(__byte_perm(gb[k >> 2], 0x4b000000u, 0x7540 | (k & 3))) - 8388608.0f;
                    };
                    const float I0 = byte_at(3 * o2), I1 = byte_at(3 * o2 + 1), I2 = byte_at(3 * o2 + 2);
                    const float q = (acc.x * fmaf(I0, k255, -0.5f) + acc.y * fmaf(I1, k255, -0.5f) + acc.z * fmaf(I2, k255, -0.5f) + acc.w) * inv_n;
                    const float qc = fminf(fmaxf(q, 0.0f), 1.0f);
                    packed[o2 >> 1] |= (uint32_t)floorf(qc * 65535.0f + 0.5f) << ((o2 & 1) * 16);
                }
                *reinterpret_cast<uint4*>(out + base) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            } else {
#pragma unroll
                for (int o2 = 0; o2 < GR; o2++) {
                    if (o2 > 0) { add4(acc, at(o2 + 2 * r)); sub4(acc, at(o2 - 1)); }
                    if (X + o2 >= gw) break;
                    const uint8_t* gp = guide + (base + o2) * 3;
                    const float I0 = (float)__ldg(gp), I1 = (float)__ldg(gp + 1), I2 = (float)__ldg(gp + 2);
                    const float q = (acc.x * fmaf(I0, k255, -0.5f) + acc.y * fmaf(I1, k255, -0.5f) + acc.z * fmaf(I2, k255, -0.5f) + acc.w) * inv_n;
                    const float qc = fminf(fmaxf(q, 0.0f), 1.0f);
                    out[base + o2] = (uint16_t)floorf(qc * 65535.0f + 0.5f);
                    if (qout) qout[base + o2] = q;
                }
            }
        }
        __syncthreads();
    }
}

#ifndef V3D_GUIDED_ANT
#define V3D_GUIDED_ANT 128
#endif
constexpr int SNT = V3D_GUIDED_ANT, SR = 4, SGR = 8;      // SNT: threads (= region columns) per CTA of the apply kernel
#ifndef V3D_GUIDED_CGR
#define V3D_GUIDED_CGR 8
#endif
#ifndef V3D_GUIDED_CNT
#define V3D_GUIDED_CNT 128
#endif
constexpr int SNC = V3D_GUIDED_CNT;      // threads (= region columns) per CTA of the coefficient kernel
constexpr int SGC = V3D_GUIDED_CGR;      // outputs per horizontal run of the coefficient kernel (4 spreads the pass over all four warps but costs more shared-memory traffic: measured slower)

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
    const size_t sm_c = coeff_s_smem<RT, SNC, SR, SGC>(win);
    constexpr int SRA = 8;            // rows per group of the apply kernel
    const size_t sm_a = (size_t)SRA * row_pitch<RT, SNT, SGR>() * 16 + (size_t)(win + SRA) * SNT * 16;
    const size_t sm_a_max = (size_t)SRA * row_pitch<RT, SNT, SGR>() * 16 + (size_t)(2 * GRMAX + 1 + SRA) * SNT * 16;
    if (!(ctx->guided_attr_set & (1 << RT))) {
        V3D_CUDA(cudaFuncSetAttribute(k_guided_coeff_s<RT, SNC, SR, SGC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)coeff_s_smem<RT, SNC, SR, SGC>(2 * GRMAX + 1)));
        V3D_CUDA(cudaFuncSetAttribute(k_guided_apply_s<RT, SNT, SRA, SGR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a_max));
        V3D_CUDA(cudaFuncSetAttribute(k_guided_apply_s<RT, SNT, SRA, SGR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a_max));
        ctx->guided_attr_set |= (1 << RT);
    }
    const bool vec = q == nullptr && (gw % 8 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)guide % 8 == 0);
    dim3 grid(strips, segs, batch);
    const int TWC = (SNC - 2 * r) & ~7;
    dim3 grid_c((gw + TWC - 1) / TWC, segs, batch);
    {
        V3dScope scope(ctx, ST_GUIDED, st);
        k_guided_coeff_s<RT, SNC, SR, SGC><<<grid_c, SNC, sm_c, st>>>(depth, w, h, guide, gw, gh, r, eps, seg, ctx->ab);
    }
    V3dScope scope(ctx, ST_GUIDED_APPLY, st);
    if (vec) k_guided_apply_s<RT, SNT, SRA, SGR, true><<<grid, SNT, sm_a, st>>>(ctx->ab, guide, gw, gh, r, seg, out, q);
    else k_guided_apply_s<RT, SNT, SRA, SGR, false><<<grid, SNT, sm_a, st>>>(ctx->ab, guide, gw, gh, r, seg, out, q);
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
    switch (r) {
        case 8: return launch_guided_stream<8>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
        case 4: return launch_guided_stream<4>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
        default: return launch_guided_stream<0>(ctx, depth, w, h, guide, gw, gh, batch, r, eps, out, q, st);
    }
}
