// Prefilter (clipped x-Sobel + intensity, BT half-pixel intervals) and the block-summed
// Birchfield-Tomasi cost volume C[b][y][x][d] (uint16, d fastest, x in window coordinates).
// Replaces the first third of cv2.StereoSGBM.compute (depth.py:341): OpenCV calcPixelCostBT plus the
// blockSize x blockSize box sum of computeDisparitySGBM.  Spec: SURVEY.md Appendix A.2.
//
// Two kernels:
//   k_prefilter_expand  writes, once per pixel, the operands of the BT cost in the exact layout the cost
//                       kernel's lanes consume them (packed int16 pairs, right image column-reversed, both
//                       pair alignments), so that
//   k_cost              stages a row with five cp.async.bulk copies (TMA, mbarrier-completed, 3 rows deep)
//                       and spends its instructions on the cost arithmetic only.
#include "v3d_internal.h"
#include "tma.cuh"

#include <type_traits>

namespace {

#ifndef V3D_COST_HN
#define V3D_COST_HN 4
#endif
#ifndef V3D_COST_HOIST_WAITS
#define V3D_COST_HOIST_WAITS 1
#endif
constexpr int TXW = 32;       // window columns per block (one warp each); TXW - 2R of them are output columns
constexpr int PADL = 32;      // front padding (elements) of the reversed right-image rows
constexpr int LPAD = 4;       // front padding (columns) of the left-image rows: a strip's halo may start R columns left of column 0 (minDisparity < 0)

// ---------------------------------------------------------------------------------------------
// Prefilter record of one pixel: {x: sobel v | lo<<8 | hi<<16,  y: intensity v | lo<<8 | hi<<16}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int gray_at(const uint8_t* img, size_t pitch, int H, int x, int y)
{
    return __ldg(img + (size_t)min(max(y, 0), H - 1) * pitch + x);
}

__device__ __forceinline__ void prefilter_px(const uint8_t* img, size_t pitch, int W, int H, int x, int y,
                                             int ftzero, int& sob, int& inten)
{
    if (x <= 0 || x >= W - 1) { sob = ftzero; inten = ftzero; return; }   // border fill hits both channels
    const int g = 2 * (gray_at(img, pitch, H, x + 1, y) - gray_at(img, pitch, H, x - 1, y)) +
                  (gray_at(img, pitch, H, x + 1, y - 1) - gray_at(img, pitch, H, x - 1, y - 1)) +
                  (gray_at(img, pitch, H, x + 1, y + 1) - gray_at(img, pitch, H, x - 1, y + 1));
    sob = min(max(g, -ftzero), ftzero) + ftzero;
    inten = gray_at(img, pitch, H, x, y);
}

// BT half-pixel interval record {v | lo << 8 | hi << 16} of one channel from the pixel and its two neighbours
// (-1 = neighbour outside the image)
__device__ __forceinline__ uint32_t bt_interval(int pm, int v, int pp)
{
    const int l = pm >= 0 ? (v + pm) >> 1 : v;
    const int r = pp >= 0 ? (v + pp) >> 1 : v;
    const int lo = min(v, min(l, r)), hi = max(v, max(l, r));
    return (uint32_t)v | ((uint32_t)lo << 8) | ((uint32_t)hi << 16);
}

// prefiltered pixel as sobel | intensity << 8, or 0xffff outside the image
__device__ __forceinline__ uint32_t prefilter_packed(const uint8_t* img, size_t pitch, int W, int H, int x, int y, int ftzero)
{
    if (x < 0 || x >= W) return 0xffffu;
    int sob, inten;
    prefilter_px(img, pitch, W, H, x, y, ftzero, sob, inten);
    return (uint32_t)sob | ((uint32_t)inten << 8);
}

// record of the pixel whose packed value is `c`, with packed neighbours `m` (x-1) and `p` (x+1); zero outside
__device__ __forceinline__ uint2 rec_from_packed(uint32_t m, uint32_t c, uint32_t p)
{
    if (c == 0xffffu) return make_uint2(0, 0);
    const int sm = m == 0xffffu ? -1 : (int)(m & 0xff), im = m == 0xffffu ? -1 : (int)(m >> 8);
    const int sp = p == 0xffffu ? -1 : (int)(p & 0xff), ip = p == 0xffffu ? -1 : (int)(p >> 8);
    return make_uint2(bt_interval(sm, (int)(c & 0xff), sp), bt_interval(im, (int)(c >> 8), ip));
}

__device__ __forceinline__ uint32_t neg16(uint32_t v) { return (0x10000u - v) & 0xffffu; }

// {v, -v, lo, -hi} of two elements packed as int16 pairs (e0 in the low halves)
__device__ __forceinline__ uint4 pack_quads(uint32_t w0, uint32_t w1)
{
    const uint32_t v0 = w0 & 0xff, l0 = (w0 >> 8) & 0xff, h0 = (w0 >> 16) & 0xff;
    const uint32_t v1 = w1 & 0xff, l1 = (w1 >> 8) & 0xff, h1 = (w1 >> 16) & 0xff;
    return make_uint4(v0 | (v1 << 16), neg16(v0) | (neg16(v1) << 16), l0 | (l1 << 16), neg16(h0) | (neg16(h1) << 16));
}

// Right image: rexp[b][y][ch][cp][w] (uint4) = quads of the reversed-order elements (2w+cp, 2w+cp+1), where
//              element e <-> image column W-1-(e-PADL).
// Left image:  lexp[b][y][LPAD + X][ch] (uint4) = {u, -u, lo, -hi} duplicated in both halves, -LPAD <= X < W + TXW (clamped).
__global__ void __launch_bounds__(256)
k_prefilter_expand(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right, size_t gpitch, size_t gstride,
                   int W, int H, int ftzero, uint4* __restrict__ rexp, int wpw, uint4* __restrict__ lexp)
{
    // Every pixel is prefiltered once per block into shared memory (the 3-pixel intervals of neighbouring
    // outputs overlap: per-thread evaluation costs 4x the loads and Sobel arithmetic).
    constexpr int NRC = 2 * 256 + 4;               // right-image columns a block touches: xmax+1 down to xmax-514
    constexpr int NLC = 256 + 2;                   // left-image columns: lb .. lb+257, lb = min(t0-1, W-2)
    __shared__ uint16_t pr[NRC], pl[NLC];
    const int t0 = blockIdx.x * 256, t = t0 + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    const uint8_t* imgR = right + (size_t)b * gstride;
    const uint8_t* imgL = left + (size_t)b * gstride;
    const int xmax = W - 1 - (2 * t0 - PADL);      // column of element 2*t0 (the block's right-most one)
    const int lb = min(max(t0 - LPAD, 0) - 1, W - 2);   // first staged left column (element t <-> column clamp(t - LPAD)); blocks past the image only repeat its last column
    if (t0 < wpw)
        for (int i = threadIdx.x; i < NRC; i += 256) pr[i] = (uint16_t)prefilter_packed(imgR, gpitch, W, H, xmax + 1 - i, y, ftzero);
    if (t0 < W + TXW + LPAD)
        for (int i = threadIdx.x; i < NLC; i += 256) pl[i] = (uint16_t)prefilter_packed(imgL, gpitch, W, H, lb + i, y, ftzero);
    __syncthreads();
    if (t < wpw) {
        // elements 2t, 2t+1, 2t+2 = columns xr0, xr0-1, xr0-2 with xr0 = xmax - 2*threadIdx.x; column x sits at pr[xmax+1-x]
        const int i0 = 1 + 2 * threadIdx.x;
        uint2 rec[3];
#pragma unroll
        for (int i = 0; i < 3; i++) rec[i] = rec_from_packed(pr[i0 + i + 1], pr[i0 + i], pr[i0 + i - 1]);
        uint4* base = rexp + ((size_t)(b * H + y) * 4) * wpw + t;
        base[0 * (size_t)wpw] = pack_quads(rec[0].x, rec[1].x);    // channel 0 (sobel), copy 0: elements 2t, 2t+1
        base[1 * (size_t)wpw] = pack_quads(rec[1].x, rec[2].x);    // channel 0, copy 1: elements 2t+1, 2t+2
        base[2 * (size_t)wpw] = pack_quads(rec[0].y, rec[1].y);    // channel 1 (intensity)
        base[3 * (size_t)wpw] = pack_quads(rec[1].y, rec[2].y);
    }
    if (t < W + TXW + LPAD) {
        const int i = min(max(t - LPAD, 0), W - 1) - lb;            // columns outside the image repeat the border one
        const uint2 rec = rec_from_packed(pl[i - 1], pl[i], pl[i + 1]);
        uint4* o = lexp + ((size_t)(b * H + y) * (W + TXW + LPAD) + t) * 2;
        o[0] = pack_quads(rec.x, rec.x);
        o[1] = pack_quads(rec.y, rec.y);
    }
}

// ---------------------------------------------------------------------------------------------
// Cost volume.  Block = a strip of TXW window columns (one warp per column) sweeping a band of rows.
//   lane l owns the disparity pairs d = 2l + 64k (+1), k < NR        (D = 64*NR)
//   vertical (2R+1)-row sum: sliding window in registers (ring of 2R+1 packed pix values)
//   horizontal (2R+1)-column sum: through shared memory, clamped in WINDOW coordinates; a warp then owns a
//   run of 4 output columns of one row and slides the window over the 4 + 2R column sums it loaded once
// ---------------------------------------------------------------------------------------------
// The row loop is unrolled by U = max(K, 3) rows and the kernel keeps U staged rows and U vertical-sum
// buffers, so that the ring slot (ph % K), the staging buffer and the vbuf slot of a row are all
// compile-time constants inside the unrolled body: no stage/parity bookkeeping instructions.
template <int NR, int R>
struct CostSmem {
    static constexpr int D = 64 * NR;
    static constexpr int K = 2 * R + 1;
    static constexpr int RPB = (R == 2 && NR <= 2) ? 2 : 1;      // rows per block barrier
    static constexpr int U = RPB > 1 ? 2 * K : (K >= 3 ? K : 3);
    static constexpr int NWORDS = D / 2 + TXW / 2 + 2;
    uint4 rbuf[U][2][2][NWORDS];      // [stage][channel][copy][word] = {v, -v, lo, -hi} pairs of the right image
    uint4 lbuf[U][TXW][2];            // [stage][column][channel]     = {u, -u, lo, -hi} of the left image
    uint32_t vbuf[U][TXW][D / 2];     // [slot][column][pair]         = vertical sums
    uint64_t bar[U];                  // TMA completion, one per staged row
    uint64_t gbar[U / RPB];           // "every warp has written its vertical sums of this row group" (32 arrivals)
};

// compile-time unrolled loop: f(std::integral_constant<int, I>) for I in [0, N)
template <int N, int I = 0, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
}

__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar_addr), "r"(parity) : "memory");
}

template <int NR, int R, bool PAD>
__global__ void __launch_bounds__(TXW * 32)
k_cost(const uint4* __restrict__ rexp, int wpw, const uint4* __restrict__ lexp, uint32_t* __restrict__ C,
       int W, int H, int W1, int band_h, int Dreal, int x0, int minD)
{
    using SM = CostSmem<NR, R>;
    constexpr int D = 64 * NR;
    constexpr int K = SM::K, U = SM::U, RPB = SM::RPB;
    constexpr int TX = TXW - 2 * R;
    constexpr int NWORDS = SM::NWORDS;
    extern __shared__ __align__(128) unsigned char cost_smem[];
    SM& sm = *reinterpret_cast<SM*>(cost_smem);

    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5;
    const int b = blockIdx.z;
    const int xs = blockIdx.x * TX;
    const int y0 = blockIdx.y * band_h, y1 = min(H, y0 + band_h);
    const int ystart = y0 - R;
    const int nrows = (y1 + R - ystart + U - 1) / U * U;    // padded to whole unrolled groups
    const int yend = ystart + nrows;

    // reversed element index of the strip's right-most image column, and the first staged word of each copy
    // image column = window column + x0 (x0 = minD + Dreal, clamped at 0); disparity index d means minD + d pixels;
    // d in [Dreal, D) is padding.  PADL covers the strip's halo columns for every minD: W - 1 - Xhi + minD >= -29.
    const int Xhi = xs - R + (TXW - 1) + x0;
    const int gbase = (W - 1 - Xhi + minD) + PADL;
    const int wlo0 = gbase >> 1, wlo1 = (gbase - 1) >> 1;
    // this warp's first element, its pair alignment and its first word inside the staged copy
    const int e0 = gbase + (TXW - 1 - c);
    const int copy = e0 & 1;
    const int wrel = ((e0 - copy) >> 1) - (copy ? wlo1 : wlo0);
    const int Xl0 = xs - R + x0;                    // image column of warp 0 in the left image

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < U; s++) mbar_init(&sm.bar[s], 1);
#pragma unroll
        for (int s = 0; s < U / RPB; s++) mbar_init(&sm.gbar[s], TXW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // one thread: stage image row clamp(r) into stage st.  The five bulk copies of a row are issued in two halves
    // (part 0: the sobel-channel copies and the tx expectation, part 1: the intensity copies and the left row)
    // so that two different warps can share the ~80 serial instructions; part < 0 = everything
    auto issue = [&](int r, int st, int part) {
        const int rr = min(max(r, 0), H - 1);
        const uint4* rrow = rexp + ((size_t)(b * H + rr) * 4) * wpw;
        constexpr uint32_t RB = NWORDS * 16, LB = TXW * 2 * 16;
        if (part != 1) {
            mbar_expect_tx(&sm.bar[st], 4 * RB + LB);
            bulk_g2s(&sm.rbuf[st][0][0][0], rrow + 0 * (size_t)wpw + wlo0, RB, &sm.bar[st]);
            bulk_g2s(&sm.rbuf[st][0][1][0], rrow + 1 * (size_t)wpw + wlo1, RB, &sm.bar[st]);
        }
        if (part != 0) {
            bulk_g2s(&sm.rbuf[st][1][0][0], rrow + 2 * (size_t)wpw + wlo0, RB, &sm.bar[st]);
            bulk_g2s(&sm.rbuf[st][1][1][0], rrow + 3 * (size_t)wpw + wlo1, RB, &sm.bar[st]);
            bulk_g2s(&sm.lbuf[st][0][0], lexp + ((size_t)(b * H + rr) * (W + TXW + LPAD) + Xl0 + LPAD) * 2, LB, &sm.bar[st]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < U - RPB; s++) issue(ystart + s, s, -1);   // U-RPB rows in flight
    }

    uint32_t ring[K][NR];
    uint32_t V[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) {
        V[k] = 0;
#pragma unroll
        for (int i = 0; i < K; i++) ring[i][k] = 0;
    }
    // per-thread shared-memory addresses (32-bit) and the clamped neighbour columns of the horizontal sum
    const uint32_t bar0 = smem_u32(&sm.bar[0]);
    const uint4* rs_p = &sm.rbuf[0][0][copy][wrel + lane];
    const uint4* ri_p = &sm.rbuf[0][1][copy][wrel + lane];
    constexpr int RSTG = 2 * 2 * NWORDS;            // uint4 per stage
    const uint4* l_p = &sm.lbuf[0][c][0];
    constexpr int LSTG = TXW * 2;
    uint32_t* v_p = &sm.vbuf[0][c][lane];
    constexpr int VSTG = TXW * (D / 2);
    // Horizontal sum: warp c owns one (row of the group, run of HN output columns, register k) item.  The
    // HN + 2R columns a run touches are loaded once and the window slides in registers.
    constexpr int HN = V3D_COST_HN, NRUN = (TX + HN - 1) / HN, NITEM = RPB * NRUN * NR;
    static_assert(NITEM <= TXW, "one horizontal item per warp");
    const int h_s = c / (NRUN * NR), h_run = (c % (NRUN * NR)) / NR, h_k = c % NR;
    const int h_x = xs + h_run * HN;                               // first output column (window coordinates)
    const int h_n = min(HN, min(xs + TX, W1) - h_x);               // output columns of this item (<= 0: none)
    const bool h_flat = xs - R >= 0 && xs - R + TXW <= W1;         // no clamping anywhere in this strip
    const uint32_t* h_vb = &sm.vbuf[0][0][32 * h_k + lane];
    uint32_t* h_out = C + (((ptrdiff_t)b * H + (ystart - R)) * W1 + h_x) * (D / 2) + 32 * h_k + lane;
    const size_t out_row = (size_t)W1 * (D / 2);
    const bool h_real = !PAD || 2 * (lane + 32 * h_k) < Dreal;     // padded disparities get a cost no real one can reach

    const uint32_t gbar0 = smem_u32(&sm.gbar[0]);
    // Phase 1 of a row group: the pixel costs of its RPB rows, the running vertical sums, their store for the
    // horizontal pass, and this warp's arrival on the group's barrier.  pgc is a compile-time constant.
    auto cost_rows = [&](auto pgc, uint32_t par) {
        constexpr int pg = decltype(pgc)::value;
#if V3D_COST_HOIST_WAITS
#pragma unroll
        for (int s = 0; s < RPB; s++) mbar_wait_a(bar0 + (pg + s) * 8, par);    // both rows' copies first: the loads of the second row may then overlap the arithmetic of the first
#endif
#pragma unroll
        for (int s = 0; s < RPB; s++) {
            const int ph = pg + s;
#if !V3D_COST_HOIST_WAITS
            mbar_wait_a(bar0 + ph * 8, par);
#endif
            const uint4 ls = l_p[ph * LSTG + 0];
            const uint4 li = l_p[ph * LSTG + 1];
#pragma unroll
            for (int k = 0; k < NR; k++) {
                const uint4 rs = rs_p[ph * RSTG + 32 * k];
                const uint4 ri = ri_p[ph * RSTG + 32 * k];
                // {x: v, y: -v, z: lo, w: -hi}
                uint32_t c0 = __vimax_s16x2_relu(__vadd2(ls.x, rs.w), __vadd2(rs.z, ls.y));
                uint32_t c1 = __vimax_s16x2_relu(__vadd2(rs.x, ls.w), __vadd2(ls.z, rs.y));
                const uint32_t bs = __vminu2(c0, c1);
                c0 = __vimax_s16x2_relu(__vadd2(li.x, ri.w), __vadd2(ri.z, li.y));
                c1 = __vimax_s16x2_relu(__vadd2(ri.x, li.w), __vadd2(li.z, ri.y));
                const uint32_t bi = __vminu2(c0, c1);
                const uint32_t pix = bs + ((bi >> 2) & 0x3fff3fffu);
                V[k] = V[k] + pix - ring[ph % K][k];     // halves never borrow: V includes the ring slot
                ring[ph % K][k] = pix;
                v_p[ph * VSTG + 32 * k] = V[k];
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gbar0 + (pg / RPB) * 8) : "memory");
    };

    // The block barrier of a group is split: a warp arrives after its cost rows, runs the cost rows of the NEXT
    // group, and only then waits for the others and does its share of the horizontal pass -- warps drift by up to
    // one group instead of meeting in lock-step every RPB rows.
    cost_rows(std::integral_constant<int, 0>{}, 0u);
    uint32_t parity = 0;
    for (int row = ystart; row < yend; row += U, parity ^= 1) {
        static_for<U / RPB>([&](auto pgic) {
            constexpr int pgi = decltype(pgic)::value, pg = pgi * RPB;
            if constexpr (pgi + 1 < U / RPB) cost_rows(std::integral_constant<int, (pgi + 1) * RPB>{}, parity);
            else if (row + U < yend) cost_rows(std::integral_constant<int, 0>{}, parity ^ 1u);
            mbar_wait_a(gbar0 + pgi * 8, parity);   // vbuf slots of this group complete; everyone is done with the previous group's stages
            // Refill the stages the rows of the PREVIOUS group used (U - RPB rows ahead of this one).  Issuing a
            // row is ~80 serial instructions of one thread: it goes, in two halves, to warps that own no
            // horizontal item, so that no warp is systematically later than the others at the next group barrier.
#pragma unroll
            for (int s = 0; s < RPB; s++) {
                const int r = row + pg + s;
                if (lane == 0 && r + U - RPB < yend) {
                    if (c == (NITEM + 2 * s) % TXW) issue(r + U - RPB, (pg + s + U - RPB) % U, 0);
                    if (c == (NITEM + 2 * s + 1) % TXW) issue(r + U - RPB, (pg + s + U - RPB) % U, 1);
                }
            }
            if (c < NITEM && h_n > 0) {
                const int ph = pg + h_s;                    // (h_s is warp-uniform, so is everything below)
                const int yo = row + ph - R;
                if (yo >= y0 && yo < y1) {
                    const uint32_t* vb = h_vb + ph * VSTG;
                    uint32_t v[HN + 2 * R];
                    if (h_flat && h_n == HN) {
#pragma unroll
                        for (int i = 0; i < HN + 2 * R; i++) v[i] = vb[(h_run * HN + i) * (D / 2)];
                    } else {                                // strip on the window border, or a short last run
#pragma unroll
                        for (int i = 0; i < HN + 2 * R; i++)
                            v[i] = i < h_n + 2 * R ? vb[(min(max(h_x - R + i, 0), W1 - 1) - (xs - R)) * (D / 2)] : 0u;
                    }
                    uint32_t acc = 0;
#pragma unroll
                    for (int i = 0; i <= 2 * R; i++) acc += v[i];
                    uint32_t* o = h_out + (size_t)ph * out_row;
#pragma unroll
                    for (int j = 0; j < HN; j++) {
                        if (j > 0) acc += v[j + 2 * R] - v[j - 1];          // halves never borrow: acc includes v[j-1]
                        // padded disparities: they never win a minimum and their path state saturates at P2,
                        // i.e. they act like the missing neighbour d = D
                        if (j < h_n) o[j * (D / 2)] = h_real ? acc : 0x20002000u;
                    }
                }
            }
        });
        h_out += (size_t)U * out_row;
    }
}

template <int NR, int R>
int launch_cost_nr(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    constexpr int U = CostSmem<NR, R>::U;
    const int band_h = U * ((130 + U - 1) / U) - 2 * R;   // band_h + 2R is a whole number of unrolled groups
    const size_t smem = sizeof(CostSmem<NR, R>) + 128;
    const bool pad = ctx->D != ctx->Dk;
    const int key = (NR * 8 + R) * 2 + (pad ? 1 : 0);
    auto kern = pad ? k_cost<NR, R, true> : k_cost<NR, R, false>;
    if (!(ctx->cost_attr_set & (1ull << key))) {
        V3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->cost_attr_set |= (1ull << key);
    }
    const int W1 = ctx->W1, H = ctx->H;
    dim3 grid((W1 + (TXW - 2 * R) - 1) / (TXW - 2 * R), (H + band_h - 1) / band_h, batch), block(TXW * 32);
    kern<<<grid, block, smem, st>>>(ctx->rexp, ctx->rexp_wpw, ctx->lexp, reinterpret_cast<uint32_t*>(ctx->C),
                                    ctx->W, H, W1, band_h, ctx->D, ctx->x0, ctx->minD);
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

template <int NR>
int launch_cost_r(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    switch (ctx->R) {
        case 0: return launch_cost_nr<NR, 0>(ctx, batch, st);
        case 1: return launch_cost_nr<NR, 1>(ctx, batch, st);
        case 2: return launch_cost_nr<NR, 2>(ctx, batch, st);
        case 3: return launch_cost_nr<NR, 3>(ctx, batch, st);
    }
    return v3d_fail(V3D_EINVAL, "blockSize %d unsupported", ctx->p.blockSize);
}

}  // namespace

// words per (row, channel, copy) of the expanded right image: the row itself, the front padding and room
// for the widest staged copy (D = 256) to read past the last column
int v3d_rexp_words(int W) { return (W + PADL) / 2 + 160; }
int v3d_lexp_cols(int W) { return W + TXW + LPAD; }

int v3d_launch_prefilter(v3d_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t gpitch,
                         size_t gstride, int batch, cudaStream_t st)
{
    V3dScope scope(ctx, ST_PREFILTER, st);
    const int n = ctx->rexp_wpw > ctx->W + TXW + LPAD ? ctx->rexp_wpw : ctx->W + TXW + LPAD;
    dim3 grid((n + 255) / 256, ctx->H, batch);
    k_prefilter_expand<<<grid, 256, 0, st>>>(left, right, gpitch, gstride, ctx->W, ctx->H, ctx->ftzero, ctx->rexp,
                                             ctx->rexp_wpw, ctx->lexp);
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

int v3d_launch_cost(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    V3dScope scope(ctx, ST_COST, st);
    switch (ctx->Dk) {
        case 64: return launch_cost_r<1>(ctx, batch, st);
        case 128: return launch_cost_r<2>(ctx, batch, st);
        case 256: return launch_cost_r<4>(ctx, batch, st);
    }
    return v3d_fail(V3D_EINVAL, "numDisparities %d unsupported (64, 128, 256)", ctx->D);
}
