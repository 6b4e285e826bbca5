// Prefilter (clipped x-Sobel + intensity, BT half-pixel intervals) and the block-summed
// Birchfield-Tomasi cost volume C[b][y][x][d] (uint16, d fastest, x in window coordinates).
// Replaces the first third of cv2.StereoSGBM.compute (depth.py:341): OpenCV calcPixelCostBT plus the
// blockSize x blockSize box sum of computeDisparitySGBM.  Spec: SURVEY.md Appendix A.2.
#include "v3d_internal.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Prefilter: one uint2 record per pixel  {x: sobel v | lo<<8 | hi<<16,  y: intensity v | lo<<8 | hi<<16}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int gray_at(const uint8_t* img, size_t pitch, int W, int H, int x, int y)
{
    return __ldg(img + (size_t)min(max(y, 0), H - 1) * pitch + x);
}

__device__ __forceinline__ void prefilter_px(const uint8_t* img, size_t pitch, int W, int H, int x, int y,
                                             int ftzero, int& sob, int& inten)
{
    if (x <= 0 || x >= W - 1) { sob = ftzero; inten = ftzero; return; }   // border fill hits both channels
    const int g = 2 * (gray_at(img, pitch, W, H, x + 1, y) - gray_at(img, pitch, W, H, x - 1, y)) +
                  (gray_at(img, pitch, W, H, x + 1, y - 1) - gray_at(img, pitch, W, H, x - 1, y - 1)) +
                  (gray_at(img, pitch, W, H, x + 1, y + 1) - gray_at(img, pitch, W, H, x - 1, y + 1));
    sob = min(max(g, -ftzero), ftzero) + ftzero;
    inten = gray_at(img, pitch, W, H, x, y);
}

__global__ void __launch_bounds__(256)
k_prefilter(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right, size_t gpitch, size_t gstride,
            int W, int H, int ftzero, uint2* __restrict__ pfL, uint2* __restrict__ pfR)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int eye = blockIdx.z & 1, b = blockIdx.z >> 1;
    if (x >= W) return;
    const uint8_t* img = (eye ? right : left) + (size_t)b * gstride;
    int s[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int xx = x - 1 + i;
        if (xx < 0 || xx >= W) { s[i] = -1; t[i] = -1; }     // missing neighbour
        else prefilter_px(img, gpitch, W, H, xx, y, ftzero, s[i], t[i]);
    }
    auto interval = [](const int (&p)[3]) -> uint32_t {
        const int v = p[1];
        const int l = p[0] >= 0 ? (v + p[0]) >> 1 : v;
        const int r = p[2] >= 0 ? (v + p[2]) >> 1 : v;
        const int lo = min(v, min(l, r)), hi = max(v, max(l, r));
        return (uint32_t)v | ((uint32_t)lo << 8) | ((uint32_t)hi << 16);
    };
    uint2 rec = make_uint2(interval(s), interval(t));
    (eye ? pfR : pfL)[((size_t)b * H + y) * W + x] = rec;
}

// ---------------------------------------------------------------------------------------------
// Cost volume.  Block = a strip of TXW window columns (one warp per column, TX = TXW-2R of them are
// output columns, the rest are halo) sweeping a band of rows top-down.
//   lane l owns the disparity pairs d = 2l + 64k (+1), k < NR        (D = 64*NR)
//   vertical (2R+1)-row sum: sliding window in registers (ring of 2R+1 packed pix values)
//   horizontal (2R+1)-column sum: through shared memory, clamped in WINDOW coordinates
// The right-image row segment the strip needs is staged once per row into shared memory as packed
// int16 {v, -v, lo, -hi} quads in reversed column order, in two copies (even / odd start) so every
// lane's 128-bit load is aligned whatever the column parity.
// ---------------------------------------------------------------------------------------------
constexpr int TXW = 16;

template <int NR, int R>
struct CostSmem {
    static constexpr int D = 64 * NR;
    static constexpr int NWORDS = D / 2 + 8;
    uint4 rbuf[2][2][2][NWORDS];   // [row parity][channel][copy][word] = {v, -v, lo, -hi} pairs
    uint4 lbuf[2][TXW][2];         // [row parity][column][channel]   = {u, -u, lo, -hi} duplicated in both halves
    uint32_t vbuf[2][TXW][D / 2];  // [row parity][column][pair]      = vertical sums
};

__device__ __forceinline__ uint32_t neg16(uint32_t v) { return (0x10000u - v) & 0xffffu; }

// Row staging is split in two so that the global load of row r+2 is in flight while row r is being
// computed: stage_load() only issues the load, stage_store() (one iteration later) expands the record
// into the shared-memory layout.
template <int NR, int R>
__device__ __forceinline__ uint2 stage_load(const uint2* __restrict__ pfL, const uint2* __restrict__ pfR, int W, int xs,
                                            int tid)
{
    constexpr int D = 64 * NR;
    constexpr int NE = TXW + D - 1;
    if (tid < NE) {
        const int Xhi = xs - R + (TXW - 1) + D;
        return __ldg(pfR + min(max(Xhi - tid, 0), W - 1));
    }
    if (tid < NE + TXW) return __ldg(pfL + min(max(xs - R + (tid - NE) + D, 0), W - 1));
    return make_uint2(0, 0);
}

template <int NR, int R>
__device__ __forceinline__ void stage_store(CostSmem<NR, R>& sm, int buf, const uint2 rec, int tid)
{
    constexpr int D = 64 * NR;
    constexpr int NE = TXW + D - 1;
    if (tid < NE) {
        const int q = tid;
        uint16_t* base = reinterpret_cast<uint16_t*>(&sm.rbuf[buf][0][0][0]);
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            const uint32_t w = ch ? rec.y : rec.x;
            const uint32_t v = w & 0xff, lo = (w >> 8) & 0xff, hi = (w >> 16) & 0xff;
            const uint16_t q4[4] = { (uint16_t)v, (uint16_t)neg16(v), (uint16_t)lo, (uint16_t)neg16(hi) };
            // copy 0 (even start): element q;  copy 1 (odd start): element q-1
#pragma unroll
            for (int cp = 0; cp < 2; cp++) {
                const int e = q - cp;
                if (e < 0) continue;
                uint16_t* p = base + ((size_t)(ch * 2 + cp) * CostSmem<NR, R>::NWORDS + (e >> 1)) * 8 + (e & 1);
#pragma unroll
                for (int k = 0; k < 4; k++) p[2 * k] = q4[k];
            }
        }
    } else if (tid < NE + TXW) {
        const int c = tid - NE;
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            const uint32_t w = ch ? rec.y : rec.x;
            const uint32_t v = w & 0xff, lo = (w >> 8) & 0xff, hi = (w >> 16) & 0xff;
            sm.lbuf[buf][c][ch] = make_uint4(v * 0x10001u, neg16(v) * 0x10001u, lo * 0x10001u, neg16(hi) * 0x10001u);
        }
    }
}

template <int NR, int R>
__global__ void __launch_bounds__(TXW * 32)
k_cost(const uint2* __restrict__ pfL, const uint2* __restrict__ pfR, uint32_t* __restrict__ C,
       int W, int H, int W1, int band_h)
{
    constexpr int D = 64 * NR;
    constexpr int K = 2 * R + 1;
    constexpr int TX = TXW - 2 * R;
    __shared__ CostSmem<NR, R> sm;

    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5;
    const int b = blockIdx.z;
    const int xs = blockIdx.x * TX;
    const int x = xs - R + c;                       // window column of this warp
    const bool valid_col = (x >= 0 && x < W1);
    const bool inner = (c >= R && c < TXW - R && x < W1);
    const int y0 = blockIdx.y * band_h, y1 = min(H, y0 + band_h);
    const int ystart = y0 - R, yend = y1 + R;
    pfL += (size_t)b * H * W;
    pfR += (size_t)b * H * W;

    const int q0 = TXW - 1 - c;
    const int copy = q0 & 1;
    const int wbase = (q0 - copy) >> 1;

    uint32_t ring[K][NR];
    uint32_t V[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) {
        V[k] = 0;
#pragma unroll
        for (int i = 0; i < K; i++) ring[i][k] = 0;
    }

    auto load_row = [&](int r) -> uint2 {
        const int rr = min(max(r, 0), H - 1);
        return stage_load<NR, R>(pfL + (size_t)rr * W, pfR + (size_t)rr * W, W, xs, tid);
    };
    uint2 rec = load_row(ystart);
    stage_store<NR, R>(sm, 0, rec, tid);
    rec = load_row(ystart + 1);
    __syncthreads();

    for (int row = ystart; row < yend; row += K) {
#pragma unroll
        for (int ph = 0; ph < K; ph++) {
            const int r = row + ph;
            if (r >= yend) break;
            const int cur = (r - ystart) & 1;
            if (r + 1 < yend) {
                stage_store<NR, R>(sm, cur ^ 1, rec, tid);    // row r+1, loaded one iteration ago
                rec = load_row(r + 2);                        // in flight while row r is computed
            }
            if (valid_col) {
                const uint4 ls = sm.lbuf[cur][c][0];
                const uint4 li = sm.lbuf[cur][c][1];
#pragma unroll
                for (int k = 0; k < NR; k++) {
                    const int w = wbase + lane + 32 * k;
                    const uint4 rs = sm.rbuf[cur][0][copy][w];
                    const uint4 ri = sm.rbuf[cur][1][copy][w];
                    // {x: v, y: -v, z: lo, w: -hi}
                    uint32_t c0 = __vimax_s16x2_relu(__vadd2(ls.x, rs.w), __vadd2(rs.z, ls.y));
                    uint32_t c1 = __vimax_s16x2_relu(__vadd2(rs.x, ls.w), __vadd2(ls.z, rs.y));
                    const uint32_t bs = __vminu2(c0, c1);
                    c0 = __vimax_s16x2_relu(__vadd2(li.x, ri.w), __vadd2(ri.z, li.y));
                    c1 = __vimax_s16x2_relu(__vadd2(ri.x, li.w), __vadd2(li.z, ri.y));
                    const uint32_t bi = __vminu2(c0, c1);
                    const uint32_t pix = bs + ((bi >> 2) & 0x3fff3fffu);
                    V[k] = V[k] + pix - ring[ph][k];     // halves never borrow: V includes ring[ph]
                    ring[ph][k] = pix;
                    sm.vbuf[cur][c][lane + 32 * k] = V[k];
                }
            }
            __syncthreads();
            const int yo = r - R;
            if (inner && yo >= y0) {
                uint32_t* out = C + (((size_t)b * H + yo) * W1 + x) * (D / 2);
#pragma unroll
                for (int k = 0; k < NR; k++) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int dx = -R; dx <= R; dx++) {
                        const int cn = min(max(x + dx, 0), W1 - 1) - (xs - R);
                        acc += sm.vbuf[cur][cn][lane + 32 * k];
                    }
                    out[lane + 32 * k] = acc;
                }
            }
        }
    }
}

template <int NR>
int launch_cost_r(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    const int band_h = 128;
    const uint2 *pfL = ctx->pfL, *pfR = ctx->pfR;
    uint32_t* C = reinterpret_cast<uint32_t*>(ctx->C);
    const int W = ctx->W, H = ctx->H, W1 = ctx->W1;
    dim3 block(TXW * 32);
#define V3D_COST_CASE(RR)                                                                              \
    case RR: {                                                                                         \
        dim3 grid((W1 + (TXW - 2 * RR) - 1) / (TXW - 2 * RR), (H + band_h - 1) / band_h, batch);        \
        k_cost<NR, RR><<<grid, block, 0, st>>>(pfL, pfR, C, W, H, W1, band_h);                          \
        break;                                                                                         \
    }
    switch (ctx->R) {
        V3D_COST_CASE(0)
        V3D_COST_CASE(1)
        V3D_COST_CASE(2)
        V3D_COST_CASE(3)
        default: return v3d_fail(V3D_EINVAL, "blockSize %d unsupported", ctx->p.blockSize);
    }
#undef V3D_COST_CASE
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

}  // namespace

int v3d_launch_prefilter(v3d_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t gpitch,
                         size_t gstride, int batch, cudaStream_t st)
{
    V3dScope scope(ctx, ST_PREFILTER, st);
    dim3 grid((ctx->W + 255) / 256, ctx->H, batch * 2);
    k_prefilter<<<grid, 256, 0, st>>>(left, right, gpitch, gstride, ctx->W, ctx->H, ctx->ftzero, ctx->pfL, ctx->pfR);
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

int v3d_launch_cost(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    V3dScope scope(ctx, ST_COST, st);
    switch (ctx->D) {
        case 64: return launch_cost_r<1>(ctx, batch, st);
        case 128: return launch_cost_r<2>(ctx, batch, st);
        case 256: return launch_cost_r<4>(ctx, batch, st);
    }
    return v3d_fail(V3D_EINVAL, "numDisparities %d unsupported (64, 128, 256)", ctx->D);
}
