// Horizontal path directions with TMA-staged rows.
// One warp walks one image row.  A row of C (and of S) is a single contiguous run of
// W1 * D * 2 bytes, so instead of one 128*NR-byte load per pixel step the warp's elected lane streams
// it through shared memory in CH-pixel chunks with cp.async.bulk (SASS: UBLKCP) completing on an
// mbarrier, NST chunks deep.  The left-to-right pass updates S in place in its shared-memory stage
// and writes the chunk back with a bulk store; the right-to-left pass only reads (S_total feeds the
// winner-takes-all directly and never returns to memory).
// Replaces the per-row part of OpenCV computeDisparitySGBM (depth.py:341).  Spec: SURVEY.md A.3/A.4.
#include "path_common.cuh"
#include "tma.cuh"

namespace {

constexpr int HW_WARPS = 8;   // warps (rows) per block
constexpr int CH = 8;         // pixel steps per staged chunk

// ------------------------------------------------------------------------------------------
// left -> right (predecessor x-1):  S += L
// ------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(HW_WARPS * 32)
k_path_lr_tma(const uint16_t* __restrict__ Cv, uint16_t* __restrict__ Sv, int W1, int rows, uint32_t P1p, uint32_t P2p)
{
    using VT = typename Vec<NR>::T;
    constexpr int NST = 3;
    constexpr int STEP_B = 128 * NR;                 // bytes of one pixel's D costs
    extern __shared__ __align__(128) unsigned char hsm[];
    __shared__ __align__(8) uint64_t bars[HW_WARPS][NST];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row = blockIdx.x * HW_WARPS + wib;
    if (row >= rows) return;
    unsigned char* cst = hsm + (size_t)wib * (2 * NST * CH * STEP_B);   // [NST][CH][STEP_B]
    unsigned char* sst = cst + NST * CH * STEP_B;
    const unsigned char* Cg = reinterpret_cast<const unsigned char*>(Cv) + (size_t)row * W1 * STEP_B;
    unsigned char* Sg = reinterpret_cast<unsigned char*>(Sv) + (size_t)row * W1 * STEP_B;
    const int nchunks = (W1 + CH - 1) / CH;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) mbar_init(&bars[wib][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int c) {        // lane 0 only
        const int st = c % NST;
        const int x0 = c * CH;
        const uint32_t bytes = (uint32_t)min(CH, W1 - x0) * STEP_B;
        mbar_expect_tx(&bars[wib][st], 2 * bytes);
        bulk_g2s(cst + st * CH * STEP_B, Cg + (size_t)x0 * STEP_B, bytes, &bars[wib][st]);
        bulk_g2s(sst + st * CH * STEP_B, Sg + (size_t)x0 * STEP_B, bytes, &bars[wib][st]);
    };
    if (lane == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }
    uint32_t M[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) M[r] = 0;
    for (int c = 0; c < nchunks; c++) {
        const int st = c % NST;
        mbar_wait(&bars[wib][st], (uint32_t)((c / NST) & 1));
        const VT* cs = reinterpret_cast<const VT*>(cst + st * CH * STEP_B) + lane;
        VT* ss = reinterpret_cast<VT*>(sst + st * CH * STEP_B) + lane;
        const int n = min(CH, W1 - c * CH);
        if (n == CH) {
#pragma unroll
            for (int j = 0; j < CH; j++) {
                uint32_t Cr[NR], Sr[NR], L[NR];
                unpack<NR>(cs[j * 32], Cr);
                unpack<NR>(ss[j * 32], Sr);
                path_step<NR>(M, Cr, L, P1p, P2p, lane);
#pragma unroll
                for (int r = 0; r < NR; r++) Sr[r] += L[r];
                ss[j * 32] = pack<NR>(Sr);
            }
        } else {
            for (int j = 0; j < n; j++) {
                uint32_t Cr[NR], Sr[NR], L[NR];
                unpack<NR>(cs[j * 32], Cr);
                unpack<NR>(ss[j * 32], Sr);
                path_step<NR>(M, Cr, L, P1p, P2p, lane);
#pragma unroll
                for (int r = 0; r < NR; r++) Sr[r] += L[r];
                ss[j * 32] = pack<NR>(Sr);
            }
        }
        fence_async_smem();          // generic-proxy writes of every lane -> visible to the bulk store
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(Sg + (size_t)c * CH * STEP_B, sst + st * CH * STEP_B, (uint32_t)n * STEP_B);
            bulk_commit();
            if (c + 2 < nchunks) {
                bulk_wait_read<1>();   // the store issued one chunk ago has finished reading stage (c+2)%NST
                issue(c + 2);
            }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------
// right -> left (predecessor x+1) fused with winner-takes-all.  Per pixel one 8-byte record:
//   .x = minS | best << 16      (best = 0xffff when the uniqueness test rejects the pixel)
//   .y = S[best-1] | S[best+1] << 16
// which k_select turns into the disparity (disp2 vote, sub-pixel, LR check).
// ------------------------------------------------------------------------------------------
template <int NR, bool TAP_S>
__device__ __forceinline__ void wta_step(uint32_t (&M)[NR], const typename Vec<NR>::T* cs, typename Vec<NR>::T* ss,
                                         typename Vec<NR>::T* Stap, uint32_t P1p, uint32_t P2p, int uniq, int lane,
                                         const uint32_t (&dc)[NR], const uint32_t (&idx)[NR], uint32_t* srow_w,
                                         int i, int x, uint2& myrec)
{
    constexpr int D = 64 * NR;
    uint32_t Cr[NR], Sr[NR], L[NR];
    unpack<NR>(*cs, Cr);
    unpack<NR>(*ss, Sr);
    path_step<NR>(M, Cr, L, P1p, P2p, lane);
#pragma unroll
    for (int r = 0; r < NR; r++) Sr[r] += L[r];
    if (TAP_S) Stap[(size_t)x * 32] = pack<NR>(Sr);
    // first argmin through (S << 8 | d) keys
    uint32_t key = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        key = min(key, __byte_perm(Sr[r], dc[r], 0x7104));
        key = min(key, __byte_perm(Sr[r], dc[r], 0x7325));
    }
    key = __reduce_min_sync(V3D_FULL_MASK, key);
    const uint32_t best = key & 0xffu, minS = key >> 8;
    // uniqueness: the smallest S over |d - best| > 1
    const uint32_t off = ((1u - best) & 0xffffu) * 0x10001u;      // t = d - best + 1 in each half
    uint32_t m2 = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const uint32_t t = __vadd2(idx[r], off);
        const uint32_t e = __vadd2(__vminu2(t, 0x00030003u), 0xfffdfffdu);   // 0xfffd..0xffff iff t in {0,1,2}
        m2 = __vminu2(m2, __vmaxu2(Sr[r], e));
    }
    m2 = __vminu2(m2, __byte_perm(m2, 0, 0x1032));
    const uint32_t minS2 = __reduce_min_sync(V3D_FULL_MASK, m2) & 0xffffu;
    const bool reject = minS2 * (uint32_t)(100 - uniq) < minS * 100u;
    // neighbours of the minimum through a per-warp shared row (double buffered by step parity)
    uint32_t* sr = srow_w + (i & 1) * (D / 2);
#pragma unroll
    for (int r = 0; r < NR; r++) sr[lane * NR + r] = Sr[r];
    __syncwarp();
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(sr);
    const uint32_t sm1 = s16[best > 0 ? best - 1 : 0];
    const uint32_t sp1 = s16[best < D - 1 ? best + 1 : D - 1];
    if (lane == (i & 31)) {
        myrec.x = (minS & 0xffffu) | ((reject ? 0xffffu : best) << 16);
        myrec.y = sm1 | (sp1 << 16);
    }
}

template <int NR, bool TAP_S>
__global__ void __launch_bounds__(HW_WARPS * 32)
k_path_rl_wta_tma(const uint16_t* __restrict__ Cv, uint16_t* __restrict__ Sv, uint2* __restrict__ rec, int W1,
                  int rows, uint32_t P1p, uint32_t P2p, int uniq)
{
    using VT = typename Vec<NR>::T;
    constexpr int D = 64 * NR;
    constexpr int NST = 3;
    constexpr int STEP_B = 128 * NR;
    extern __shared__ __align__(128) unsigned char hsm[];
    __shared__ __align__(8) uint64_t bars[HW_WARPS][NST];
    __shared__ uint32_t srow[HW_WARPS][2][D / 2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row = blockIdx.x * HW_WARPS + wib;
    if (row >= rows) return;
    unsigned char* cst = hsm + (size_t)wib * (2 * NST * CH * STEP_B);
    unsigned char* sst = cst + NST * CH * STEP_B;
    const unsigned char* Cg = reinterpret_cast<const unsigned char*>(Cv) + (size_t)row * W1 * STEP_B;
    const unsigned char* Sg = reinterpret_cast<const unsigned char*>(Sv) + (size_t)row * W1 * STEP_B;
    VT* Stap = reinterpret_cast<VT*>(Sv) + (size_t)row * W1 * 32 + lane;
    rec += (size_t)row * W1;
    const int nchunks = (W1 + CH - 1) / CH;

    uint32_t dc[NR], idx[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const uint32_t d0 = 2 * NR * lane + 2 * r;
        dc[r] = d0 | ((d0 + 1) << 8);
        idx[r] = d0 | ((d0 + 1) << 16);
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) mbar_init(&bars[wib][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // chunk c covers the pixels [lo, hi) counted from the right end of the row
    auto issue = [&](int c) {
        const int st = c % NST;
        const int hi = W1 - c * CH, lo = max(hi - CH, 0);
        const uint32_t bytes = (uint32_t)(hi - lo) * STEP_B;
        mbar_expect_tx(&bars[wib][st], 2 * bytes);
        bulk_g2s(cst + st * CH * STEP_B, Cg + (size_t)lo * STEP_B, bytes, &bars[wib][st]);
        bulk_g2s(sst + st * CH * STEP_B, Sg + (size_t)lo * STEP_B, bytes, &bars[wib][st]);
    };
    if (lane == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
        if (nchunks > 2) issue(2);
    }
    uint32_t M[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) M[r] = 0;
    uint2 myrec = make_uint2(0, 0);
    int i = 0;                                   // step counter, x = W1 - 1 - i
    for (int c = 0; c < nchunks; c++) {
        const int st = c % NST;
        mbar_wait(&bars[wib][st], (uint32_t)((c / NST) & 1));
        const int hi = W1 - c * CH, lo = max(hi - CH, 0), n = hi - lo;
        const VT* cs = reinterpret_cast<const VT*>(cst + st * CH * STEP_B) + lane;
        VT* ss = reinterpret_cast<VT*>(sst + st * CH * STEP_B) + lane;
        if (n == CH) {
#pragma unroll
            for (int j = CH - 1; j >= 0; j--, i++) {
                wta_step<NR, TAP_S>(M, cs + j * 32, ss + j * 32, Stap, P1p, P2p, uniq, lane, dc, idx, &srow[wib][0][0],
                                    i, lo + j, myrec);
                if ((i & 31) == 31) rec[W1 - 1 - i + (31 - lane)] = myrec;     // 32 records, coalesced
            }
        } else {
            for (int j = n - 1; j >= 0; j--, i++) {
                wta_step<NR, TAP_S>(M, cs + j * 32, ss + j * 32, Stap, P1p, P2p, uniq, lane, dc, idx, &srow[wib][0][0],
                                    i, lo + j, myrec);
                if ((i & 31) == 31) rec[W1 - 1 - i + (31 - lane)] = myrec;
            }
        }
        __syncwarp();                            // every lane is done reading this stage
        if (lane == 0 && c + NST < nchunks) issue(c + NST);
    }
    // flush the last partial group of records: steps i0 .. i-1 live in lanes 0 .. (i-1-i0)
    if (i & 31) {
        const int i0 = i & ~31;
        if (i0 + lane < i) rec[W1 - 1 - (i0 + lane)] = myrec;
    }
}

template <int NR>
int launch_h(v3d_ctx* ctx, int batch, cudaStream_t st, bool tap_s)
{
    const int rows = batch * ctx->H;
    const uint32_t P1p = (uint32_t)ctx->P1 * 0x10001u, P2p = (uint32_t)ctx->P2 * 0x10001u;
    const size_t smem = (size_t)HW_WARPS * 2 * 3 * CH * 128 * NR;
    dim3 grid((rows + HW_WARPS - 1) / HW_WARPS), block(HW_WARPS * 32);
    if (!(ctx->h_attr_set & (1 << NR))) {
        V3D_CUDA(cudaFuncSetAttribute(k_path_lr_tma<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        V3D_CUDA(cudaFuncSetAttribute(k_path_rl_wta_tma<NR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        V3D_CUDA(cudaFuncSetAttribute(k_path_rl_wta_tma<NR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->h_attr_set |= (1 << NR);
    }
    {
        V3dScope scope(ctx, ST_PATHS, st);
        k_path_lr_tma<NR><<<grid, block, smem, st>>>(ctx->C, ctx->S, ctx->W1, rows, P1p, P2p);
        V3D_LAUNCHED(ctx, 1);
    }
    {
        V3dScope scope(ctx, ST_WTA, st);
        if (tap_s)
            k_path_rl_wta_tma<NR, true><<<grid, block, smem, st>>>(ctx->C, ctx->S, ctx->rec, ctx->W1, rows, P1p, P2p, ctx->uniq);
        else
            k_path_rl_wta_tma<NR, false><<<grid, block, smem, st>>>(ctx->C, ctx->S, ctx->rec, ctx->W1, rows, P1p, P2p, ctx->uniq);
        V3D_LAUNCHED(ctx, 1);
    }
    return V3D_OK;
}

}  // namespace

int v3d_launch_paths_horizontal(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    const bool tap_s = ctx->debug_taps != 0;   // parity tests ask the WTA pass to also store S_total
    switch (ctx->D) {
        case 64: return launch_h<1>(ctx, batch, st, tap_s);
        case 128: return launch_h<2>(ctx, batch, st, tap_s);
        case 256: return launch_h<4>(ctx, batch, st, tap_s);
    }
    return v3d_fail(V3D_EINVAL, "numDisparities %d unsupported (64, 128, 256)", ctx->D);
}
