// Horizontal path directions with TMA-staged rows.
// One warp walks one image row.  A row of C (and of S) is a single contiguous run of
// W1 * D * 2 bytes, so instead of one 128*NR-byte load per pixel step the warp's elected lane streams
// it through shared memory in CH-pixel chunks with cp.async.bulk (SASS: UBLKCP) completing on an
// mbarrier, NST chunks deep.  Neither horizontal direction writes a volume: the left-to-right pass only
// leaves a checkpoint of its path state per chunk, the right-to-left pass (last kernel of the chain) re-runs
// the left-to-right steps of each chunk from its checkpoint, adds both directions onto the vertical
// directions' S in shared memory and feeds the winner-takes-all directly -- S_total never returns to memory.
// Replaces the per-row part of OpenCV computeDisparitySGBM (depth.py:341).  Spec: SURVEY.md A.3/A.4.
#include "path_common.cuh"
#include "tma.cuh"

namespace {

constexpr int HW_WARPS = 8;   // warps (rows) per block
constexpr int CH = 8;         // pixel steps per staged chunk
#ifndef V3D_WTA_NST
#define V3D_WTA_NST 2       // staged chunks per warp of the last path kernel (2: three blocks per SM; 3 measured 3 % slower)
#endif
__host__ __device__ constexpr int wta_stages(int nr) { return nr == 4 ? 3 : V3D_WTA_NST; }   // D = 256: one block per SM either way, deeper staging wins

// ------------------------------------------------------------------------------------------
// left -> right (predecessor x-1), CHECKPOINT pass.  The left-to-right costs L are never written: this kernel
// only streams C in (one volume pass) and keeps, for every CH-pixel chunk of the right-to-left kernel below, the
// path state M that ENTERS the chunk (1/CH of a volume).  The right-to-left kernel re-runs the CH left-to-right
// steps of a chunk from its checkpoint, on the C chunk it has staged anyway, so the left-to-right direction
// costs 1 + 2/CH volume passes instead of the 3 of "write S, read S back" -- the recomputed steps land on a
// kernel that is HBM-bound and has the issue slots.
//   ckpt[row][c][d], c = chunk index counted from the RIGHT end of the row (chunk c = pixels [W1-8c-8, W1-8c)),
//   = M before the chunk's first pixel; the left-most chunk starts a path (M = 0) and has no checkpoint.
// ------------------------------------------------------------------------------------------
#ifndef V3D_LRC_NST
#define V3D_LRC_NST 4
#endif
__host__ __device__ constexpr int lrc_stages(int nr) { return nr == 4 ? 3 : V3D_LRC_NST; }   // D = 256: two blocks per SM
template <int NR>
__global__ void __launch_bounds__(HW_WARPS * 32)
k_path_lr_ckpt(const uint16_t* __restrict__ Cv, uint16_t* __restrict__ ckpt, int W1, int rows, uint32_t P1p, uint32_t P2p)
{
    using VT = typename Vec<NR>::T;
    constexpr int NST = lrc_stages(NR);
    constexpr int STEP_B = 128 * NR;                 // bytes of one pixel's D costs
    extern __shared__ __align__(128) unsigned char hsm[];
    __shared__ __align__(8) uint64_t bars[HW_WARPS][NST];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row = blockIdx.x * HW_WARPS + wib;
    if (row >= rows) return;
    unsigned char* cst = hsm + (size_t)wib * (NST * CH * STEP_B);   // [NST][CH][STEP_B]
    const unsigned char* Cg = reinterpret_cast<const unsigned char*>(Cv) + (size_t)row * W1 * STEP_B;
    const int nchunks = (W1 + CH - 1) / CH;
    VT* ck = reinterpret_cast<VT*>(ckpt) + (size_t)row * nchunks * 32 + lane;
    const int jstar = W1 & (CH - 1);                 // position inside a left-aligned chunk where a right-aligned chunk starts
    const int c_of_k0 = (W1 / CH) - 1;               // right-aligned chunk that starts inside left-aligned chunk 0

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
        mbar_expect_tx(&bars[wib][st], bytes);
        bulk_g2s(cst + st * CH * STEP_B, Cg + (size_t)x0 * STEP_B, bytes, &bars[wib][st]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST - 1; s++)
            if (s < nchunks) issue(s);
    }
    uint32_t M[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) M[r] = 0;
#if V3D_STEP_CARRY
    uint32_t mm = 0;                                 // M holds L, mm its minimum over d (path_step_carry)
    auto step = [&](const uint32_t (&Cr)[NR]) { path_step_carry<NR>(M, mm, Cr, P1p, P2p, lane); };
    auto checkpoint = [&](VT* dst) { uint32_t Ms[NR]; carry_to_state<NR>(M, mm, P2p, Ms); *dst = pack<NR>(Ms); };
#else
    auto step = [&](const uint32_t (&Cr)[NR]) { uint32_t L[NR]; path_step<NR>(M, Cr, L, P1p, P2p, lane); };
    auto checkpoint = [&](VT* dst) { *dst = pack<NR>(M); };
#endif
    for (int c = 0; c < nchunks; c++) {
        const int st = c % NST;
        mbar_wait(&bars[wib][st], (uint32_t)((c / NST) & 1));
        // every lane has left chunk c-1 (the __syncwarp below): its stage takes chunk c+NST-1
        if (lane == 0 && c + NST - 1 < nchunks) issue(c + NST - 1);
        const VT* cs = reinterpret_cast<const VT*>(cst + st * CH * STEP_B) + lane;
        const int n = min(CH, W1 - c * CH);
        const int cr = c_of_k0 - c;                  // checkpoint taken inside this chunk (if any)
        const bool take = cr >= 0 && (c > 0 || jstar > 0);
        if (n == CH) {
#pragma unroll
            for (int j = 0; j < CH; j++) {
                if (take && j == jstar) checkpoint(&ck[(size_t)cr * 32]);
                uint32_t Cr[NR];
                unpack<NR>(cs[j * 32], Cr);
                step(Cr);
            }
        } else {
            for (int j = 0; j < n; j++) {
                if (take && j == jstar) checkpoint(&ck[(size_t)cr * 32]);
                uint32_t Cr[NR];
                unpack<NR>(cs[j * 32], Cr);
                step(Cr);
            }
        }
        __syncwarp();                                // every lane is done reading this stage
    }
}

// ------------------------------------------------------------------------------------------
// right -> left (predecessor x+1) fused with the re-run of the left -> right steps and winner-takes-all.
// Per pixel one 8-byte record:
//   .x = minS | best << 16      (best = 0xffff when the uniqueness test rejects the pixel)
//   .y = S[best-1] | S[best+1] << 16
// which k_select turns into the disparity (disp2 vote, sub-pixel, LR check).
//
// Per chunk of CH = 8 pixels the warp stages C, S (the vertical directions' sum) and the left-to-right
// checkpoint of the chunk, then walks the chunk in BOTH directions at once: step i runs the left-to-right
// recurrence at pixel i and the right-to-left one at pixel 7-i -- two independent dependency chains in one
// instruction stream -- and adds both L onto S in place.  After that the chunk holds S_total and the warp
// re-maps its lanes to (pixel = lane / 4, quarter = lane % 4) so that all 8 pixels do their winner-takes-all at
// once: each lane scans its quarter of one pixel's disparities from shared memory and the four quarters
// meet through two shuffle-xor steps.  That replaces two warp-wide reductions per PIXEL by two
// 4-lane reductions per CHUNK.  S_total never returns to memory.
// ------------------------------------------------------------------------------------------
template <int NR, bool TAP_S, bool PAD>
__global__ void __launch_bounds__(HW_WARPS * 32, NR <= 2 ? 3 : 1)      // 3 blocks per SM by shared memory: up to 80 registers
k_path_rl_wta_tma(const uint16_t* __restrict__ Cv, uint16_t* __restrict__ Sv, const uint16_t* __restrict__ ckpt,
                  uint2* __restrict__ rec, int W1, int rows, uint32_t P1p, uint32_t P2p, int uniq, int Dreal)
{
    using VT = typename Vec<NR>::T;
    constexpr int D = 64 * NR;
    constexpr int NST = wta_stages(NR);
    constexpr int STEP_B = 128 * NR;
    constexpr int NU = 2 * NR;                   // uint4 (8 disparities each) per lane in the WTA phase
    static_assert(CH == 8, "the WTA lane mapping assumes 8 pixels per chunk");
    extern __shared__ __align__(128) unsigned char hsm[];
    __shared__ __align__(8) uint64_t bars[HW_WARPS][NST];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row = blockIdx.x * HW_WARPS + wib;
    if (row >= rows) return;
    unsigned char* cst = hsm + (size_t)wib * ((2 * CH + 1) * NST * STEP_B);   // [NST][CH][STEP_B] C, the same for S, [NST][STEP_B] checkpoints
    unsigned char* sst = cst + NST * CH * STEP_B;
    unsigned char* kst = sst + NST * CH * STEP_B;
    const unsigned char* Cg = reinterpret_cast<const unsigned char*>(Cv) + (size_t)row * W1 * STEP_B;
    const unsigned char* Sg = reinterpret_cast<const unsigned char*>(Sv) + (size_t)row * W1 * STEP_B;
    uint4* Stap = reinterpret_cast<uint4*>(Sv) + (size_t)row * W1 * (STEP_B / 16);
    rec += (size_t)row * W1;
    const int nchunks = (W1 + CH - 1) / CH;
    const unsigned char* Kg = reinterpret_cast<const unsigned char*>(ckpt) + (size_t)row * nchunks * STEP_B;

    // WTA-phase identity of this lane
    const int wp = lane >> 2, wq = lane & 3;
    // visit the NU uint4 in a lane-dependent rotation so that a quarter-warp's 128-bit loads hit 8 different
    // 16-byte bank groups (the natural order is a 4-way conflict for D = 128)
    const int rot = NR == 2 ? (((wq >> 1) + 2 * (wp & 1)) & 3) : NR == 4 ? ((wq + 4 * (wp & 1)) & 7) : (wp & 1);   // D = 256: the natural order is an 8-way conflict

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
        mbar_expect_tx(&bars[wib][st], 2 * bytes + (lo > 0 ? STEP_B : 0));
        bulk_g2s(cst + st * CH * STEP_B, Cg + (size_t)lo * STEP_B, bytes, &bars[wib][st]);
        bulk_g2s(sst + st * CH * STEP_B, Sg + (size_t)lo * STEP_B, bytes, &bars[wib][st]);
        if (lo > 0) bulk_g2s(kst + st * STEP_B, Kg + (size_t)c * STEP_B, STEP_B, &bars[wib][st]);   // the left-most chunk starts a path
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++)
            if (s < nchunks) issue(s);
    }
    uint32_t M[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) M[r] = 0;
#if V3D_STEP_CARRY
    uint32_t mmr = 0, mml = 0;                       // M / Ml hold L of the last pixel, mmr / mml its minimum over d
    // one step of a direction: state in, this pixel's L out
    auto step_r = [&](const uint32_t (&Cr)[NR], uint32_t (&L)[NR]) {
        path_step_carry<NR>(M, mmr, Cr, P1p, P2p, lane);
#pragma unroll
        for (int r = 0; r < NR; r++) L[r] = M[r];
    };
#else
    auto step_r = [&](const uint32_t (&Cr)[NR], uint32_t (&L)[NR]) { path_step<NR>(M, Cr, L, P1p, P2p, lane); };
#endif
    for (int c = 0; c < nchunks; c++) {
        const int st = c % NST;
        mbar_wait(&bars[wib][st], (uint32_t)((c / NST) & 1));
        const int hi = W1 - c * CH, lo = max(hi - CH, 0), n = hi - lo;
        const VT* cs = reinterpret_cast<const VT*>(cst + st * CH * STEP_B) + lane;
        VT* ss = reinterpret_cast<VT*>(sst + st * CH * STEP_B) + lane;
        // ---- phase 1: the path steps of both horizontal directions; S_total replaces S in the stage ----
        uint32_t Ml[NR];                             // left-to-right state entering the chunk
        if (lo > 0) unpack<NR>(reinterpret_cast<const VT*>(kst + st * STEP_B)[lane], Ml);
        else {
#pragma unroll
            for (int r = 0; r < NR; r++) Ml[r] = 0;
        }
#if V3D_STEP_CARRY
        mml = 0;                                     // a checkpoint is M: (L = M, min = 0)
        auto step_l = [&](const uint32_t (&Cr)[NR], uint32_t (&L)[NR]) {
            path_step_carry<NR>(Ml, mml, Cr, P1p, P2p, lane);
#pragma unroll
            for (int r = 0; r < NR; r++) L[r] = Ml[r];
        };
#else
        auto step_l = [&](const uint32_t (&Cr)[NR], uint32_t (&L)[NR]) { path_step<NR>(Ml, Cr, L, P1p, P2p, lane); };
#endif
        if (n == CH) {
            // Steps 0..3 keep their two L in registers; steps 4..7 meet the pixels the other direction has already
            // visited, so every pixel's S is read and written ONCE (S + L_left-to-right + L_right-to-left).
            uint32_t keepA[CH / 2][NR], keepB[CH / 2][NR];
#pragma unroll
            for (int i = 0; i < CH; i++) {
                const int j = CH - 1 - i;
                uint32_t Ca[NR], Cb[NR], Sa[NR], Sb[NR], La[NR], Lb[NR];
                unpack<NR>(cs[i * 32], Ca);
                unpack<NR>(cs[j * 32], Cb);
                step_l(Ca, La);                                  // left to right at pixel i
                step_r(Cb, Lb);                                  // right to left at pixel 7 - i
                if (i < CH / 2) {
#pragma unroll
                    for (int r = 0; r < NR; r++) { keepA[i][r] = La[r]; keepB[i][r] = Lb[r]; }
                } else {
                    unpack<NR>(ss[i * 32], Sa);                  // pixel i: right-to-left was here at step 7 - i = j
#pragma unroll
                    for (int r = 0; r < NR; r++) Sa[r] += La[r] + keepB[j][r];
                    ss[i * 32] = pack<NR>(Sa);
                    unpack<NR>(ss[j * 32], Sb);                  // pixel j: left-to-right was here at step j
#pragma unroll
                    for (int r = 0; r < NR; r++) Sb[r] += Lb[r] + keepA[j][r];
                    ss[j * 32] = pack<NR>(Sb);
                }
            }
        } else {
            for (int j = 0; j < n; j++) {
                uint32_t Cr[NR], Sr[NR], L[NR];
                unpack<NR>(cs[j * 32], Cr);
                unpack<NR>(ss[j * 32], Sr);
                step_l(Cr, L);
#pragma unroll
                for (int r = 0; r < NR; r++) Sr[r] += L[r];
                ss[j * 32] = pack<NR>(Sr);
            }
            for (int j = n - 1; j >= 0; j--) {
                uint32_t Cr[NR], Sr[NR], L[NR];
                unpack<NR>(cs[j * 32], Cr);
                unpack<NR>(ss[j * 32], Sr);
                step_r(Cr, L);
#pragma unroll
                for (int r = 0; r < NR; r++) Sr[r] += L[r];
                ss[j * 32] = pack<NR>(Sr);
            }
        }
        __syncwarp();
        // ---- phase 2: winner-takes-all of the chunk's pixels, 4 lanes per pixel ----
        {
            const uint4* px = reinterpret_cast<const uint4*>(sst + st * CH * STEP_B) + wp * (STEP_B / 16) + wq * NU;
            uint4 v[NU];
            uint32_t key = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < NU; k++) {
                const int u = (k + rot) & (NU - 1);
                v[k] = px[u];
                const uint32_t d0 = (uint32_t)(wq * NU + u) * 8;
                if (TAP_S && wp < n) Stap[(size_t)(lo + wp) * (STEP_B / 16) + wq * NU + u] = v[k];
                if (PAD && (int)d0 >= Dreal) v[k] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);   // padding
                const uint32_t dc0 = d0 | ((d0 + 1) << 8);
                const uint32_t w[4] = { v[k].x, v[k].y, v[k].z, v[k].w };
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t dc = dc0 + t * 0x0202u;
                    key = __vimin3_u32(key, __byte_perm(w[t], dc, 0x7104), __byte_perm(w[t], dc, 0x7325));
                }
            }
            key = min(key, __shfl_xor_sync(V3D_FULL_MASK, key, 1));
            key = min(key, __shfl_xor_sync(V3D_FULL_MASK, key, 2));
            const uint32_t best = key & 0xffu, minS = key >> 8;
            // uniqueness: the smallest S over |d - best| > 1.  One lane per pixel takes the sub-pixel neighbours and
            // then blanks S[best-1 .. best+1] in the staged chunk, so that the second minimum is a plain packed minimum
            // over the re-read row (1 instruction per two disparities instead of 5 for a masked one).
            uint16_t* s16 = reinterpret_cast<uint16_t*>(sst + st * CH * STEP_B + wp * STEP_B);
            uint32_t sm1 = 0, sp1 = 0;
            if (wq == 0) {
                sm1 = s16[best > 0 ? best - 1 : 0];
                sp1 = s16[best < D - 1 ? best + 1 : D - 1];
                s16[best] = 0xffffu;
                if (best > 0) s16[best - 1] = 0xffffu;
                if (best < D - 1) s16[best + 1] = 0xffffu;
            }
            __syncwarp();
            uint32_t m2 = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < NU; k++) {
                const int u = (k + rot) & (NU - 1);
                uint4 w4 = px[u];
                if (PAD && (int)((wq * NU + u) * 8) >= Dreal) w4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                m2 = __vimin3_u16x2(m2, w4.x, w4.y);
                m2 = __vimin3_u16x2(m2, w4.z, w4.w);
            }
            m2 = __vminu2(m2, __byte_perm(m2, 0, 0x1032));
            m2 = __vminu2(m2, __shfl_xor_sync(V3D_FULL_MASK, m2, 1));
            m2 = __vminu2(m2, __shfl_xor_sync(V3D_FULL_MASK, m2, 2));
            const uint32_t minS2 = m2 & 0xffffu;
            if (wq == 0 && wp < n) {
                const bool reject = minS2 * (uint32_t)(100 - uniq) < minS * 100u;
                rec[lo + wp] = make_uint2((minS & 0xffffu) | ((reject ? 0xffffu : best) << 16), sm1 | (sp1 << 16));
            }
        }
        __syncwarp();                            // every lane is done reading this stage
        if (lane == 0 && c + NST < nchunks) issue(c + NST);
    }
}

template <int NR>
int launch_lr(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    const int rows = batch * ctx->H;
    const uint32_t P1p = (uint32_t)ctx->P1 * 0x10001u, P2p = (uint32_t)ctx->P2 * 0x10001u;
    const size_t smem = (size_t)HW_WARPS * lrc_stages(NR) * CH * 128 * NR;
    dim3 grid((rows + HW_WARPS - 1) / HW_WARPS), block(HW_WARPS * 32);
    V3D_CUDA(cudaFuncSetAttribute(k_path_lr_ckpt<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    V3dScope scope(ctx, ST_LR, st);
    k_path_lr_ckpt<NR><<<grid, block, smem, st>>>(ctx->C, ctx->ckpt, ctx->W1, rows, P1p, P2p);
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

template <int NR>
int launch_wta(v3d_ctx* ctx, int batch, cudaStream_t st, bool tap_s)
{
    const int rows = batch * ctx->H;
    const uint32_t P1p = (uint32_t)ctx->P1 * 0x10001u, P2p = (uint32_t)ctx->P2 * 0x10001u;
    const size_t smem = (size_t)HW_WARPS * (2 * CH + 1) * wta_stages(NR) * 128 * NR;
    dim3 grid((rows + HW_WARPS - 1) / HW_WARPS), block(HW_WARPS * 32);
    const bool pad = ctx->D != ctx->Dk;
    auto wta = tap_s ? (pad ? k_path_rl_wta_tma<NR, true, true> : k_path_rl_wta_tma<NR, true, false>)
                     : (pad ? k_path_rl_wta_tma<NR, false, true> : k_path_rl_wta_tma<NR, false, false>);
    V3D_CUDA(cudaFuncSetAttribute(wta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    V3dScope scope(ctx, ST_WTA, st);
    wta<<<grid, block, smem, st>>>(ctx->C, ctx->S, ctx->ckpt, ctx->rec, ctx->W1, rows, P1p, P2p, ctx->uniq, ctx->D);
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

}  // namespace

// left-to-right checkpoint pass: reads C, writes the path state entering every 8-pixel chunk
int v3d_launch_path_lr(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    switch (ctx->Dk) {
        case 64: return launch_lr<1>(ctx, batch, st);
        case 128: return launch_lr<2>(ctx, batch, st);
        case 256: return launch_lr<4>(ctx, batch, st);
    }
    return v3d_fail(V3D_EINVAL, "numDisparities %d unsupported (64, 128, 256)", ctx->D);
}

// last path kernel: both horizontal directions (left-to-right re-run from the checkpoints) fused with winner-takes-all
int v3d_launch_path_rl_wta(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    const bool tap_s = ctx->debug_taps != 0;   // parity tests ask the WTA pass to also store S_total
    switch (ctx->Dk) {
        case 64: return launch_wta<1>(ctx, batch, st, tap_s);
        case 128: return launch_wta<2>(ctx, batch, st, tap_s);
        case 256: return launch_wta<4>(ctx, batch, st, tap_s);
    }
    return v3d_fail(V3D_EINVAL, "numDisparities %d unsupported (64, 128, 256)", ctx->D);
}
