// Semi-global path aggregation and winner-takes-all.
// Replaces the middle of cv2.StereoSGBM.compute (depth.py:341): OpenCV computeDisparitySGBM's
// L_r recurrences, S = sum_r L_r, per-pixel argmin, uniqueness test and the S[best-1], S[best+1]
// taps the sub-pixel step needs.  Spec: SURVEY.md Appendix A.3 / A.4.
//
// One WARP owns one path (a row for the horizontal directions, a wrapped column/diagonal for the
// vertical ones) and walks it sequentially.  Lane l holds the 2*NR consecutive disparities
// d = 2*NR*l .. 2*NR*l + 2*NR-1 as NR packed int16x2 registers, so one warp-step is one coalesced
// 128*NR-byte load of C, the DPX packed min/add recurrence and one CREDUX warp-min.  S is written by the
// top-down vertical sweep, accumulated by the bottom-up one (MODE_HH, L2 reductions) and consumed by the last
// kernel (both row directions + winner-takes-all, k_paths_h.cu); sums of integers, so the order does not matter.
//   state kept per path:  M[d] = min(L[d] - min_d L, P2)      (the P2 clamp folds the "minL + P2" term)
//   step:                 L[d] = C[d] + min(M[d], min(M[d-1], M[d+1]) + P1)
// A predecessor outside the window contributes L = 0, i.e. M = 0, which is also the initial state.
#include "path_common.cuh"
#include "tma.cuh"

#include <cooperative_groups.h>
#include <type_traits>
namespace cg = cooperative_groups;

namespace {



// ------------------------------------------------------------------------------------------
// Vertical / diagonal directions.  sx = x step along the path (-dx), sy = y step (-dy).
// Warp k of a frame walks the path that is at column k in the first row; columns wrap, and a wrap
// is a path start.
// ------------------------------------------------------------------------------------------
template <int NR, int SMODE, int PF>
__global__ void __launch_bounds__(256)
k_path_vert(const uint16_t* __restrict__ Cv, uint16_t* __restrict__ Sv, int W1, int H, int sx, int sy,
            uint32_t P1p, uint32_t P2p)
{
    using VT = typename Vec<NR>::T;
    constexpr int D = 64 * NR;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= W1) return;
    const size_t frame = (size_t)blockIdx.y * H * W1;
    const VT* C = reinterpret_cast<const VT*>(Cv) + frame * 32 + lane;
    VT* S = reinterpret_cast<VT*>(Sv) + frame * 32 + lane;

    int xp = k, yp = sy > 0 ? 0 : H - 1;                // prefetch cursor
    VT cq[PF], sq[PF];
#pragma unroll
    for (int j = 0; j < PF; j++) {
        if (j < H) {
            const size_t o = ((size_t)yp * W1 + xp) * 32;
            cq[j] = __ldg(C + o);
            if (SMODE == S_ACCUM) sq[j] = S[o];
            xp += sx; if (xp < 0) xp += W1; if (xp >= W1) xp -= W1;
            yp += sy;
        }
    }
    uint32_t M[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) M[r] = 0;
    int x = k, y = sy > 0 ? 0 : H - 1;
    for (int base = 0; base < H; base += PF) {
#pragma unroll
        for (int j = 0; j < PF; j++) {
            const int i = base + j;
            if (i >= H) break;
            uint32_t Cr[NR], Sr[NR], L[NR];
            unpack<NR>(cq[j], Cr);
            if (SMODE == S_ACCUM) unpack<NR>(sq[j], Sr);
            if (i + PF < H) {
                const size_t o = ((size_t)yp * W1 + xp) * 32;
                cq[j] = __ldg(C + o);
                if (SMODE == S_ACCUM) sq[j] = S[o];
                xp += sx; if (xp < 0) xp += W1; if (xp >= W1) xp -= W1;
                yp += sy;
            }
            path_step<NR>(M, Cr, L, P1p, P2p, lane);
#pragma unroll
            for (int r = 0; r < NR; r++) Sr[r] = (SMODE == S_ACCUM) ? Sr[r] + L[r] : L[r];
            S[((size_t)y * W1 + x) * 32] = pack<NR>(Sr);
            // advance; leaving the window on either side starts a new path at the far edge
            x += sx; y += sy;
            if (x < 0 || x >= W1) {
                x = x < 0 ? x + W1 : x - W1;
#pragma unroll
                for (int r = 0; r < NR; r++) M[r] = 0;
            }
        }
    }
    (void)D;
}

// ------------------------------------------------------------------------------------------
// The three directions of one vertical sweep fused: (x, y-sy), (x-1, y-sy), (x+1, y-sy).
// One thread-block CLUSTER of V3_CL CTAs owns a whole frame; warp g of the cluster owns the CPW
// consecutive columns [g*CPW, (g+1)*CPW) for every row and every direction, so C is read once and
// S is touched once per sweep (a plain store in WRITE mode, RED.ADD in ACCUM mode) instead of three reads of C
// and three read-modify-writes of S.
// Path state M lives in shared memory ([dir][column][d]); a diagonal path moves to the neighbouring
// column every row, so the only inter-warp traffic is one 2*NR*64-byte state vector per warp
// boundary and direction per row.  Each warp PUSHES its two boundary vectors into its neighbours' inboxes
// (st.async through distributed shared memory, completion counted on the receiver's mbarrier), so a warp
// synchronises with its two neighbours only: no cluster-wide barrier, no memory fence, no remote loads.
// ------------------------------------------------------------------------------------------
// V3_CL CTAs per cluster, V3_NW warps per CTA: 8 x 32 for D <= 128 (portable cluster size); 16 x 18 for
// D = 256, whose 512-byte state vectors need the shared memory of 16 SMs per frame (non-portable size).

// C loads of the D = 256 fused sweep (V3D_V3_D256_HINTS).  That sweep is latency-bound (18 warps per SM, about half of
// its issue slots used, 45-53 % of its stall samples on the long scoreboard: profiles/r02e_vert3_d256_stalls.txt):
// ptxas gives ALL global loads of the row body one scoreboard, so the row's first use of C waits for the youngest
// outstanding load -- the one issued less than a column earlier for the last column of the next row -- and the kernel's
// few spilled registers (the st.async target addresses, re-read every row) are evicted from L1 by the C stream.
//   * every C value is used once: the loads do not allocate in L1 (LDG.E.NA), the spill slots stay resident;
//   * the row after the next one is prefetched into L2 at the top of every row (CCTL.E.PF2), so that the wait is an L2
//     hit instead of a DRAM round trip.
// Measured on B200: the two sweeps of cfg5 7.34 -> 7.05 ms per 7 frames.  For D <= 128 (32 warps per SM, integer-pipe and
// issue bound) the same two changes are neutral / 3 % slower (3.47 -> 3.48 / 3.59 ms), so they stay off there.
#ifndef V3D_V3_D256_HINTS
#define V3D_V3_D256_HINTS 1
#endif
template <int NR> __device__ __forceinline__ typename Vec<NR>::T ld_c(const typename Vec<NR>::T* p) { return __ldg(p); }
#if V3D_V3_D256_HINTS
template <> __device__ __forceinline__ uint4 ld_c<4>(const uint4* p)
{
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NR, int CPW, int SMODE, int V3_CL, int V3_NW>
__global__ void __launch_bounds__(V3_NW * 32, 1)
k_path_vert3(const uint16_t* __restrict__ Cv, uint16_t* __restrict__ Sv, int W1, int H, int sy,
             uint32_t P1p, uint32_t P2p, int balanced)
{
    using VT = typename Vec<NR>::T;
    extern __shared__ uint4 v3smem[];
    constexpr int COLS = V3_NW * CPW;
    VT* Mst = reinterpret_cast<VT*>(v3smem);          // [3][COLS][32]   path state
    VT* inbox = Mst + 3 * COLS * 32;                  // [2][V3_NW][2][32] boundary vectors from the neighbours, double buffered
    uint64_t* mb = reinterpret_cast<uint64_t*>(inbox + 2 * V3_NW * 2 * 32);   // [2][V3_NW] one mbarrier per warp and buffer
    uint4* nbr = reinterpret_cast<uint4*>(mb + 2 * V3_NW);                    // [V3_NW] lane 0's st.async targets {left inbox, left barrier, right inbox, right barrier}
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int frame = blockIdx.x / V3_CL;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // Columns per warp: `balanced` spreads the W1 columns over ALL warps of the cluster (CPW or CPW - 1 each, so every
    // SM of the cluster carries the same share of the row); otherwise warps are filled with CPW columns in order and
    // the last ones stay empty (narrow frames, where a balanced share would drop below the 3 columns the
    // arrive / interior / wait schedule needs).
    constexpr int NWT = V3_CL * V3_NW;
    const int gw_ = rank * V3_NW + w;
    constexpr bool CAN_BALANCE = NR <= 2;           // (D = 256 never balances: see launch_vert3)
    const int gcol0 = (CAN_BALANCE && balanced) ? (int)((long long)gw_ * W1 / NWT) : gw_ * CPW;       // first column this warp owns
    const int lc0 = w * CPW;
    // Addresses.  WALK: the kernel parameters stay in the constant bank and ONE 64-bit element offset walks down (or up)
    // the rows -- the row top is two wide adds instead of two 64-bit multiplies and there are no row-invariant base
    // pointers to spill (D = 256: the last spills go, 7.06 -> 6.93 ms per 7 frames of cfg5; D = 64: 2.62 -> 2.59 ms).
    // At D = 128 ptxas turns the same source into a body that re-reads spilled values in the middle of a row
    // (3.38 -> 3.68 ms), so that instantiation keeps the frame pointers and multiplies the row index.
#ifndef V3D_V3_WALK2
#define V3D_V3_WALK2 0
#endif
    constexpr bool WALK = NR != 2 || V3D_V3_WALK2;
    const VT* const C = WALK ? reinterpret_cast<const VT*>(Cv) : reinterpret_cast<const VT*>(Cv) + (size_t)frame * H * W1 * 32 + lane;
    VT* const S = WALK ? reinterpret_cast<VT*>(Sv) : reinterpret_cast<VT*>(Sv) + (size_t)frame * H * W1 * 32 + lane;
    const ptrdiff_t rstride = (ptrdiff_t)sy * W1 * 32;                       // elements from a row to the next one of the sweep
    size_t off = (((size_t)frame * H + (sy > 0 ? 0 : H - 1)) * W1 + gcol0) * 32 + lane;   // WALK: this warp's first column in the current row

    uint32_t zr[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) zr[r] = 0;
    const VT zero = pack<NR>(zr);
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int j = 0; j < CPW; j++) Mst[((d * COLS) + lc0 + j) * 32 + lane] = zero;
    if (threadIdx.x < 2 * V3_NW) mbar_init(mb + threadIdx.x, 1);
    fence_mbar_init_cluster();
    cluster.sync();

    constexpr int PARSTRIDE = V3_NW * 2 * 32;         // inbox elements per buffer
    const int nv = (CAN_BALANCE && balanced) ? (int)((long long)(gw_ + 1) * W1 / NWT) - gcol0
                                             : min(max(W1 - gcol0, 0), CPW);      // columns of this warp inside the window
    const bool has_right = CAN_BALANCE ? gcol0 + nv < W1 : gcol0 + CPW < W1;   // a column to the right of this warp's last one exists
    V3D_DASSERT(nv >= 0 && nv <= CPW);
    VT cq[CPW];
    {
        const int y = sy > 0 ? 0 : H - 1;
        const VT* Crow = WALK ? C + off : C + ((size_t)y * W1 + gcol0) * 32;
#pragma unroll
        for (int j = 0; j < CPW; j++)
            if (j < nv) cq[j] = ld_c<NR>(Crow + j * 32);
    }
    VT* Md = Mst + (0 * COLS + lc0) * 32 + lane;      // this warp's state rows, one per direction
    VT* Ml = Mst + (1 * COLS + lc0) * 32 + lane;
    VT* Mr = Mst + (2 * COLS + lc0) * 32 + lane;
    // inbox slot 0 receives the left neighbour's (x-1)-direction state of its last column, slot 1 the right
    // neighbour's (x+1)-direction state of its first column
    const bool has_left = gcol0 > 0 && nv > 0;
    const VT* in_l = inbox + (w * 2 + 0) * 32 + lane;
    const VT* in_r = inbox + (w * 2 + 1) * 32 + lane;
    uint64_t* my_bar = mb + w;
    const uint32_t rx_bytes = ((has_left ? 1u : 0u) + (has_right ? 1u : 0u)) * 32u * (uint32_t)sizeof(VT);
    uint32_t to_l = 0, to_l_bar = 0, to_r = 0, to_r_bar = 0;      // cluster addresses in the neighbours' shared memory
    V3D_DASSERT(rank >= 0 && rank < V3_CL && (int)cluster.num_blocks() == V3_CL);
    V3D_DASSERT(nv == 0 || (gcol0 < W1 && gcol0 + nv <= W1));
    if (has_left) {
        const int nr = w > 0 ? rank : rank - 1, nw = w > 0 ? w - 1 : V3_NW - 1;
        V3D_DASSERT(nr >= 0 && nr < V3_CL && nw >= 0 && nw < V3_NW);      // the left neighbour's inbox slot exists
        to_l = mapa_u32(smem_u32(inbox + (nw * 2 + 1) * 32 + lane), nr);
        to_l_bar = mapa_u32(smem_u32(mb + nw), nr);
    }
    if (has_right) {
        const int nr = w < V3_NW - 1 ? rank : rank + 1, nw = w < V3_NW - 1 ? w + 1 : 0;
        V3D_DASSERT(nr >= 0 && nr < V3_CL && nw >= 0 && nw < V3_NW);      // a column to the right implies a CTA / warp to the right
        to_r = mapa_u32(smem_u32(inbox + (nw * 2 + 0) * 32 + lane), nr);
        to_r_bar = mapa_u32(smem_u32(mb + nw), nr);
    }
    // The four cluster addresses are needed once per row and would otherwise be the registers ptxas spills to local
    // memory (the kernel sits on its register cap): lane 0's values live in shared memory, one broadcast LDS.128 per row
    // brings them back (a lane's inbox slot is lane 0's plus lane * sizeof(VT): mapa only replaces the CTA bits).
    if (lane == 0) nbr[w] = make_uint4(to_l, to_l_bar, to_r, to_r_bar);
    __syncwarp();
    const uint32_t nbr_addr = smem_u32(nbr + w);
    const uint32_t lane_off = (uint32_t)lane * (uint32_t)sizeof(VT);
    // One column: the three path steps, the state write-back and the S store.  inl / inr are the previous
    // row's states arriving from the left / right neighbour column.
    auto do_col = [&](int j, const uint32_t (&inl)[NR], const uint32_t (&inr)[NR], bool more, const VT* Cnext,
                      const VT* Snext, VT* Srow) {
        uint32_t Cr[NR], Sr[NR], L[NR], M[NR];
        unpack<NR>(cq[j], Cr);
        // next row's C (on the last row Cnext points at the current row again: an unconditional load goes
        // straight into cq, a predicated one costs a dependent move that waits for it)
        cq[j] = ld_c<NR>(Cnext + j * 32);
        unpack<NR>(Md[j * 32], M);                               // (x, y-sy)
        path_step<NR>(M, Cr, L, P1p, P2p, lane);
        Md[j * 32] = pack<NR>(M);
#pragma unroll
        for (int r = 0; r < NR; r++) { Sr[r] = L[r]; M[r] = inl[r]; }
        path_step<NR>(M, Cr, L, P1p, P2p, lane);                 // (x-1, y-sy)
        Ml[j * 32] = pack<NR>(M);
#pragma unroll
        for (int r = 0; r < NR; r++) { Sr[r] += L[r]; M[r] = inr[r]; }
        path_step<NR>(M, Cr, L, P1p, P2p, lane);                 // (x+1, y-sy)
        Mr[j * 32] = pack<NR>(M);
#pragma unroll
        for (int r = 0; r < NR; r++) Sr[r] += L[r];
        if (SMODE == S_ACCUM) {
            // S += (the three L): reduction performed by the L2 (RED, no return value), so the sweep neither
            // loads S nor waits for it.  Packed 16-bit sums never carry (S_total < 2^16 by construction), so
            // wide integer adds are exact.
            if (NR == 1) {
                atomicAdd(reinterpret_cast<unsigned int*>(Srow + j * 32), Sr[0]);
            } else {
                unsigned long long* sp = reinterpret_cast<unsigned long long*>(Srow + j * 32);
#pragma unroll
                for (int r = 0; r < NR; r += 2)
                    atomicAdd(sp + r / 2, (unsigned long long)Sr[r] | ((unsigned long long)Sr[r + (NR > 1 ? 1 : 0)] << 32));
            }
        } else {
            Srow[j * 32] = pack<NR>(Sr);
        }
    };
    const uint32_t zeros[NR] = {};
    for (int i = 0; i < H; i++) {
        const int y = sy > 0 ? i : H - 1 - i;
        const int yn = sy > 0 ? i + 1 : H - 2 - i;
        const int par = i & 1;
        const bool more = i + 1 < H;
        const size_t off_next = more ? off + rstride : off;
        const VT* Cnext = WALK ? C + off_next : C + ((size_t)(more ? yn : y) * W1 + gcol0) * 32;
        VT* Srow = WALK ? S + off : S + ((size_t)y * W1 + gcol0) * 32;
        const VT* Snext = WALK ? S + off_next : S + ((size_t)(more ? yn : y) * W1 + gcol0) * 32;
#if V3D_V3_D256_HINTS
        if constexpr (NR == 4) {
            // row i + 2 (clamped to the last row: a redundant prefetch is cheaper than a branch)
            const VT* C2 = C + (i + 2 < H ? off + 2 * rstride : off_next);
#pragma unroll
            for (int j = 0; j < CPW; j++)
                if (j < nv) prefetch_l2(C2 + j * 32);
        }
#endif
        const uint32_t phase = (i >> 1) & 1;
        V3D_DASSERT(par == 0 || par == 1);                               // double-buffered inboxes: [par][warp][side]
        V3D_DASSERT((char*)(in_r + par * PARSTRIDE) + sizeof(VT) <= (char*)mb && y >= 0 && y < H);
        // (a warp with a right neighbour is full unless the columns are balanced; a compile-time index where possible)
        uint4 nb;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(nb.x), "=r"(nb.y), "=r"(nb.z), "=r"(nb.w) : "r"(nbr_addr));
        const uint32_t par_off = lane_off + par * PARSTRIDE * (uint32_t)sizeof(VT);
        V3D_DASSERT(!has_left || (nb.x + lane_off == to_l && nb.y == to_l_bar));      // the table holds this lane's own mapa results
        V3D_DASSERT(!has_right || (nb.z + lane_off == to_r && nb.w == to_r_bar));
        if (has_right) st_async(nb.z + par_off, Ml[((NR <= 2 && CPW >= 4 ? nv : CPW) - 1) * 32], nb.w + par * V3_NW * 8);
        if (has_left) st_async(nb.x + par_off, Mr[0], nb.y + par * V3_NW * 8);
        if (lane == 0 && rx_bytes) mbar_expect_tx(my_bar + par * V3_NW, rx_bytes);
        // A warp with NV >= 3 columns: the interior columns need nothing from other warps, so they run between the
        // barrier's arrive and wait; only the two edge columns wait for the neighbours' states.
        auto fast_row = [&](auto nvc) {
            constexpr int NV = decltype(nvc)::value;
            uint32_t carry[NR], save_r1[NR], inr[NR];
            unpack<NR>(Ml[0], carry);                            // old state of column 0 -> column 1
            unpack<NR>(Mr[1 * 32], save_r1);                     // old state of column 1 -> column 0 (used last)
#pragma unroll
            for (int j = 1; j < NV - 1; j++) {
                uint32_t nextcarry[NR];
                unpack<NR>(Ml[j * 32], nextcarry);
                unpack<NR>(Mr[(j + 1) * 32], inr);
                do_col(j, carry, inr, more, Cnext, Snext, Srow);
#pragma unroll
                for (int r = 0; r < NR; r++) carry[r] = nextcarry[r];
            }
            if (rx_bytes) mbar_wait(my_bar + par * V3_NW, phase);
            uint32_t edge[NR];
            if (gcol0 > 0) unpack<NR>(in_l[par * PARSTRIDE], edge);
            else {
#pragma unroll
                for (int r = 0; r < NR; r++) edge[r] = 0;
            }
            do_col(0, edge, save_r1, more, Cnext, Snext, Srow);
            if (has_right) unpack<NR>(in_r[par * PARSTRIDE], edge);
            else {
#pragma unroll
                for (int r = 0; r < NR; r++) edge[r] = 0;
            }
            do_col(NV - 1, carry, edge, more, Cnext, Snext, Srow);
        };
        // (the second instantiation of the row body only where the balanced distribution uses it: at D = 256 it costs
        // registers -- spills -- and instruction cache for nothing)
        constexpr bool TWO_WIDTHS = NR <= 2 && CPW >= 4;
        if (nv == CPW && CPW >= 3) {
            if constexpr (TWO_WIDTHS) {
                fast_row(std::integral_constant<int, CPW>{});
            } else {
                // written out (not through the generic lambda): the D = 256 sweep is register-starved (96 registers, 18
                // warps) and loses 15 % when the body goes through the lambda
                uint32_t carry[NR], save_r1[NR], inr[NR];
                unpack<NR>(Ml[0], carry);
                unpack<NR>(Mr[1 * 32], save_r1);
#pragma unroll
                for (int j = 1; j < CPW - 1; j++) {
                    uint32_t nextcarry[NR];
                    unpack<NR>(Ml[j * 32], nextcarry);
                    unpack<NR>(Mr[(j + 1) * 32], inr);
                    do_col(j, carry, inr, more, Cnext, Snext, Srow);
#pragma unroll
                    for (int r = 0; r < NR; r++) carry[r] = nextcarry[r];
                }
                if (rx_bytes) mbar_wait(my_bar + par * V3_NW, phase);
                uint32_t edge[NR];
                if (gcol0 > 0) unpack<NR>(in_l[par * PARSTRIDE], edge);
                else {
#pragma unroll
                    for (int r = 0; r < NR; r++) edge[r] = 0;
                }
                do_col(0, edge, save_r1, more, Cnext, Snext, Srow);
                if (has_right) unpack<NR>(in_r[par * PARSTRIDE], edge);
                else {
#pragma unroll
                    for (int r = 0; r < NR; r++) edge[r] = 0;
                }
                do_col(CPW - 1, carry, edge, more, Cnext, Snext, Srow);
            }
        } else if (TWO_WIDTHS && nv == CPW - 1) {
            if constexpr (TWO_WIDTHS) fast_row(std::integral_constant<int, CPW - 1>{});
        } else {
            if (rx_bytes) mbar_wait(my_bar + par * V3_NW, phase);
            uint32_t carry[NR], inr[NR];
            if (has_left) unpack<NR>(in_l[par * PARSTRIDE], carry);
            else {
#pragma unroll
                for (int r = 0; r < NR; r++) carry[r] = 0;
            }
#pragma unroll
            for (int j = 0; j < CPW; j++) {
                if (j < nv) {
                    uint32_t nextcarry[NR];
                    unpack<NR>(Ml[j * 32], nextcarry);
                    if (j + 1 < nv) {
                        if (j + 1 < CPW) unpack<NR>(Mr[(j + 1) * 32], inr);
                    } else if (has_right) {
                        unpack<NR>(in_r[par * PARSTRIDE], inr);
                    } else {
#pragma unroll
                        for (int r = 0; r < NR; r++) inr[r] = 0;
                    }
                    do_col(j, carry, inr, more, Cnext, Snext, Srow);
#pragma unroll
                    for (int r = 0; r < NR; r++) carry[r] = nextcarry[r];
                }
            }
        }
        if constexpr (WALK) off = off_next;
    }
    (void)zeros;
    cluster.sync();   // nobody may exit while a neighbour can still read its shared memory
}

// Launch (or, with query != nullptr, only ask how many clusters are co-resident) one fused sweep configuration.
template <int NR, int CPW, int V3_CL, int V3_NW>
int launch_vert3(v3d_ctx* ctx, int batch, int sy, bool accum, cudaStream_t st, int* query)
{
    using VT = typename Vec<NR>::T;
    const size_t smem = ((size_t)3 * V3_NW * CPW * 32 + (size_t)2 * V3_NW * 2 * 32) * sizeof(VT) + (size_t)2 * V3_NW * 8 + (size_t)V3_NW * 16;
    auto kw = k_path_vert3<NR, CPW, S_WRITE, V3_CL, V3_NW>;
    auto ka = k_path_vert3<NR, CPW, S_ACCUM, V3_CL, V3_NW>;
    V3D_CUDA(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    V3D_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (V3_CL > 8) {
        V3D_CUDA(cudaFuncSetAttribute(kw, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        V3D_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(V3_CL * batch);
    cfg.blockDim = dim3(V3_NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = V3_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (query) {
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kw, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        *query = n;
        return V3D_OK;
    }
    const uint16_t* C = ctx->C;
    uint16_t* S = ctx->S;
    const uint32_t P1p = (uint32_t)ctx->P1 * 0x10001u, P2p = (uint32_t)ctx->P2 * 0x10001u;
    // every warp of the cluster gets CPW or CPW - 1 columns when that leaves each at least 3 (see the kernel)
    // (not for D = 256: two instantiations of the row body per kernel thrash the instruction cache there -- measured 8.05
    // vs 7.25 ms for the two sweeps of cfg5)
    const int balanced = NR <= 2 && ctx->W1 / (V3_CL * V3_NW) >= 3 ? 1 : 0;
    V3D_CUDA(cudaLaunchKernelEx(&cfg, accum ? ka : kw, C, S, ctx->W1, ctx->H, sy, P1p, P2p, balanced));
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

// Columns per warp the frame needs with CL x NW warps per cluster -> the instantiation that holds them.
// Returns 1 when the configuration ran (or was queried), 0 when the frame is wider than the cluster's shared memory.
template <int NR, int CL, int NW>
int vert3_dispatch(v3d_ctx* ctx, int batch, int sy, bool accum, cudaStream_t st, int* query, int& rc)
{
    const int need = (ctx->W1 + CL * NW - 1) / (CL * NW);
    if (need <= 2) rc = launch_vert3<NR, 2, CL, NW>(ctx, batch, sy, accum, st, query);
    else if (need <= 4) rc = launch_vert3<NR, 4, CL, NW>(ctx, batch, sy, accum, st, query);
    else if (need <= 5) rc = launch_vert3<NR, 5, CL, NW>(ctx, batch, sy, accum, st, query);
    else if (need <= 6) rc = launch_vert3<NR, 6, CL, NW>(ctx, batch, sy, accum, st, query);
    else if (need <= 7) rc = launch_vert3<NR, 7, CL, NW>(ctx, batch, sy, accum, st, query);
    else if (need <= 8 && NR <= 2) rc = launch_vert3<NR, (NR <= 2 ? 8 : 7), CL, NW>(ctx, batch, sy, accum, st, query);
    else return 0;      // wider than one cluster's shared memory: per-direction kernels
    return 1;
}

// Returns 1 when the fused sweep ran, 0 when the shape does not fit it (caller falls back), < 0 on error.
// Cluster size: 16 CTAs for D = 256 (512-byte state vectors); for D <= 128 whichever of 9 and 8 CTAs keeps more SMs
// busy on THIS device -- clusters live inside a GPC, and on B200 15 clusters are co-resident at either size, so 9 CTAs
// per frame use 135 SMs where 8 use 120 (tools/probe_clusters.cu lists every size).
template <int NR>
int try_vert3(v3d_ctx* ctx, int batch, int sy, bool accum, cudaStream_t st)
{
    if (ctx->no_fused_vertical) return 0;
#ifndef V3D_V3_NW4
#define V3D_V3_NW4 18
#endif
#ifndef V3D_V3_NW2
#define V3D_V3_NW2 32
#endif
#ifndef V3D_V3_CL2
#define V3D_V3_CL2 9
#endif
    constexpr int NW = NR <= 2 ? V3D_V3_NW2 : V3D_V3_NW4;
    constexpr int CLA = NR <= 2 ? V3D_V3_CL2 : 16, CLB = NR <= 2 ? 8 : 16;     // preferred / portable cluster size
    constexpr int CLS = NR <= 2 ? 4 : 16;                                       // small cluster for narrow frames
    int rc = V3D_OK;
    if (!ctx->v3_cl) {
        // Narrow frames (W1 <= 4 * NW * 8 columns: e.g. 960-wide eyes) fit the shared memory of 4 CTAs: 33 such
        // clusters are co-resident on B200 (132 SMs) and every warp keeps 7 columns instead of 3-4, i.e. enough
        // interior columns to run while the neighbours' edge states are in flight.
        int ns = 0;
        if (CLS != CLA && ctx->W1 <= CLS * NW * 8 && ctx->W1 / (CLS * NW) >= 3) {
            const int fs = vert3_dispatch<NR, CLS, NW>(ctx, batch, sy, accum, st, &ns, rc);
            if (rc) return rc;
            if (fs && ns > 0) { ctx->v3_cl = CLS; ctx->max_clusters = ns; }
        }
    }
    if (!ctx->v3_cl) {
        int na = 0, nb = 0;
        const int fa = vert3_dispatch<NR, CLA, NW>(ctx, batch, sy, accum, st, &na, rc);
        if (rc) return rc;
        const int fb = CLB != CLA ? vert3_dispatch<NR, CLB, NW>(ctx, batch, sy, accum, st, &nb, rc) : 0;
        if (rc) return rc;
        if (fa && na > 0 && (!fb || na * CLA >= nb * CLB)) { ctx->v3_cl = CLA; ctx->max_clusters = na; }
        else if (fb && nb > 0) { ctx->v3_cl = CLB; ctx->max_clusters = nb; }
        else {                             // too wide, or this device cannot co-schedule such a cluster:
            ctx->max_clusters = 0;         // use the one-direction-per-launch kernels from now on
            ctx->no_fused_vertical = 1;
            return 0;
        }
    }
    if (CLS != CLA && ctx->v3_cl == CLS) {
        const int fit = vert3_dispatch<NR, CLS, NW>(ctx, batch, sy, accum, st, nullptr, rc);
        return rc ? rc : fit;
    }
    const int fit = ctx->v3_cl == CLA ? vert3_dispatch<NR, CLA, NW>(ctx, batch, sy, accum, st, nullptr, rc)
                                      : vert3_dispatch<NR, CLB, NW>(ctx, batch, sy, accum, st, nullptr, rc);
    if (rc) return rc;
    return fit;
}

template <int NR>
int launch_paths_nr(v3d_ctx* ctx, int batch, cudaStream_t st, bool tap_s)
{
    constexpr int PF = NR == 4 ? 4 : 8;
    const int W1 = ctx->W1, H = ctx->H;
    const uint32_t P1p = (uint32_t)ctx->P1 * 0x10001u, P2p = (uint32_t)ctx->P2 * 0x10001u;
    const uint16_t* C = ctx->C;
    uint16_t* S = ctx->S;
    const int wpb = 8;
    dim3 block(wpb * 32);
    dim3 gv((W1 + wpb - 1) / wpb, batch);
    // Order of the chain (integer sums, so any order gives the same S): the top-down vertical sweep WRITES S (C in,
    // S out: two volume passes, no read-modify-write), a bottom-up sweep (MODE_HH) accumulates onto it, the
    // left-to-right checkpoint pass reads C once more, and the last kernel reads C and S, re-runs left-to-right
    // from the checkpoints, runs right-to-left and feeds the winner-takes-all: 6.25 volume passes per frame
    // with the cost kernel's write, where "every direction reads C and read-modify-writes S" would need 16.
    // The checkpoint pass depends on C only, so it runs on a side stream next to the vertical sweep (which is
    // ALU-bound and leaves the SMs its clusters cannot occupy, and most of the HBM bandwidth, unused).
    const bool side = ctx->side_stream != nullptr && !ctx->timing;
    if (side) {
        V3D_CUDA(cudaEventRecord(ctx->ev_fork, st));
        V3D_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
        int rc = v3d_launch_path_lr(ctx, batch, ctx->side_stream);
        if (rc) return rc;
        V3D_CUDA(cudaEventRecord(ctx->ev_join, ctx->side_stream));
    }
    {
        V3dScope scope(ctx, ST_PATHS, st);
        // top-down sweep: predecessors (x, y-1), (x-1, y-1), (x+1, y-1)
        int fused = try_vert3<NR>(ctx, batch, +1, false, st);
        if (fused < 0) return fused;
        if (!fused) {
            k_path_vert<NR, S_WRITE, PF><<<gv, block, 0, st>>>(C, S, W1, H, 0, +1, P1p, P2p);
            k_path_vert<NR, S_ACCUM, PF><<<gv, block, 0, st>>>(C, S, W1, H, +1, +1, P1p, P2p);
            k_path_vert<NR, S_ACCUM, PF><<<gv, block, 0, st>>>(C, S, W1, H, -1, +1, P1p, P2p);
            V3D_LAUNCHED(ctx, 3);
        }
        if (ctx->p.mode == V3D_MODE_HH) {
            // bottom-up sweep: predecessors (x, y+1), (x+1, y+1), (x-1, y+1)
            fused = try_vert3<NR>(ctx, batch, -1, true, st);
            if (fused < 0) return fused;
            if (!fused) {
                k_path_vert<NR, S_ACCUM, PF><<<gv, block, 0, st>>>(C, S, W1, H, 0, -1, P1p, P2p);
                k_path_vert<NR, S_ACCUM, PF><<<gv, block, 0, st>>>(C, S, W1, H, -1, -1, P1p, P2p);
                k_path_vert<NR, S_ACCUM, PF><<<gv, block, 0, st>>>(C, S, W1, H, +1, -1, P1p, P2p);
                V3D_LAUNCHED(ctx, 3);
            }
        }
    }
    if (side) V3D_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    else if (int rc = v3d_launch_path_lr(ctx, batch, st)) return rc;
    return v3d_launch_path_rl_wta(ctx, batch, st);
}

}  // namespace

int v3d_launch_paths(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    const bool tap_s = ctx->debug_taps != 0;   // parity tests ask the WTA pass to also store S_total
    switch (ctx->Dk) {
        case 64: return launch_paths_nr<1>(ctx, batch, st, tap_s);
        case 128: return launch_paths_nr<2>(ctx, batch, st, tap_s);
        case 256: return launch_paths_nr<4>(ctx, batch, st, tap_s);
    }
    return v3d_fail(V3D_EINVAL, "numDisparities %d unsupported (64, 128, 256)", ctx->D);
}
