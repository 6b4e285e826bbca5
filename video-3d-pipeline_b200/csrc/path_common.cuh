// Device helpers shared by the path-aggregation kernels (k_paths.cu, k_paths_h.cu).
#pragma once
#include "v3d_internal.h"

namespace {

template <int NR> struct Vec;
template <> struct Vec<1> { using T = uint32_t; };
template <> struct Vec<2> { using T = uint2; };
template <> struct Vec<4> { using T = uint4; };

template <int NR> __device__ __forceinline__ void unpack(const typename Vec<NR>::T& v, uint32_t (&r)[NR]);
template <> __device__ __forceinline__ void unpack<1>(const uint32_t& v, uint32_t (&r)[1]) { r[0] = v; }
template <> __device__ __forceinline__ void unpack<2>(const uint2& v, uint32_t (&r)[2]) { r[0] = v.x; r[1] = v.y; }
template <> __device__ __forceinline__ void unpack<4>(const uint4& v, uint32_t (&r)[4]) { r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
template <int NR> __device__ __forceinline__ typename Vec<NR>::T pack(const uint32_t (&r)[NR]);
template <> __device__ __forceinline__ uint32_t pack<1>(const uint32_t (&r)[1]) { return r[0]; }
template <> __device__ __forceinline__ uint2 pack<2>(const uint32_t (&r)[2]) { return make_uint2(r[0], r[1]); }
template <> __device__ __forceinline__ uint4 pack<4>(const uint32_t (&r)[4]) { return make_uint4(r[0], r[1], r[2], r[3]); }

// One step of the recurrence.  M in/out, C in, L out.  P1p / P2p are P1, P2 duplicated in both halves.
// V3D_STEP_MIN3 = 1: the neighbours are shifted copies of M + P1 and the three-way minimum is one VIMNMX3.U16x2
// (4 instructions on the half-rate integer pipe per register instead of 5; the extra add can issue on the other pipe).
#ifndef V3D_STEP_MIN3
#define V3D_STEP_MIN3 1
#endif
#ifndef V3D_STEP_SUB
#define V3D_STEP_SUB 1       // M = min(L - min, P2) as a 32-bit subtraction + VIMNMX (only for D <= 128: one more instruction per register beyond that)
#endif
#ifndef V3D_STEP_UNEG
#define V3D_STEP_UNEG 0      // measured slower: the path kernels are issue-bound, not only integer-pipe-bound
#endif
template <int NR>
__device__ __forceinline__ void path_step(uint32_t (&M)[NR], const uint32_t (&C)[NR], uint32_t (&L)[NR],
                                          uint32_t P1p, uint32_t P2p, int lane)
{
#if V3D_STEP_MIN3
    uint32_t Mp[NR];                                   // M + P1; halves never carry: P2 + P1 < 2^15
#pragma unroll
    for (int k = 0; k < NR; k++) Mp[k] = M[k] + P1p;
#else
    const uint32_t (&Mp)[NR] = M;
#endif
    const uint32_t up = __shfl_up_sync(V3D_FULL_MASK, Mp[NR - 1], 1);
    const uint32_t dn = __shfl_down_sync(V3D_FULL_MASK, Mp[0], 1);
    // d = -1 and d = D do not exist.  Substituting the cell's own value for the missing neighbour is
    // exact (M[d] + P1 never beats M[d]), and costs nothing: only the byte-permute selector differs.
    const uint32_t sel_up = lane == 0 ? 0x5454u : 0x5432u;
    const uint32_t sel_dn = lane == 31 ? 0x3232u : 0x5432u;
    uint32_t sh[NR + 1];
    sh[0] = __byte_perm(up, Mp[0], sel_up);           // (M[d-1] for the even d, M[d-1] for the odd d)
#pragma unroll
    for (int k = 1; k < NR; k++) sh[k] = __byte_perm(Mp[k - 1], Mp[k], 0x5432);
    sh[NR] = __byte_perm(Mp[NR - 1], dn, sel_dn);
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < NR; k++) {
#if V3D_STEP_MIN3
        L[k] = C[k] + __vimin3_u16x2(sh[k], sh[k + 1], M[k]);
#else
        const uint32_t nb = __vminu2(sh[k], sh[k + 1]);
        L[k] = C[k] + __viaddmin_u16x2(nb, P1p, M[k]);   // halves never carry: C + P2 < 2^15
#endif
        m = __vminu2(m, L[k]);
    }
#if V3D_STEP_UNEG
    // The warp minimum and its negation stay off the half-rate integer pipe: the lane's smaller half is moved into
    // the HIGH half with a multiply (FMA pipe), CREDUX leaves (min << 16) in a uniform register, and -min in both
    // halves comes from a negate, a multiply-high and an add, all IMAD forms (3 integer-pipe instructions fewer per
    // step than PRMT + VIMNMX ... LOP3 + VIADD.16x2).
    m = __vminu2(m, m * 0x10000u);                           // high half = min(hi, lo); low half = 0
    const uint32_t r = __reduce_min_sync(V3D_FULL_MASK, m);  // = min_d L << 16
    const uint32_t x = 0u - r;                               // (-min & 0xffff) << 16
    const uint32_t neg = x + __umulhi(x, 0x10001u);          // + (x >> 16) as a multiply-high: -min in both halves
#pragma unroll
    for (int k = 0; k < NR; k++) M[k] = __viaddmin_s16x2(L[k], neg, P2p);
#else
    m = __vminu2(m, __byte_perm(m, 0, 0x1032));          // both halves = this lane's minimum
    const uint32_t mm = __reduce_min_sync(V3D_FULL_MASK, m);
    // V3D_STEP_SUB: every half of L is >= the minimum, so a plain 32-bit subtraction never borrows between the halves: it
    // can issue on either pipe (IADD3 / IMAD.IADD), and the clamp is a plain VIMNMX -- two instructions per register
    // instead of LOP3 + VIADD.16x2 per step and one VIADDMNMX per register: the same instruction count for 128
    // disparities with one or two of them off the half-rate integer pipe (1: D <= 128 only, 2: every D).
    constexpr bool sub = V3D_STEP_SUB == 2 || (V3D_STEP_SUB == 1 && NR <= 2);
    if constexpr (sub) {
#pragma unroll
        for (int k = 0; k < NR; k++) M[k] = __vminu2(L[k] - mm, P2p);
    } else {
        const uint32_t neg = __vadd2(~mm, 0x00010001u);   // -min in both halves
#pragma unroll
        for (int k = 0; k < NR; k++) M[k] = __viaddmin_s16x2(L[k], neg, P2p);
    }
#endif
}

// The same recurrence with the state carried as (L, min_d L) instead of M = min(L - min_d L, P2):
//   L'[d] = C[d] + min(L[d], L[d-1] + P1, L[d+1] + P1, P2 + min) - min
// (the clamp at P2 moves inside the minimum; the P2 + P1 terms of the clamped neighbours are dominated by P2).  Exact for
// the same reasons as above: all halves stay below 2^15 and C + x - min never borrows because x >= min.  For kernels that
// keep the path state in registers (the row kernels); a state that enters as M is (L = M, min = 0).  19 instead of 21
// instructions per 128 disparities and the warp reduction no longer sits between a step's result and the next step's
// shuffles -- but MEASURED SLOWER on B200 (last path kernel 2.86 vs 2.77 ms, checkpoint pass 1.34 vs 1.30 ms per 15
// frames of cfg2; bit-exact): ptxas turns C + x - min into IADD3, so the half-rate integer pipe keeps its 12
// instructions per step while two independent chains per warp already hid the reduction's latency.  Off by default.
#ifndef V3D_STEP_CARRY
#define V3D_STEP_CARRY 0
#endif
template <int NR>
__device__ __forceinline__ void path_step_carry(uint32_t (&Lp)[NR], uint32_t& mm, const uint32_t (&C)[NR], uint32_t P1p,
                                                uint32_t P2p, int lane)
{
    uint32_t Mp[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) Mp[k] = Lp[k] + P1p;
    const uint32_t up = __shfl_up_sync(V3D_FULL_MASK, Mp[NR - 1], 1);
    const uint32_t dn = __shfl_down_sync(V3D_FULL_MASK, Mp[0], 1);
    const uint32_t sel_up = lane == 0 ? 0x5454u : 0x5432u;
    const uint32_t sel_dn = lane == 31 ? 0x3232u : 0x5432u;
    uint32_t sh[NR + 1];
    sh[0] = __byte_perm(up, Mp[0], sel_up);
#pragma unroll
    for (int k = 1; k < NR; k++) sh[k] = __byte_perm(Mp[k - 1], Mp[k], 0x5432);
    sh[NR] = __byte_perm(Mp[NR - 1], dn, sel_dn);
    const uint32_t T = P2p + mm;
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        const uint32_t x = __vminu2(__vimin3_u16x2(sh[k], sh[k + 1], Lp[k]), T);
        Lp[k] = C[k] + x - mm;
        m = __vminu2(m, Lp[k]);
    }
    m = __vminu2(m, __byte_perm(m, 0, 0x1032));
    mm = __reduce_min_sync(V3D_FULL_MASK, m);            // min_d L in both halves
}
// (L, min) -> M, the form the checkpoints and the vertical sweep's shared-memory state use
template <int NR>
__device__ __forceinline__ void carry_to_state(const uint32_t (&Lp)[NR], uint32_t mm, uint32_t P2p, uint32_t (&M)[NR])
{
    const uint32_t neg = __vadd2(~mm, 0x00010001u);
#pragma unroll
    for (int k = 0; k < NR; k++) M[k] = __viaddmin_s16x2(Lp[k], neg, P2p);
}

enum { S_WRITE = 0, S_ACCUM = 1 };

}  // namespace
