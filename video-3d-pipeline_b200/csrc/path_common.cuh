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
#else
    m = __vminu2(m, __byte_perm(m, 0, 0x1032));          // both halves = this lane's minimum
    const uint32_t mm = __reduce_min_sync(V3D_FULL_MASK, m);
    const uint32_t neg = __vadd2(~mm, 0x00010001u);       // -min in both halves
#endif
#pragma unroll
    for (int k = 0; k < NR; k++) M[k] = __viaddmin_s16x2(L[k], neg, P2p);
}

enum { S_WRITE = 0, S_ACCUM = 1 };

}  // namespace
