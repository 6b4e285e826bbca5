// Disparity selection (disp2 vote, sub-pixel, LR check), 3x3 median, speckle filter and the
// reference's float / uint16 epilogue.
// Replaces the tail of cv2.StereoSGBM.compute (depth.py:341: OpenCV computeDisparitySGBM's per-row
// epilogue, medianBlur(disp, 3), filterSpeckles) and depth.py:341 (/16), :374 (<=0 -> 0), :400-403
// (per-frame min-max -> uint16).  Spec: SURVEY.md Appendix A.4 - A.6.
#include "v3d_internal.h"

namespace {

// The invalid value is (minDisparity - 1) * 16 (-16 for the reference's minDisparity = 0): a kernel argument `inv`.

// ---------------------------------------------------------------------------------------------
// One block per image row.  cv2 walks x descending and keeps, for every right-image column x2, the
// cheapest vote (strict '>' => among equal costs the largest x wins); atomicMin on
// (minS << 16 | 0xffff - x) reproduces that independent of thread order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_select(const uint2* __restrict__ rec, int16_t* __restrict__ raw, int W, int W1, int D, int maxdiff, int x0, int minD, int inv)
{
    // window column x <-> image column x + x0; disparity index d <-> minD + d pixels; x0 = max(minD + D, 0)
    extern __shared__ uint32_t smem[];
    uint32_t* key = smem;                                   // [W]
    int16_t* d16row = reinterpret_cast<int16_t*>(smem + W); // [W]
    const size_t row = blockIdx.x;
    rec += row * W1;
    raw += row * W;
    for (int X = threadIdx.x; X < W; X += blockDim.x) { key[X] = 0xffffffffu; d16row[X] = (int16_t)inv; }
    __syncthreads();
    for (int x = threadIdx.x; x < W1; x += blockDim.x) {
        const uint2 r = __ldg(rec + x);
        const uint32_t best = r.x >> 16;
        if (best == 0xffffu) continue;                      // failed the uniqueness test
        const int minS = (int)(r.x & 0xffffu);
        atomicMin(&key[x + x0 - minD - (int)best], ((uint32_t)minS << 16) | (0xffffu - (uint32_t)x));
        int d16 = ((int)best + minD) * 16;
        if (best > 0 && (int)best < D - 1) {
            const int sm1 = (int)(r.y & 0xffffu), sp1 = (int)(r.y >> 16);
            const int den = max(sm1 + sp1 - 2 * minS, 1);
            d16 += ((sm1 - sp1) * 16 + den) / (den * 2);    // C truncating division
        }
        d16row[x + x0] = (int16_t)d16;
    }
    __syncthreads();
    for (int X = threadIdx.x; X < W; X += blockDim.x) {
        int d1 = d16row[X];
        if (d1 != inv) {                                     // (columns outside the window kept the initial value)
            const int dl = d1 >> 4, dh = (d1 + 15) >> 4;
            const int xl = X - dl, xh = X - dh;
            bool bad = true;
            {
                bool t = false;
                if (xl >= 0 && xl < W) {
                    const uint32_t k = key[xl];
                    // a column nobody voted for holds cv2's initial value, the SCALED invalid disparity (minD - 1) * 16,
                    // which passes cv2's "disp2 >= minD" test when minD >= 2 (harmless quirk at minD = 0: -16 < 0)
                    const int d2 = k != 0xffffffffu ? (int)(0xffffu - (k & 0xffffu)) + x0 - xl : inv;
                    t = d2 >= minD && abs(d2 - dl) > maxdiff;
                }
                bad = bad && t;
            }
            {
                bool t = false;
                if (xh >= 0 && xh < W) {
                    const uint32_t k = key[xh];
                    // a column nobody voted for holds cv2's initial value, the SCALED invalid disparity (minD - 1) * 16,
                    // which passes cv2's "disp2 >= minD" test when minD >= 2 (harmless quirk at minD = 0: -16 < 0)
                    const int d2 = k != 0xffffffffu ? (int)(0xffffu - (k & 0xffffu)) + x0 - xh : inv;
                    t = d2 >= minD && abs(d2 - dh) > maxdiff;
                }
                bad = bad && t;
            }
            if (bad) d1 = inv;
        }
        raw[X] = (int16_t)d1;
    }
}

// ---------------------------------------------------------------------------------------------
// 3x3 median, replicate border, invalid values take part (cv2.medianBlur on CV_16S).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(int& a, int& b) { const int t = min(a, b); b = max(a, b); a = t; }

__global__ void __launch_bounds__(256)
k_median3(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t dpitch_e, size_t dstride_e, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    src += (size_t)b * W * H;
    int v[9];
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
        const int16_t* r = src + (size_t)min(max(y + dy, 0), H - 1) * W;
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) v[(dy + 1) * 3 + dx + 1] = __ldg(r + min(max(x + dx, 0), W - 1));
    }
    // 19-exchange median-of-9 network
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
    cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
    cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
    cswap(v[4], v[2]);
    dst[(size_t)b * dstride_e + (size_t)y * dpitch_e + x] = (int16_t)v[4];
}

// The same median on 8 pixels per thread: 128-bit loads and stores, the exchange network on packed
// int16 pairs (two pixels per VIMNMX.S16x2).  Needs W % 8 == 0 and 16-byte aligned rows on both sides.
__device__ __forceinline__ void cswap2(uint32_t& a, uint32_t& b)
{
    const uint32_t t = __vmins2(a, b);
    b = __vmaxs2(a, b);
    a = t;
}

__global__ void __launch_bounds__(256)
k_median3_v8(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t dpitch_e, size_t dstride_e, int W, int H,
             int n_groups)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups) return;
    const int gpr = W >> 3;                          // groups per row
    const int row = t / gpr, x0 = (t - row * gpr) << 3;
    const int b = row / H, y = row - b * H;
    const int16_t* img = src + (size_t)b * W * H;
    uint32_t v[4][9];                                // [pixel pair][3 rows x (left, centre, right)]
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
        const int16_t* r = img + (size_t)min(max(y + dy, 0), H - 1) * W;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(r + x0));
        const uint32_t w[4] = { q.x, q.y, q.z, q.w };
        const uint32_t lw = (uint32_t)(uint16_t)__ldg(r + max(x0 - 1, 0)) << 16;      // left neighbour in the high half
        const uint32_t rw = (uint32_t)(uint16_t)__ldg(r + min(x0 + 8, W - 1));        // right neighbour in the low half
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k][(dy + 1) * 3 + 0] = __byte_perm(k ? w[k - 1] : lw, w[k], 0x5432);   // (p[x-1], p[x])
            v[k][(dy + 1) * 3 + 1] = w[k];                                            // (p[x],   p[x+1])
            v[k][(dy + 1) * 3 + 2] = __byte_perm(w[k], k < 3 ? w[k + 1] : rw, 0x5432); // (p[x+1], p[x+2])
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t (&a)[9] = v[k];
        cswap2(a[1], a[2]); cswap2(a[4], a[5]); cswap2(a[7], a[8]);
        cswap2(a[0], a[1]); cswap2(a[3], a[4]); cswap2(a[6], a[7]);
        cswap2(a[1], a[2]); cswap2(a[4], a[5]); cswap2(a[7], a[8]);
        cswap2(a[0], a[3]); cswap2(a[5], a[8]); cswap2(a[4], a[7]);
        cswap2(a[3], a[6]); cswap2(a[1], a[4]); cswap2(a[2], a[5]);
        cswap2(a[4], a[7]); cswap2(a[4], a[2]); cswap2(a[6], a[4]);
        cswap2(a[4], a[2]);
        o[k] = a[4];
    }
    *reinterpret_cast<uint4*>(dst + (size_t)b * dstride_e + (size_t)y * dpitch_e + x0) = make_uint4(o[0], o[1], o[2], o[3]);
}

// true when 8-pixel (16-byte) vector accesses are legal on an int16 image with these pitches
__host__ inline bool vec8_ok(const void* p, size_t pitch_e, size_t stride_e, int W)
{
    return (W & 7) == 0 && (pitch_e & 7) == 0 && (stride_e & 7) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// ---------------------------------------------------------------------------------------------
// Speckle filter = connected components (4-neighbour, edge iff both valid and |a-b| <= maxDiff)
// with a size threshold.  Lock-free union-find on pixel indices; sizes are kept per horizontal RUN (at the
// run's first pixel) and added to the component's root once per run, so only run heads ever walk the
// union-find forest.  Order independent, like cv2.filterSpeckles.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* L, int i)
{
    V3D_DASSERT(i >= 0);
    int p = L[i];
    // parents only ever decrease (atomicMin), so a chain is strictly descending: it cannot cycle or leave the frame upwards
    while (p != i) { V3D_DASSERT(p >= 0 && p < i); i = p; p = L[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b)
{
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); V3D_DASSERT(old >= 0 && old <= b); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); V3D_DASSERT(old >= 0 && old <= a); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

__device__ __forceinline__ bool ccl_edge(int a, int b, int maxDiff, int inv)
{
    return a != inv && b != inv && abs(a - b) <= maxDiff;
}

// Pass 1, one block per image row: every valid pixel is labelled with the index of the first pixel
// of its horizontal run (inclusive max-scan of "run starts here" positions), so that union-find
// chains start one hop deep instead of a row long.
// labels are indices into the whole batch buffer (frame b occupies [b*n, (b+1)*n))
__global__ void __launch_bounds__(256)
k_ccl_rows(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int* __restrict__ L, int inv,
           int* __restrict__ sizes, int W, int H, int maxDiff)
{
    __shared__ int wmax[8];
    __shared__ int carry_s;
    const int y = blockIdx.x % H, b = blockIdx.x / H;
    const int16_t* d = disp + (size_t)b * dstride_e + (size_t)y * dpitch_e;
    const int rowbase = (b * H + y) * W;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = -1;
    __syncthreads();
    for (int x0 = 0; x0 < W; x0 += 256) {
        const int x = x0 + threadIdx.x;
        int start = -1;
        if (x < W) {
            const int v = d[x];
            const bool left = x > 0 && ccl_edge(v, d[x - 1], maxDiff, inv);
            if (!left) start = x;                     // a run (or an invalid pixel) starts here
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(V3D_FULL_MASK, start, o);
            if (lane >= o) start = max(start, t);
        }
        if (lane == 31) wmax[wid] = start;
        __syncthreads();
        int pre = carry_s;
        for (int w = 0; w < wid; w++) pre = max(pre, wmax[w]);
        start = max(start, pre);
        if (x < W) L[rowbase + x] = rowbase + start;
        // run length, accumulated at the run head (sizes[] was zeroed before the launch): one atomic per
        // stretch of equal `start` inside a warp; invalid pixels are their own start and count nothing
        {
            const bool valid = x < W && d[x] != inv;
            const int key = valid ? start : -2 - (int)threadIdx.x;
            const int prev = __shfl_up_sync(V3D_FULL_MASK, key, 1);
            const bool head = lane == 0 || key != prev;
            const unsigned heads = __ballot_sync(V3D_FULL_MASK, head);
            if (head && valid) {
                const unsigned later = lane == 31 ? 0u : (heads >> (lane + 1));
                const int len = later ? __ffs(later) : 32 - lane;
                atomicAdd(&sizes[rowbase + start], len);
            }
        }
        __syncthreads();
        if (threadIdx.x == 255) carry_s = start;
        __syncthreads();
    }
}

// Pass 1 on 8 pixels per thread (W % 8 == 0, W <= 2048: the whole row in one step, one block scan instead of
// one per 256 pixels).  Threads whose 8 pixels lie in one run aggregate their length over the warp (one atomic
// per stretch of such threads); the others add their own stretches.
__global__ void __launch_bounds__(256)
k_ccl_rows_v8(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int* __restrict__ L, int inv,
              int* __restrict__ sizes, int W, int H, int maxDiff)
{
    __shared__ int wmax[8];
    const int y = blockIdx.x % H, b = blockIdx.x / H;
    const int16_t* d = disp + (size_t)b * dstride_e + (size_t)y * dpitch_e;
    const int rowbase = (b * H + y) * W;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x0 = threadIdx.x * 8;
    const bool act = x0 < W;
    int v[8], st[8];
    int cur = -1;                                    // -1: still inside the run that entered from the left
    if (act) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(d + x0));
        const uint32_t w[4] = { q.x, q.y, q.z, q.w };
        int pl = x0 > 0 ? (int)__ldg(d + x0 - 1) : inv;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = (int16_t)(w[i >> 1] >> ((i & 1) * 16));
            if (!ccl_edge(v[i], pl, maxDiff, inv)) cur = x0 + i;      // a run (or an invalid pixel) starts here
            st[i] = cur;
            pl = v[i];
        }
    }
    // exclusive max-scan of every thread's last start over the block
    int incl = cur;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(V3D_FULL_MASK, incl, o);
        if (lane >= o) incl = max(incl, t);
    }
    if (lane == 31) wmax[wid] = incl;
    int pre = __shfl_up_sync(V3D_FULL_MASK, incl, 1);
    if (lane == 0) pre = -1;
    __syncthreads();
    for (int w = 0; w < wid; w++) pre = max(pre, wmax[w]);
    bool uniform = act;                              // all 8 pixels valid and in one run
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (st[i] < 0) st[i] = pre;
            uniform = uniform && v[i] != inv && st[i] == st[0];
        }
        int4* lp = reinterpret_cast<int4*>(L + rowbase + x0);
        lp[0] = make_int4(rowbase + st[0], rowbase + st[1], rowbase + st[2], rowbase + st[3]);
        lp[1] = make_int4(rowbase + st[4], rowbase + st[5], rowbase + st[6], rowbase + st[7]);
    }
    // run lengths, accumulated at the run heads (sizes[] was zeroed before the launch)
    const int key = uniform ? st[0] : -2 - (int)threadIdx.x;
    const int prev = __shfl_up_sync(V3D_FULL_MASK, key, 1);
    const bool head = lane == 0 || key != prev;
    const unsigned heads = __ballot_sync(V3D_FULL_MASK, head);
    if (uniform) {
        if (head) {
            const unsigned later = lane == 31 ? 0u : (heads >> (lane + 1));
            const int cnt = later ? __ffs(later) : 32 - lane;
            atomicAdd(&sizes[rowbase + st[0]], 8 * cnt);
        }
    } else if (act) {
        int s = -1, len = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (v[i] == inv) continue;
            if (st[i] != s) {
                if (len) atomicAdd(&sizes[rowbase + s], len);
                s = st[i]; len = 0;
            }
            len++;
        }
        if (len) atomicAdd(&sizes[rowbase + s], len);
    }
}

// Pass 2: vertical unions.  A union is skipped when the pixel's left neighbour already carries it
// (x-1,y)~(x,y), (x-1,y-1)~(x,y-1) and (x-1,y)~(x-1,y-1) all hold.
__global__ void __launch_bounds__(256)
k_ccl_merge(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int* __restrict__ L, int inv,
            int W, int H, int maxDiff)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W || y == 0) return;
    const int16_t* d = disp + (size_t)b * dstride_e;
    const int v = d[(size_t)y * dpitch_e + x];
    const int u = d[(size_t)(y - 1) * dpitch_e + x];
    if (!ccl_edge(v, u, maxDiff, inv)) return;
    if (x > 0) {
        const int vl = d[(size_t)y * dpitch_e + x - 1], ul = d[(size_t)(y - 1) * dpitch_e + x - 1];
        if (ccl_edge(v, vl, maxDiff, inv) && ccl_edge(u, ul, maxDiff, inv) && ccl_edge(vl, ul, maxDiff, inv)) return;
    }
    const int i = (b * H + y) * W + x;
    uf_union(L, i, i - W);
}

// The same pass on 8 pixels per thread (128-bit loads of both rows; the left neighbours of a pixel are the
// previous loop iteration's values).
__global__ void __launch_bounds__(256)
k_ccl_merge_v8(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int* __restrict__ L, int inv,
               int W, int H, int maxDiff, int n_groups)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups) return;
    const int gpr = W >> 3;
    const int row = t / gpr, x0 = (t - row * gpr) << 3;
    const int b = row / H, y = row - b * H;
    if (y == 0) return;
    const int16_t* r1 = disp + (size_t)b * dstride_e + (size_t)y * dpitch_e + x0;
    const int16_t* r0 = r1 - dpitch_e;
    const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(r1)), q0 = __ldg(reinterpret_cast<const uint4*>(r0));
    const uint32_t w1[4] = { q1.x, q1.y, q1.z, q1.w }, w0[4] = { q0.x, q0.y, q0.z, q0.w };
    bool has_left = x0 > 0;
    int vl = has_left ? (int)__ldg(r1 - 1) : inv, ul = has_left ? (int)__ldg(r0 - 1) : inv;
    const int base = row * W + x0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int v = (int16_t)(w1[i >> 1] >> ((i & 1) * 16)), u = (int16_t)(w0[i >> 1] >> ((i & 1) * 16));
        if (ccl_edge(v, u, maxDiff, inv)) {
            const bool covered = has_left && ccl_edge(v, vl, maxDiff, inv) && ccl_edge(u, ul, maxDiff, inv) && ccl_edge(vl, ul, maxDiff, inv);
            if (!covered) uf_union(L, base + i, base + i - W);
        }
        vl = v; ul = u; has_left = true;
    }
}

// Pass 3: every run head that is not its component's root adds its run length to the root (and is flattened
// onto it).  sizes[i] != 0 exactly at run heads; nobody adds to a non-root, so reading it here is race free.
__global__ void __launch_bounds__(256)
k_ccl_count(int* __restrict__ L, int* __restrict__ sizes, int n_total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const int len = sizes[i];
    if (len == 0) return;
    const int root = uf_find(L, i);
    if (root != i) {
        L[i] = root;                           // roots never change after the merge kernel
        atomicAdd(&sizes[root], len);
    }
}

__global__ void __launch_bounds__(256)
k_ccl_apply(int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, const int* __restrict__ L,
            const int* __restrict__ sizes, int W, int H, int maxSize, int newVal)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    int16_t* p = disp + (size_t)b * dstride_e + (size_t)y * dpitch_e + x;
    if (*p == newVal) return;                        // newVal is the invalid value itself
    const int root = L[L[(b * H + y) * W + x]];      // pixel -> run head -> root (heads were flattened by k_ccl_count)
    if (sizes[root] <= maxSize) *p = (int16_t)newVal;
}

// 8 pixels per thread; pixels of one horizontal run share their label, so the two dependent look-ups
// (run head -> root, root -> size) happen once per run instead of once per pixel.
__global__ void __launch_bounds__(256)
k_ccl_apply_v8(int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, const int* __restrict__ L,
               const int* __restrict__ sizes, int W, int H, int maxSize, int newVal, int n_groups)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups) return;
    const int gpr = W >> 3;
    const int row = t / gpr, x0 = (t - row * gpr) << 3;
    const int b = row / H, y = row - b * H;
    uint4* p = reinterpret_cast<uint4*>(disp + (size_t)b * dstride_e + (size_t)y * dpitch_e + x0);
    const uint4 q = *p;
    uint32_t w[4] = { q.x, q.y, q.z, q.w };
    const uint32_t inv2 = ((uint32_t)(uint16_t)newVal << 16) | (uint16_t)newVal;   // newVal is the invalid value itself
    if (q.x == inv2 && q.y == inv2 && q.z == inv2 && q.w == inv2) return;
    const int4* lp = reinterpret_cast<const int4*>(L + (size_t)row * W + x0);      // row * W + x0 is a multiple of 8
    const int4 l0 = __ldg(lp), l1 = __ldg(lp + 1);
    const int lab[8] = { l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w };
    int prev = -1;
    bool kill = false, changed = false;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int16_t v = (int16_t)(w[i >> 1] >> ((i & 1) * 16));
        if (v == (int16_t)newVal) continue;
        if (lab[i] != prev) {
            prev = lab[i];
            kill = sizes[L[prev]] <= maxSize;      // run head -> root (flattened by k_ccl_count) -> component size
        }
        if (kill) {
            w[i >> 1] = (i & 1) ? (w[i >> 1] & 0x0000ffffu) | ((uint32_t)(uint16_t)newVal << 16)
                                : (w[i >> 1] & 0xffff0000u) | (uint16_t)newVal;
            changed = true;
        }
    }
    if (changed) *p = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------------
// Epilogue.  float32(disp)/16 with <=0 -> 0 is max(disp,0)/16, exact, so the per-frame min/max of
// the float map are the min/max of max(disp,0) taken in integers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_minmax_init(int* mm, int batch)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) { mm[2 * i] = 0x7fffffff; mm[2 * i + 1] = -0x7fffffff - 1; }
}

__global__ void __launch_bounds__(256)
k_minmax(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int W, int H, int* __restrict__ mm)
{
    const int b = blockIdx.y;
    const int16_t* d = disp + (size_t)b * dstride_e;
    int lo = 0x7fffffff, hi = -0x7fffffff;
    const int n = W * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        const int v = max((int)d[(size_t)y * dpitch_e + x], 0);
        lo = min(lo, v); hi = max(hi, v);
    }
    lo = __reduce_min_sync(V3D_FULL_MASK, lo);
    hi = __reduce_max_sync(V3D_FULL_MASK, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * b], lo); atomicMax(&mm[2 * b + 1], hi); }
}

__global__ void __launch_bounds__(256)
k_epilogue(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int W, int H,
           const int* __restrict__ mm, float* __restrict__ f32, uint16_t* __restrict__ u16, int fixed, float lo, float hi)
{
    const int b = blockIdx.y;
    const int n = W * H;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int y = i / W, x = i - y * W;
    const int v = max((int)disp[(size_t)b * dstride_e + (size_t)y * dpitch_e + x], 0);
    const float f = __fdiv_rn((float)v, 16.0f);
    if (f32) f32[(size_t)b * n + i] = f;
    if (u16 && fixed) {      // opt-in clip-level scale (v3d_set_depth_scale)
        const float t = __fdiv_rn(__fsub_rn(f, lo), __fsub_rn(hi, lo));
        u16[(size_t)b * n + i] = (uint16_t)__fmul_rn(fminf(fmaxf(t, 0.0f), 1.0f), 65535.0f);
    } else if (u16) {
        const float mn = __fdiv_rn((float)mm[2 * b], 16.0f), mx = __fdiv_rn((float)mm[2 * b + 1], 16.0f);
        uint16_t o = 0;
        if (mx > mn) {
            // three separately rounded IEEE operations, like numpy (depth.py:401)
            const float t = __fmul_rn(__fdiv_rn(__fsub_rn(f, mn), __fsub_rn(mx, mn)), 65535.0f);
            o = (uint16_t)t;
        }
        u16[(size_t)b * n + i] = o;
    }
}

// 8 pixels per thread (128-bit loads); max(disp, 0) and the running min / max on packed int16 pairs.
__global__ void __launch_bounds__(256)
k_minmax_v8(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int W, int H, int* __restrict__ mm)
{
    const int b = blockIdx.y;
    const int16_t* d = disp + (size_t)b * dstride_e;
    const int gpr = W >> 3, n_groups = gpr * H;
    uint32_t lo = 0x7fff7fffu, hi = 0u;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_groups; t += gridDim.x * blockDim.x) {
        const int y = t / gpr, x0 = (t - y * gpr) << 3;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(d + (size_t)y * dpitch_e + x0));
        const uint32_t a = __vmaxs2(q.x, 0u), c = __vmaxs2(q.y, 0u), e = __vmaxs2(q.z, 0u), g = __vmaxs2(q.w, 0u);
        lo = __vminu2(lo, __vminu2(__vminu2(a, c), __vminu2(e, g)));      // non-negative from here: unsigned order
        hi = __vmaxu2(hi, __vmaxu2(__vmaxu2(a, c), __vmaxu2(e, g)));
    }
    int l = (int)min(lo & 0xffffu, lo >> 16), h = (int)max(hi & 0xffffu, hi >> 16);
    if (lo == 0x7fff7fffu && hi == 0u) { l = 0x7fffffff; h = -0x7fffffff; }     // this thread saw no pixel
    l = __reduce_min_sync(V3D_FULL_MASK, l);
    h = __reduce_max_sync(V3D_FULL_MASK, h);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * b], l); atomicMax(&mm[2 * b + 1], h); }
}

__global__ void __launch_bounds__(256)
k_epilogue_v8(const int16_t* __restrict__ disp, size_t dpitch_e, size_t dstride_e, int W, int H,
              const int* __restrict__ mm, float* __restrict__ f32, uint16_t* __restrict__ u16, int fixed, float lo, float hi)
{
    const int b = blockIdx.y;
    const int gpr = W >> 3;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= gpr * H) return;
    const int y = t / gpr, x0 = (t - y * gpr) << 3;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(disp + (size_t)b * dstride_e + (size_t)y * dpitch_e + x0));
    const uint32_t w[4] = { q.x, q.y, q.z, q.w };
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int v = max((int)(int16_t)(w[i >> 1] >> ((i & 1) * 16)), 0);
        f[i] = __fdiv_rn((float)v, 16.0f);
    }
    const size_t o = (size_t)b * W * H + (size_t)y * W + x0;       // dense outputs; a multiple of 8
    if (f32) {
        float4* fp = reinterpret_cast<float4*>(f32 + o);
        fp[0] = make_float4(f[0], f[1], f[2], f[3]);
        fp[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    if (!u16) return;
    uint32_t r[8];
    if (fixed) {             // opt-in clip-level scale (v3d_set_depth_scale)
        const float den = __fsub_rn(hi, lo);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float tt = __fdiv_rn(__fsub_rn(f[i], lo), den);
            r[i] = (uint16_t)__fmul_rn(fminf(fmaxf(tt, 0.0f), 1.0f), 65535.0f);
        }
    } else {
        const float mn = __fdiv_rn((float)mm[2 * b], 16.0f), mx = __fdiv_rn((float)mm[2 * b + 1], 16.0f);
        const float den = __fsub_rn(mx, mn);
#pragma unroll
        for (int i = 0; i < 8; i++)     // three separately rounded IEEE operations, like numpy (depth.py:401)
            r[i] = mx > mn ? (uint32_t)(uint16_t)__fmul_rn(__fdiv_rn(__fsub_rn(f[i], mn), den), 65535.0f) : 0u;
    }
    *reinterpret_cast<uint4*>(u16 + o) = make_uint4(r[0] | (r[1] << 16), r[2] | (r[3] << 16), r[4] | (r[5] << 16),
                                                    r[6] | (r[7] << 16));
}

// save_depth_map on an arbitrary float map (depth.py:397-403): per-frame min/max then the same three
// rounded operations.  Floats are ordered through the usual sign-flip integer key.
__device__ __forceinline__ int fkey(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float funkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
k_minmax_f32(const float* __restrict__ in, size_t n, int* __restrict__ mm)
{
    const int b = blockIdx.y;
    const float* d = in + (size_t)b * n;
    int lo = 0x7fffffff, hi = -0x7fffffff - 1;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = fkey(d[i]);
        lo = min(lo, k); hi = max(hi, k);
    }
    lo = __reduce_min_sync(V3D_FULL_MASK, lo);
    hi = __reduce_max_sync(V3D_FULL_MASK, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * b], lo); atomicMax(&mm[2 * b + 1], hi); }
}

__global__ void __launch_bounds__(256)
k_normalize_f32(const float* __restrict__ in, size_t n, const int* __restrict__ mm, uint16_t* __restrict__ out,
                int fixed, float lo, float hi)
{
    const int b = blockIdx.y;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fixed) {             // opt-in clip-level scale (v3d_set_depth_scale)
        const float t = __fdiv_rn(__fsub_rn(in[(size_t)b * n + i], lo), __fsub_rn(hi, lo));
        out[(size_t)b * n + i] = (uint16_t)__fmul_rn(fminf(fmaxf(t, 0.0f), 1.0f), 65535.0f);
        return;
    }
    const float mn = funkey(mm[2 * b]), mx = funkey(mm[2 * b + 1]);
    uint16_t o = 0;
    if (mx > mn) o = (uint16_t)__fmul_rn(__fdiv_rn(__fsub_rn(in[(size_t)b * n + i], mn), __fsub_rn(mx, mn)), 65535.0f);
    out[(size_t)b * n + i] = o;
}

}  // namespace

int v3d_launch_normalize_f32(v3d_ctx* ctx, const float* in, size_t n, int batch, uint16_t* out, cudaStream_t st)
{
    V3dScope scope(ctx, ST_POST, st);
    if (!ctx->fixed_scale) {
        k_minmax_init<<<(batch + 255) / 256, 256, 0, st>>>(ctx->minmax, batch);
        dim3 g(148 * 2, batch);
        k_minmax_f32<<<g, 256, 0, st>>>(in, n, ctx->minmax);
        V3D_LAUNCHED(ctx, 2);
    }
    dim3 grid((unsigned)((n + 255) / 256), batch);
    k_normalize_f32<<<grid, 256, 0, st>>>(in, n, ctx->minmax, out, ctx->fixed_scale, ctx->scale_lo, ctx->scale_hi);
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

int v3d_launch_select(v3d_ctx* ctx, int batch, cudaStream_t st)
{
    V3dScope scope(ctx, ST_SELECT, st);
    const size_t smem = (size_t)ctx->W * 4 + (size_t)ctx->W * 2 + 16;      // one row of votes and disparities
    if (smem > 48 * 1024 && !ctx->select_attr_set) {
        V3D_CUDA(cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->select_attr_set = 1;
    }
    k_select<<<batch * ctx->H, 256, smem, st>>>(ctx->rec, ctx->raw, ctx->W, ctx->W1, ctx->D, ctx->maxdiff, ctx->x0, ctx->minD, ctx->inv);
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

int v3d_launch_median(v3d_ctx* ctx, int batch, int16_t* dst, size_t dpitch, size_t dstride, cudaStream_t st)
{
    V3dScope scope(ctx, ST_MEDIAN, st);
    if (vec8_ok(dst, dpitch / 2, dstride / 2, ctx->W)) {
        const int n_groups = batch * ctx->H * (ctx->W / 8);
        k_median3_v8<<<(n_groups + 255) / 256, 256, 0, st>>>(ctx->raw, dst, dpitch / 2, dstride / 2, ctx->W, ctx->H, n_groups);
    } else {
        dim3 grid((ctx->W + 255) / 256, ctx->H, batch);
        k_median3<<<grid, 256, 0, st>>>(ctx->raw, dst, dpitch / 2, dstride / 2, ctx->W, ctx->H);
    }
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}

int v3d_launch_speckle(v3d_ctx* ctx, int batch, int16_t* disp, size_t dpitch, size_t dstride, cudaStream_t st)
{
    if (ctx->p.speckleWindowSize <= 0) return V3D_OK;
    V3dScope scope(ctx, ST_SPECKLE, st);
    const int W = ctx->W, H = ctx->H;
    const int maxDiff = 16 * ctx->p.speckleRange;
    dim3 grid((W + 255) / 256, H, batch);
    const int n_total = batch * W * H;
    V3D_CUDA(cudaMemsetAsync(ctx->sizes, 0, (size_t)n_total * sizeof(int), st));
    const bool vec = vec8_ok(disp, dpitch / 2, dstride / 2, W);
    if (vec && W <= 2048) k_ccl_rows_v8<<<batch * H, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, ctx->labels, ctx->inv, ctx->sizes, W, H, maxDiff);
    else k_ccl_rows<<<batch * H, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, ctx->labels, ctx->inv, ctx->sizes, W, H, maxDiff);
    const int n_groups = vec ? batch * H * (W / 8) : 0;
    if (vec) k_ccl_merge_v8<<<(n_groups + 255) / 256, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, ctx->labels, ctx->inv, W, H, maxDiff, n_groups);
    else k_ccl_merge<<<grid, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, ctx->labels, ctx->inv, W, H, maxDiff);
    k_ccl_count<<<(n_total + 255) / 256, 256, 0, st>>>(ctx->labels, ctx->sizes, n_total);
    if (vec) {
        k_ccl_apply_v8<<<(n_groups + 255) / 256, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, ctx->labels, ctx->sizes, W, H,
                                                               ctx->p.speckleWindowSize, ctx->inv, n_groups);
    } else {
        k_ccl_apply<<<grid, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, ctx->labels, ctx->sizes, W, H,
                                          ctx->p.speckleWindowSize, ctx->inv);
    }
    V3D_LAUNCHED(ctx, 4);
    return V3D_OK;
}

int v3d_launch_post(v3d_ctx* ctx, const int16_t* disp, size_t dpitch, size_t dstride, int batch,
                    float* f32, uint16_t* u16, cudaStream_t st)
{
    V3dScope scope(ctx, ST_POST, st);
    const int W = ctx->W, H = ctx->H, n = W * H;
    const bool vec = vec8_ok(disp, dpitch / 2, dstride / 2, W) && (!f32 || (reinterpret_cast<uintptr_t>(f32) & 15) == 0) &&
                     (!u16 || (reinterpret_cast<uintptr_t>(u16) & 15) == 0);
    if (u16 && !ctx->fixed_scale) {
        k_minmax_init<<<(batch + 255) / 256, 256, 0, st>>>(ctx->minmax, batch);
        dim3 g(148 * 2, batch);
        if (vec) k_minmax_v8<<<g, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, W, H, ctx->minmax);
        else k_minmax<<<g, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, W, H, ctx->minmax);
        V3D_LAUNCHED(ctx, 2);
    }
    if (vec) {
        dim3 grid((n / 8 + 255) / 256, batch);
        k_epilogue_v8<<<grid, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, W, H, ctx->minmax, f32, u16, ctx->fixed_scale,
                                            ctx->scale_lo, ctx->scale_hi);
    } else {
        dim3 grid((n + 255) / 256, batch);
        k_epilogue<<<grid, 256, 0, st>>>(disp, dpitch / 2, dstride / 2, W, H, ctx->minmax, f32, u16, ctx->fixed_scale,
                                         ctx->scale_lo, ctx->scale_hi);
    }
    V3D_LAUNCHED(ctx, 1);
    return V3D_OK;
}
