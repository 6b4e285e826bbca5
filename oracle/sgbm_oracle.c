/*
 * oracle/sgbm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the arithmetic the reference
 * reaches through its third-party dependency on the depth hot path:
 *
 *   reference call sites (under /root/reference/src/video_3d_pipeline/):
 *     depth.py:250-268   split_sbs_frame (+ cv2.resize INTER_LANCZOS4 unsqueeze)
 *     depth.py:274-275   cv2.cvtColor BGR2RGB
 *     depth.py:337-338   cv2.cvtColor RGB2GRAY
 *     depth.py:315-325   cv2.StereoSGBM_create(...)
 *     depth.py:341       stereo.compute(left_gray, right_gray)
 *     depth.py:374       disparity[disparity <= 0] = 0
 *     depth.py:397-406   save_depth_map (per-frame min-max -> uint16)
 *
 * The dependency is opencv-python (pinned 4.11.0.86 in uv.lock:1164-1165; this
 * image has opencv-python-headless 4.13.0.92).  Its source is NOT vendored in
 * /root/reference, so this file restates the published algorithm of OpenCV's
 * calib3d StereoSGBM (calcPixelCostBT, computeDisparitySGBM,
 * StereoSGBMImpl::compute), imgproc medianBlur / resize / cvtColor and calib3d
 * filterSpeckles, as specified in SURVEY.md Appendix A.  Parity is PINNED by
 * tests/test_oracle_vs_cv2.py, which asserts this restatement bit-equal to the
 * live cv2 on seeded inputs, and by the committed golden vectors under
 * tests/golden/ (generated from cv2 by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this file.  The product (libv3d.so) never links or calls it.
 *
 * Layouts: images are row-major uint8.  Cost volumes are [H][W1][D] uint16 with
 * d fastest, W1 = W - D (window columns, image column = window column + D).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int minDisparity;      /* only 0 is supported (depth.py:316) */
    int numDisparities;    /* D, multiple of 16 */
    int blockSize;         /* odd, >= 1 */
    int P1, P2;
    int disp12MaxDiff;
    int preFilterCap;
    int uniquenessRatio;
    int speckleWindowSize;
    int speckleRange;
    int mode;              /* 0 = MODE_SGBM (5 paths), 1 = MODE_HH (8 paths) */
} orc_params;

#define ORC_INV (-16)
#define ORC_OK 0
#define ORC_EINVAL (-1)
#define ORC_ENOMEM (-2)

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------ A.1 -- */

/* BGR -> gray as cvtColor(BGR2RGB) then cvtColor(RGB2GRAY) does it
 * (depth.py:274-275, 337-338): 15-bit fixed point. */
void orc_bgr_to_gray(const uint8_t* bgr, long npix, uint8_t* gray)
{
    for (long i = 0; i < npix; i++) {
        int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((9798 * r + 19235 * g + 3735 * b + 16384) >> 15);
    }
}

/* cv2.resize(eye, (2*w, h), INTER_LANCZOS4) of depth.py:265-266: horizontal x2,
 * vertical identity.  8 taps in 11-bit fixed point, replicate border. */
void orc_unsqueeze_x2(const uint8_t* src, int h, int w, int ch, long src_pitch,
                      uint8_t* dst, long dst_pitch)
{
    static const int tap_even[8] = { -8, 64, -188, 579, 1830, -312, 114, -31 };
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (long)y * src_pitch;
        uint8_t* o = dst + (long)y * dst_pitch;
        for (int dx = 0; dx < 2 * w; dx++) {
            int odd = dx & 1;
            int sx = odd ? (dx - 1) / 2 : dx / 2 - 1;
            for (int c = 0; c < ch; c++) {
                int acc = 0;
                for (int k = 0; k < 8; k++) {
                    int t = odd ? tap_even[7 - k] : tap_even[k];
                    int xs = iclamp(sx - 3 + k, 0, w - 1);
                    acc += t * s[xs * ch + c];
                }
                o[dx * ch + c] = (uint8_t)iclamp((acc + 1024) >> 11, 0, 255);
            }
        }
    }
}

/* Split an SBS BGR frame into two gray eyes, the chain of depth.py:250-268,
 * 274-275, 337-338.  eye width = W_sbs/2 (or W_sbs when unsqueeze). */
int orc_split_gray(const uint8_t* sbs_bgr, int H, int W_sbs, int unsqueeze,
                   uint8_t* left_gray, uint8_t* right_gray)
{
    if (W_sbs & 1) return ORC_EINVAL;          /* depth.py:254-255 */
    int half = W_sbs / 2;
    int We = unsqueeze ? 2 * half : half;
    uint8_t* tmp = (uint8_t*)malloc((size_t)We * 3);
    if (!tmp) return ORC_ENOMEM;
    for (int eye = 0; eye < 2; eye++) {
        uint8_t* out = eye ? right_gray : left_gray;
        for (int y = 0; y < H; y++) {
            const uint8_t* row = sbs_bgr + ((long)y * W_sbs + (long)eye * half) * 3;
            if (unsqueeze) {
                orc_unsqueeze_x2(row, 1, half, 3, 0, tmp, 0);
                orc_bgr_to_gray(tmp, We, out + (long)y * We);
            } else {
                orc_bgr_to_gray(row, We, out + (long)y * We);
            }
        }
    }
    free(tmp);
    return ORC_OK;
}

/* ------------------------------------------------------------------ A.2 -- */

/* Per-image prefilter rows (calcPixelCostBT's row preparation): clipped
 * x-Sobel and raw intensity, both with columns 0 and W-1 forced to ftzero. */
void orc_prefilter(const uint8_t* img, int W, int H, int preFilterCap,
                   uint8_t* sob, uint8_t* inten)
{
    int ftzero = imax(preFilterCap, 15) | 1;
    for (int y = 0; y < H; y++) {
        const uint8_t* r1 = img + (long)y * W;
        const uint8_t* r0 = img + (long)(y > 0 ? y - 1 : y) * W;
        const uint8_t* r2 = img + (long)(y < H - 1 ? y + 1 : y) * W;
        uint8_t* so = sob + (long)y * W;
        uint8_t* io = inten + (long)y * W;
        for (int x = 0; x < W; x++) {
            if (x == 0 || x == W - 1) {
                so[x] = (uint8_t)ftzero;
                io[x] = (uint8_t)ftzero;
            } else {
                int g = 2 * (r1[x + 1] - r1[x - 1]) + (r0[x + 1] - r0[x - 1]) + (r2[x + 1] - r2[x - 1]);
                so[x] = (uint8_t)(iclamp(g, -ftzero, ftzero) + ftzero);
                io[x] = r1[x];
            }
        }
    }
}

/* Birchfield-Tomasi half-pixel interval of a row. */
static void bt_interval(const uint8_t* p, int W, uint8_t* lo, uint8_t* hi)
{
    for (int x = 0; x < W; x++) {
        int v = p[x];
        int l = x > 0 ? (v + p[x - 1]) / 2 : v;
        int r = x < W - 1 ? (v + p[x + 1]) / 2 : v;
        lo[x] = (uint8_t)imin(v, imin(l, r));
        hi[x] = (uint8_t)imax(v, imax(l, r));
    }
}

/* Block-summed pixel cost Cb[H][W1][D] (uint16). */
int orc_cost_volume(const uint8_t* left, const uint8_t* right, int W, int H,
                    const orc_params* p, uint16_t* Cb)
{
    int D = p->numDisparities, W1 = W - D, r = p->blockSize / 2;
    if (p->minDisparity != 0 || D <= 0 || (D % 16) || W1 <= r || p->blockSize < 1 || !(p->blockSize & 1))
        return ORC_EINVAL;
    size_t plane = (size_t)W * H;
    uint8_t* pf = (uint8_t*)malloc(plane * 4);
    uint8_t* lohi = (uint8_t*)malloc((size_t)W * 8);
    size_t vol = (size_t)H * W1 * D;
    uint8_t* pix = (uint8_t*)malloc(vol);
    uint16_t* hs = (uint16_t*)malloc(vol * sizeof(uint16_t));
    if (!pf || !lohi || !pix || !hs) { free(pf); free(lohi); free(pix); free(hs); return ORC_ENOMEM; }
    uint8_t *sobL = pf, *intL = pf + plane, *sobR = pf + 2 * plane, *intR = pf + 3 * plane;
    orc_prefilter(left, W, H, p->preFilterCap, sobL, intL);
    orc_prefilter(right, W, H, p->preFilterCap, sobR, intR);

    for (int y = 0; y < H; y++) {
        uint8_t* row = pix + (size_t)y * W1 * D;
        memset(row, 0, (size_t)W1 * D);
        for (int ch = 0; ch < 2; ch++) {
            const uint8_t* u = (ch ? intL : sobL) + (long)y * W;
            const uint8_t* v = (ch ? intR : sobR) + (long)y * W;
            uint8_t *ulo = lohi, *uhi = lohi + W, *vlo = lohi + 2 * W, *vhi = lohi + 3 * W;
            bt_interval(u, W, ulo, uhi);
            bt_interval(v, W, vlo, vhi);
            int shift = ch ? 2 : 0;
            for (int x = 0; x < W1; x++) {
                int X = x + D;
                for (int d = 0; d < D; d++) {
                    int xr = X - d;
                    int c0 = imax(0, imax(u[X] - vhi[xr], vlo[xr] - u[X]));
                    int c1 = imax(0, imax(v[xr] - uhi[X], ulo[X] - v[xr]));
                    row[(size_t)x * D + d] += (uint8_t)(imin(c0, c1) >> shift);
                }
            }
        }
    }
    /* horizontal sum, clamp in WINDOW coordinates */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W1; x++) {
            uint16_t* o = hs + ((size_t)y * W1 + x) * D;
            memset(o, 0, sizeof(uint16_t) * D);
            for (int dx = -r; dx <= r; dx++) {
                const uint8_t* s = pix + ((size_t)y * W1 + iclamp(x + dx, 0, W1 - 1)) * D;
                for (int d = 0; d < D; d++) o[d] = (uint16_t)(o[d] + s[d]);
            }
        }
    /* vertical sum, clamp to image rows */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W1; x++) {
            uint16_t* o = Cb + ((size_t)y * W1 + x) * D;
            memset(o, 0, sizeof(uint16_t) * D);
            for (int dy = -r; dy <= r; dy++) {
                const uint16_t* s = hs + ((size_t)iclamp(y + dy, 0, H - 1) * W1 + x) * D;
                for (int d = 0; d < D; d++) o[d] = (uint16_t)(o[d] + s[d]);
            }
        }
    free(pf); free(lohi); free(pix); free(hs);
    return ORC_OK;
}

/* ------------------------------------------------------------------ A.3 -- */

/* One path direction: predecessor of (x,y) is (x+dx, y+dy).  Adds L into S32.
 * If Ldir != NULL the per-direction L volume is also written (debug tap). */
static int aggregate_dir(const uint16_t* Cb, int W1, int H, int D, int P1, int P2,
                         int dx, int dy, uint32_t* S32, uint16_t* Ldir)
{
    /* previous-row state (for dy != 0) and running state (for dy == 0) */
    int* Lprev = (int*)calloc((size_t)W1 * D, sizeof(int));
    int* Lcur = (int*)calloc((size_t)W1 * D, sizeof(int));
    int* mprev = (int*)calloc((size_t)W1, sizeof(int));
    int* mcur = (int*)calloc((size_t)W1, sizeof(int));
    if (!Lprev || !Lcur || !mprev || !mcur) { free(Lprev); free(Lcur); free(mprev); free(mcur); return ORC_ENOMEM; }
    int ystart = dy > 0 ? H - 1 : 0, yend = dy > 0 ? -1 : H, ystep = dy > 0 ? -1 : 1;
    for (int y = ystart; y != yend; y += ystep) {
        int have_prev_row = (dy != 0) && (y + dy >= 0) && (y + dy < H);
        int xstart = dx > 0 ? W1 - 1 : 0, xend = dx > 0 ? -1 : W1, xstep = dx > 0 ? -1 : 1;
        for (int x = xstart; x != xend; x += xstep) {
            const uint16_t* c = Cb + ((size_t)y * W1 + x) * D;
            int* L = Lcur + (size_t)x * D;
            int qx = x + dx;
            const int* Lq = NULL;
            int mq = 0;
            if (qx >= 0 && qx < W1) {
                if (dy == 0) { Lq = Lcur + (size_t)qx * D; mq = mcur[qx]; }
                else if (have_prev_row) { Lq = Lprev + (size_t)qx * D; mq = mprev[qx]; }
            }
            int m = 1 << 30;
            for (int d = 0; d < D; d++) {
                int best;
                if (Lq) {
                    int a = Lq[d];
                    int b = d > 0 ? Lq[d - 1] + P1 : 32767;
                    int e = d < D - 1 ? Lq[d + 1] + P1 : 32767;
                    best = imin(imin(a, b), imin(e, mq + P2)) - mq;
                } else {
                    /* predecessor outside the window: L(q,.) = 0, minL(q) = 0 */
                    best = 0;
                }
                int l = c[d] + best;
                L[d] = l;
                if (l < m) m = l;
            }
            mcur[x] = m;
            uint32_t* s = S32 + ((size_t)y * W1 + x) * D;
            for (int d = 0; d < D; d++) s[d] += (uint32_t)L[d];
            if (Ldir) {
                uint16_t* lo = Ldir + ((size_t)y * W1 + x) * D;
                for (int d = 0; d < D; d++) lo[d] = (uint16_t)L[d];
            }
        }
        if (dy != 0) {
            int* t = Lprev; Lprev = Lcur; Lcur = t;
            int* tm = mprev; mprev = mcur; mcur = tm;
        }
    }
    free(Lprev); free(Lcur); free(mprev); free(mcur);
    return ORC_OK;
}

/* cv2: P1 = P1 > 0 ? P1 : 2;  P2 = max(P2 > 0 ? P2 : 5, P1 + 1). */
static void effective_penalties(const orc_params* p, int* P1, int* P2)
{
    *P1 = p->P1 > 0 ? p->P1 : 2;
    *P2 = imax(p->P2 > 0 ? p->P2 : 5, *P1 + 1);
}

static const int ORC_DIRS[8][2] = {
    { -1, 0 }, { -1, -1 }, { 0, -1 }, { 1, -1 }, { 1, 0 },   /* MODE_SGBM */
    { 1, 1 }, { 0, 1 }, { -1, 1 }                            /* + MODE_HH */
};

/* S[H][W1][D] = min(32767, sum over directions of L).  S_unsat (optional,
 * uint16) receives the unsaturated sum (fits: <= 8*4725). */
int orc_aggregate(const uint16_t* Cb, int W1, int H, const orc_params* p,
                  uint16_t* S, uint16_t* S_unsat)
{
    int D = p->numDisparities;
    int ndirs = p->mode == 1 ? 8 : 5;
    if (p->mode != 0 && p->mode != 1) return ORC_EINVAL;
    size_t vol = (size_t)H * W1 * D;
    int P1, P2;
    effective_penalties(p, &P1, &P2);
    uint32_t* S32 = (uint32_t*)calloc(vol, sizeof(uint32_t));
    if (!S32) return ORC_ENOMEM;
    for (int k = 0; k < ndirs; k++) {
        int rc = aggregate_dir(Cb, W1, H, D, P1, P2,
                               ORC_DIRS[k][0], ORC_DIRS[k][1], S32, NULL);
        if (rc) { free(S32); return rc; }
    }
    for (size_t i = 0; i < vol; i++) {
        S[i] = (uint16_t)(S32[i] > 32767u ? 32767u : S32[i]);
        if (S_unsat) S_unsat[i] = (uint16_t)S32[i];
    }
    free(S32);
    return ORC_OK;
}

/* Single-direction tap used by the stage-level GPU parity tests. */
int orc_aggregate_one(const uint16_t* Cb, int W1, int H, const orc_params* p,
                      int dir_index, uint16_t* Ldir)
{
    int D = p->numDisparities;
    if (dir_index < 0 || dir_index > 7) return ORC_EINVAL;
    size_t vol = (size_t)H * W1 * D;
    int P1, P2;
    effective_penalties(p, &P1, &P2);
    uint32_t* S32 = (uint32_t*)calloc(vol, sizeof(uint32_t));
    if (!S32) return ORC_ENOMEM;
    int rc = aggregate_dir(Cb, W1, H, D, P1, P2,
                           ORC_DIRS[dir_index][0], ORC_DIRS[dir_index][1], S32, Ldir);
    free(S32);
    return rc;
}

/* ------------------------------------------------------------------ A.4 -- */

/* WTA + uniqueness + sub-pixel + disp2 vote + LR check.  disp is [H][W] int16. */
int orc_select(const uint16_t* S, int W, int H, const orc_params* p, int16_t* disp)
{
    int D = p->numDisparities, W1 = W - D;
    int uniq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    int maxdiff = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    int* disp2 = (int*)malloc(sizeof(int) * W);
    int* disp2cost = (int*)malloc(sizeof(int) * W);
    if (!disp2 || !disp2cost) { free(disp2); free(disp2cost); return ORC_ENOMEM; }
    for (int y = 0; y < H; y++) {
        int16_t* dr = disp + (long)y * W;
        for (int X = 0; X < W; X++) { dr[X] = ORC_INV; disp2[X] = ORC_INV; disp2cost[X] = 32767; }
        for (int x = W1 - 1; x >= 0; x--) {
            const uint16_t* s = S + ((size_t)y * W1 + x) * D;
            int best = 0, minS = s[0];
            for (int d = 1; d < D; d++) if (s[d] < minS) { minS = s[d]; best = d; }
            int d;
            for (d = 0; d < D; d++)
                if ((int)s[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
            if (d < D) continue;
            int x2 = x + D - best;
            if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = best; }
            int d16;
            if (best > 0 && best < D - 1) {
                int den = imax((int)s[best - 1] + (int)s[best + 1] - 2 * (int)s[best], 1);
                d16 = best * 16 + (((int)s[best - 1] - (int)s[best + 1]) * 16 + den) / (den * 2);
            } else {
                d16 = best * 16;
            }
            dr[x + D] = (int16_t)d16;
        }
        for (int X = D; X < W; X++) {
            int d1 = dr[X];
            if (d1 == ORC_INV) continue;
            int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
            int _x = X - _d, x_ = X - d_;
            if (0 <= _x && _x < W && disp2[_x] >= 0 && abs(disp2[_x] - _d) > maxdiff &&
                0 <= x_ && x_ < W && disp2[x_] >= 0 && abs(disp2[x_] - d_) > maxdiff)
                dr[X] = ORC_INV;
        }
    }
    free(disp2); free(disp2cost);
    return ORC_OK;
}

/* ------------------------------------------------------------------ A.5 -- */

static inline void sort2(int16_t* a, int16_t* b) { if (*a > *b) { int16_t t = *a; *a = *b; *b = t; } }

void orc_median3(const int16_t* src, int W, int H, int16_t* dst)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int16_t v[9];
            int k = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    v[k++] = src[(long)iclamp(y + dy, 0, H - 1) * W + iclamp(x + dx, 0, W - 1)];
            /* insertion sort is plenty for 9 values */
            for (int i = 1; i < 9; i++)
                for (int j = i; j > 0; j--) sort2(&v[j - 1], &v[j]);
            dst[(long)y * W + x] = v[4];
        }
}

/* filterSpeckles(disp, newVal, maxSize, maxDiff): 4-connected components of
 * pixels != newVal, edge iff |a-b| <= maxDiff; size <= maxSize -> newVal. */
int orc_filter_speckles(int16_t* disp, int W, int H, int newVal, int maxSize, int maxDiff)
{
    long n = (long)W * H;
    int* label = (int*)calloc((size_t)n, sizeof(int));
    long* stack = (long*)malloc(sizeof(long) * (size_t)n);
    long* members = (long*)malloc(sizeof(long) * (size_t)n);
    if (!label || !stack || !members) { free(label); free(stack); free(members); return ORC_ENOMEM; }
    int cur = 0;
    for (long i = 0; i < n; i++) {
        if (disp[i] == newVal || label[i]) continue;
        cur++;
        long sp = 0, cnt = 0;
        stack[sp++] = i; label[i] = cur;
        while (sp) {
            long q = stack[--sp];
            members[cnt++] = q;
            int qy = (int)(q / W), qx = (int)(q % W);
            int v = disp[q];
            long nb[4]; int nn = 0;
            if (qx + 1 < W) nb[nn++] = q + 1;
            if (qx > 0) nb[nn++] = q - 1;
            if (qy + 1 < H) nb[nn++] = q + W;
            if (qy > 0) nb[nn++] = q - W;
            for (int k = 0; k < nn; k++) {
                long t = nb[k];
                if (!label[t] && disp[t] != newVal && abs(v - disp[t]) <= maxDiff) {
                    label[t] = cur; stack[sp++] = t;
                }
            }
        }
        if (cnt <= maxSize) {
            /* Values may only be overwritten after the flood: members are all
             * labelled already so later floods never look at them again. */
            for (long k = 0; k < cnt; k++) label[members[k]] = -cur;
        }
    }
    for (long i = 0; i < n; i++) if (label[i] < 0) disp[i] = (int16_t)newVal;
    free(label); free(stack); free(members);
    return ORC_OK;
}

/* ------------------------------------------------- whole compute() ----- */

/* cv2.StereoSGBM.compute(left, right) -> int16 disparity x16, invalid = -16.
 * Optional taps (may be NULL): Cb, S (saturated), raw (pre-median), med
 * (post-median, pre-speckle). */
int orc_sgbm_compute(const uint8_t* left, const uint8_t* right, int W, int H,
                     const orc_params* p, int16_t* disp,
                     uint16_t* tapC, uint16_t* tapS, int16_t* tapRaw, int16_t* tapMed)
{
    int D = p->numDisparities, W1 = W - D;
    if (W1 <= p->blockSize / 2) return ORC_EINVAL;   /* cv2 raises here */
    size_t vol = (size_t)H * W1 * D;
    uint16_t* Cb = tapC ? tapC : (uint16_t*)malloc(vol * 2);
    uint16_t* S = tapS ? tapS : (uint16_t*)malloc(vol * 2);
    int16_t* raw = tapRaw ? tapRaw : (int16_t*)malloc((size_t)W * H * 2);
    int rc = ORC_ENOMEM;
    if (Cb && S && raw) {
        rc = orc_cost_volume(left, right, W, H, p, Cb);
        if (!rc) rc = orc_aggregate(Cb, W1, H, p, S, NULL);
        if (!rc) rc = orc_select(S, W, H, p, raw);
        if (!rc) {
            orc_median3(raw, W, H, disp);
            if (tapMed) memcpy(tapMed, disp, (size_t)W * H * 2);
            if (p->speckleWindowSize > 0)
                rc = orc_filter_speckles(disp, W, H, (p->minDisparity - 1) * 16,
                                         p->speckleWindowSize, 16 * p->speckleRange);
        }
    }
    if (!tapC) free(Cb);
    if (!tapS) free(S);
    if (!tapRaw) free(raw);
    return rc;
}

/* ------------------------------------------------------------------ A.6 -- */

/* depth.py:341,374: float32(disp)/16, <=0 -> 0. */
void orc_disp_to_float(const int16_t* disp, long n, float* out)
{
    for (long i = 0; i < n; i++) {
        float f = (float)disp[i] / 16.0f;
        out[i] = f <= 0.0f ? 0.0f : f;
    }
}

/* depth.py:397-403: ((d-min)/(max-min)*65535).astype(uint16) in fp32,
 * three separately rounded IEEE operations. */
void orc_normalize_u16(const float* d, long n, uint16_t* out)
{
    float mn = d[0], mx = d[0];
    for (long i = 1; i < n; i++) { if (d[i] < mn) mn = d[i]; if (d[i] > mx) mx = d[i]; }
    if (!(mx > mn)) { memset(out, 0, sizeof(uint16_t) * (size_t)n); return; }
    volatile float range = mx - mn;
    for (long i = 0; i < n; i++) {
        volatile float a = d[i] - mn;
        volatile float b = a / range;
        volatile float c = b * 65535.0f;
        out[i] = (uint16_t)c;
    }
}
