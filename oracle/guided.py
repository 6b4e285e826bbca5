"""Float64 definition of the guided 1080p -> 2160p depth upscale (test infrastructure).

PARITY UNPINNED BY THE REFERENCE: the reference advertises a guided filter
(readme.md:97,119) but ships none -- upscale.py:47-59 is ffmpeg ``scale`` -- and
the installed cv2 has no ``ximgproc``.  This file is therefore the normative
definition the CUDA kernels are tested against (tolerance: 0.5 LSB of the 16-bit
output on q, i.e. 0.5/65535 absolute, and <= 1 LSB on the rounded uint16).

Definition (He, Sun, Tang "Guided Image Filtering", colour-guide form):
  p   = bilinear upsample of depth_u16/65535 to the guide size, half-pixel
        centres, coordinates clamped to the source (cv2.resize INTER_LINEAR
        convention);
  I   = guide RGB uint8 / 255;
  box = normalised (2r+1)^2 mean, border REFLECT ``fedcba|abcdef``
        (numpy 'symmetric', cv2.BORDER_REFLECT);
  Sigma = box(I I^T) - box(I) box(I)^T + eps*Id;   cov = box(I p) - box(I) box(p)
  a = Sigma^-1 cov;  b = box(p) - a . box(I);  q = box(a) . I + box(b)
  out = floor(clip(q, 0, 1) * 65535 + 0.5)  as uint16.
"""
import numpy as np


def bilinear_upsample(src, out_h, out_w):
    """Half-pixel-centre bilinear resize, source coordinates clamped (float64)."""
    src = np.asarray(src, np.float64)
    h, w = src.shape

    def taps(n_out, n_in):
        c = (np.arange(n_out, dtype=np.float64) + 0.5) * (n_in / n_out) - 0.5
        i0 = np.floor(c)
        f = c - i0
        i0 = i0.astype(np.int64)
        return np.clip(i0, 0, n_in - 1), np.clip(i0 + 1, 0, n_in - 1), f

    y0, y1, fy = taps(out_h, h)
    x0, x1, fx = taps(out_w, w)
    top = src[y0][:, x0] * (1 - fx)[None, :] + src[y0][:, x1] * fx[None, :]
    bot = src[y1][:, x0] * (1 - fx)[None, :] + src[y1][:, x1] * fx[None, :]
    return top * (1 - fy)[:, None] + bot * fy[:, None]


def box_mean(a, r):
    """Normalised (2r+1)^2 box mean, REFLECT (symmetric) border, float64."""
    a = np.asarray(a, np.float64)
    k = 2 * r + 1
    pad = np.pad(a, ((r, r), (r, r)), mode="symmetric")
    cs = np.cumsum(pad, axis=0)
    cs = np.concatenate([np.zeros((1, cs.shape[1])), cs], axis=0)
    v = cs[k:] - cs[:-k]
    cs = np.cumsum(v, axis=1)
    cs = np.concatenate([np.zeros((cs.shape[0], 1)), cs], axis=1)
    return (cs[:, k:] - cs[:, :-k]) / float(k * k)


def guided_coefficients(p, I, r, eps):
    """Per-pixel linear coefficients a (H,W,3) and b (H,W) of the colour guided filter."""
    mI = [box_mean(I[..., c], r) for c in range(3)]
    mp = box_mean(p, r)
    cov = [box_mean(I[..., c] * p, r) - mI[c] * mp for c in range(3)]
    var = {}
    for i in range(3):
        for j in range(i, 3):
            var[(i, j)] = box_mean(I[..., i] * I[..., j], r) - mI[i] * mI[j]
    s00, s01, s02 = var[(0, 0)] + eps, var[(0, 1)], var[(0, 2)]
    s11, s12, s22 = var[(1, 1)] + eps, var[(1, 2)], var[(2, 2)] + eps
    # closed-form inverse of the symmetric 3x3
    c00 = s11 * s22 - s12 * s12
    c01 = s02 * s12 - s01 * s22
    c02 = s01 * s12 - s02 * s11
    c11 = s00 * s22 - s02 * s02
    c12 = s01 * s02 - s00 * s12
    c22 = s00 * s11 - s01 * s01
    det = s00 * c00 + s01 * c01 + s02 * c02
    a0 = (c00 * cov[0] + c01 * cov[1] + c02 * cov[2]) / det
    a1 = (c01 * cov[0] + c11 * cov[1] + c12 * cov[2]) / det
    a2 = (c02 * cov[0] + c12 * cov[1] + c22 * cov[2]) / det
    b = mp - a0 * mI[0] - a1 * mI[1] - a2 * mI[2]
    return np.stack([a0, a1, a2], axis=-1), b


def guided_upscale(depth_u16, guide_rgb, r=8, eps=1e-3):
    """Returns (q float64 HxW, out uint16 HxW)."""
    guide_rgb = np.asarray(guide_rgb)
    H, W = guide_rgb.shape[:2]
    p = bilinear_upsample(np.asarray(depth_u16, np.float64) / 65535.0, H, W)
    I = guide_rgb.astype(np.float64) / 255.0
    a, b = guided_coefficients(p, I, r, eps)
    q = sum(box_mean(a[..., c], r) * I[..., c] for c in range(3)) + box_mean(b, r)
    out = np.floor(np.clip(q, 0.0, 1.0) * 65535.0 + 0.5).astype(np.uint16)
    return q, out


def guided_upscale_cv2(depth_u16, guide_rgb, r=8, eps=1e-3):
    """Fast CPU port of the same filter (fp32, cv2.resize + cv2.boxFilter).  Used ONLY as the timed
    CPU baseline in bench.py; tests check it against guided_upscale() above."""
    import cv2
    guide_rgb = np.asarray(guide_rgb)
    H, W = guide_rgb.shape[:2]
    k = (2 * r + 1, 2 * r + 1)

    def box(a):
        return cv2.boxFilter(a, cv2.CV_32F, k, borderType=cv2.BORDER_REFLECT)

    p = cv2.resize(np.asarray(depth_u16, np.float32) * np.float32(1.0 / 65535.0), (W, H), interpolation=cv2.INTER_LINEAR)
    I = guide_rgb.astype(np.float32) * np.float32(1.0 / 255.0)
    # centre the guide and the input about their global means (box sums are shift covariant)
    I = I - I.reshape(-1, 3).mean(axis=0).astype(np.float32)
    p0 = np.float32(p.mean())
    p = p - p0
    Ic = [np.ascontiguousarray(I[..., c]) for c in range(3)]
    mI = [box(c) for c in Ic]
    mp = box(p)
    cov = [box(Ic[c] * p) - mI[c] * mp for c in range(3)]
    var = {(i, j): box(Ic[i] * Ic[j]) - mI[i] * mI[j] for i in range(3) for j in range(i, 3)}
    s00, s01, s02 = var[(0, 0)] + eps, var[(0, 1)], var[(0, 2)]
    s11, s12, s22 = var[(1, 1)] + eps, var[(1, 2)], var[(2, 2)] + eps
    c00 = s11 * s22 - s12 * s12
    c01 = s02 * s12 - s01 * s22
    c02 = s01 * s12 - s02 * s11
    c11 = s00 * s22 - s02 * s02
    c12 = s01 * s02 - s00 * s12
    c22 = s00 * s11 - s01 * s01
    det = s00 * c00 + s01 * c01 + s02 * c02
    a0 = (c00 * cov[0] + c01 * cov[1] + c02 * cov[2]) / det
    a1 = (c01 * cov[0] + c11 * cov[1] + c12 * cov[2]) / det
    a2 = (c02 * cov[0] + c12 * cov[1] + c22 * cov[2]) / det
    b = mp - a0 * mI[0] - a1 * mI[1] - a2 * mI[2]
    q = box(a0) * Ic[0] + box(a1) * Ic[1] + box(a2) * Ic[2] + box(b) + p0
    return q, np.floor(np.clip(q, 0.0, 1.0) * 65535.0 + 0.5).astype(np.uint16)
