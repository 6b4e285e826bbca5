"""Seeded synthetic SBS clips and 4K guides.

No files, no codec: frame t of a clip with seed s is a pure function of (s, t),
so every rank / test / bench leg regenerates exactly the same pixels.
"""
import cv2
import numpy as np


def _stretch_u8(a):
    a = a.astype(np.float32)
    lo, hi = float(a.min()), float(a.max())
    if hi <= lo:
        return np.zeros(a.shape, np.uint8)
    return ((a - lo) * (255.0 / (hi - lo))).astype(np.uint8)


def stereo_pair(seed, t, eye_w, eye_h, num_disp):
    """One textured gray stereo pair with piecewise-constant true disparity.

    Returns (left, right, true_disparity) -- uint8 HxW, uint8 HxW, int32 HxW.
    """
    D, W, H = num_disp, eye_w, eye_h
    rng = np.random.default_rng(seed * 1_000_003 + t)
    tex = rng.integers(0, 256, size=(H, W + D + 8), dtype=np.uint8)
    tex = _stretch_u8(cv2.GaussianBlur(tex.astype(np.float32), (0, 0), 1.2))
    disp = np.full((H, W), max(D // 6, 1), np.int32)
    for _ in range(6):
        rh, rw = max(H // 4, 1), max(W // 4, 1)
        y0 = int(rng.integers(0, max(H - rh, 1)))
        x0 = int(rng.integers(0, max(W - rw, 1)))
        disp[y0:y0 + rh, x0:x0 + rw] = int(rng.integers(1, max(D - 2, 2)))
    right = tex[:, D:D + W]
    cols = np.arange(W)[None, :] - disp + D
    left = np.take_along_axis(tex, cols, axis=1)
    return np.ascontiguousarray(left), np.ascontiguousarray(right), disp


def sbs_frame(seed, t, eye_w, eye_h, num_disp):
    """BGR uint8 side-by-side frame H x 2*eye_w x 3 (left eye | right eye).

    The three channels are the plane and two 1-pixel rolls of it so that the
    BGR->gray conversion is exercised with three different values per pixel.
    """
    left, right, _ = stereo_pair(seed, t, eye_w, eye_h, num_disp)

    def bgr(p):
        return np.stack([p, np.roll(p, 1, axis=1), np.roll(p, 1, axis=0)], axis=-1)

    return np.ascontiguousarray(np.concatenate([bgr(left), bgr(right)], axis=1))


def guide_frame(seed, t, w4k, h4k):
    """RGB uint8 guide frame (h4k x w4k x 3): blurred noise with hard edges."""
    rng = np.random.default_rng(seed * 7_000_003 + t)
    g = rng.integers(0, 256, size=(h4k, w4k, 3), dtype=np.uint8).astype(np.float32)
    g = cv2.GaussianBlur(g, (0, 0), 3.0)
    out = np.empty((h4k, w4k, 3), np.uint8)
    for c in range(3):
        out[..., c] = _stretch_u8(g[..., c])
    # a few rectangles so the filter has real edges to follow
    for _ in range(4):
        rh, rw = max(h4k // 5, 1), max(w4k // 5, 1)
        y0 = int(rng.integers(0, max(h4k - rh, 1)))
        x0 = int(rng.integers(0, max(w4k - rw, 1)))
        out[y0:y0 + rh, x0:x0 + rw] = rng.integers(0, 256, size=3, dtype=np.uint8)
    return out


def depth_u16(seed, t, w, h):
    """Smooth synthetic uint16 depth map (h x w) for upscale-only runs."""
    rng = np.random.default_rng(seed * 9_000_011 + t)
    d = rng.random((h, w), dtype=np.float32)
    d = cv2.GaussianBlur(d, (0, 0), 9.0)
    lo, hi = float(d.min()), float(d.max())
    return ((d - lo) / max(hi - lo, 1e-12) * 65535.0).astype(np.uint16)
