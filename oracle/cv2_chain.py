"""The reference's own CPU call chain, calling the same cv2 functions it calls.

This is the reference implementation of the hot path (test infrastructure and
``bench.py --impl reference`` only).  Each step cites the reference line it
repeats; files are under /root/reference/src/video_3d_pipeline/.
"""
import cv2
import numpy as np


def split_sbs_frame(sbs_frame, unsqueeze=True):
    """depth.py:250-268."""
    height, width = sbs_frame.shape[:2]
    if width % 2 != 0:
        raise ValueError("SBS frame width must be even")       # depth.py:254-255
    half = width // 2
    left, right = sbs_frame[:, :half], sbs_frame[:, half:]      # depth.py:258-259
    if unsqueeze:                                               # depth.py:263-266
        left = cv2.resize(left, (half * 2, height), interpolation=cv2.INTER_LANCZOS4)
        right = cv2.resize(right, (half * 2, height), interpolation=cv2.INTER_LANCZOS4)
    return left, right


def to_gray(bgr):
    """depth.py:274-275 (BGR2RGB) then depth.py:337-338 (RGB2GRAY)."""
    return cv2.cvtColor(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB), cv2.COLOR_RGB2GRAY)


def make_matcher(numDisparities=64, mode=0, **kw):
    """depth.py:315-325 with D and mode opened up for BASELINE configs 2 and 5."""
    args = dict(minDisparity=0, numDisparities=numDisparities, blockSize=5, P1=8 * 3 * 5 ** 2,
                P2=32 * 3 * 5 ** 2, disp12MaxDiff=1, uniquenessRatio=10, speckleWindowSize=100,
                speckleRange=32, mode=mode)
    args.update(kw)
    return cv2.StereoSGBM_create(**args)


def disparity_s16(left_gray, right_gray, matcher):
    """depth.py:341 before the float conversion: int16, x16 fixed point, invalid -16."""
    return matcher.compute(left_gray, right_gray)


def depth_from_sbs(sbs_bgr, matcher, unsqueeze):
    """depth.py:257-266, 274-275, 337-341, 374: SBS BGR frame -> float32 disparity."""
    left, right = split_sbs_frame(sbs_bgr, unsqueeze)
    lg, rg = to_gray(left), to_gray(right)
    disparity = matcher.compute(lg, rg).astype(np.float32) / 16.0   # depth.py:341
    disparity[disparity <= 0] = 0                                   # depth.py:374
    return disparity.astype(np.float32)


def normalize_u16(depth_map):
    """depth.py:400-403."""
    if depth_map.max() > depth_map.min():
        return ((depth_map - depth_map.min()) / (depth_map.max() - depth_map.min()) * 65535).astype(np.uint16)
    return np.zeros_like(depth_map, dtype=np.uint16)
