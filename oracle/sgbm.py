"""ctypes wrapper around oracle/sgbm_oracle.c (test infrastructure only)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle_sgbm.so"
_lib = None

INV = -16


class Params(C.Structure):
    """Mirror of orc_params; defaults are the reference's literals (depth.py:315-325)."""
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]

    def __init__(self, numDisparities=64, mode=0, blockSize=5, P1=8 * 3 * 5 ** 2, P2=32 * 3 * 5 ** 2,
                 disp12MaxDiff=1, preFilterCap=0, uniquenessRatio=10, speckleWindowSize=100,
                 speckleRange=32, minDisparity=0):
        super().__init__(minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff,
                         preFilterCap, uniquenessRatio, speckleWindowSize, speckleRange, mode)


def build(force=False):
    """Compile the C oracle next to its source (gcc only)."""
    src = _HERE / "sgbm_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math",
                               "-shared", "-o", str(_LIB_PATH), str(src)])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _check(rc, what):
    if rc == -1:
        raise ValueError(f"{what}: invalid argument (cv2 would raise here)")
    if rc:
        raise MemoryError(f"{what}: rc={rc}")


def bgr_to_gray(bgr):
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    out = np.empty(bgr.shape[:-1], np.uint8)
    lib().orc_bgr_to_gray(_p(bgr, C.c_uint8), C.c_long(out.size), _p(out, C.c_uint8))
    return out


def unsqueeze_x2(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((h, 2 * w) + img.shape[2:], np.uint8)
    lib().orc_unsqueeze_x2(_p(img, C.c_uint8), h, w, ch, C.c_long(w * ch), _p(out, C.c_uint8), C.c_long(2 * w * ch))
    return out


def split_gray(sbs_bgr, unsqueeze):
    sbs_bgr = np.ascontiguousarray(sbs_bgr, dtype=np.uint8)
    H, Ws = sbs_bgr.shape[:2]
    We = Ws if unsqueeze else Ws // 2
    left = np.empty((H, We), np.uint8)
    right = np.empty((H, We), np.uint8)
    _check(lib().orc_split_gray(_p(sbs_bgr, C.c_uint8), H, Ws, int(bool(unsqueeze)),
                                _p(left, C.c_uint8), _p(right, C.c_uint8)), "split_gray")
    return left, right


def prefilter(img, preFilterCap=0):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    sob = np.empty((H, W), np.uint8)
    inten = np.empty((H, W), np.uint8)
    lib().orc_prefilter(_p(img, C.c_uint8), W, H, preFilterCap, _p(sob, C.c_uint8), _p(inten, C.c_uint8))
    return sob, inten


def cost_volume(left, right, params):
    left = np.ascontiguousarray(left, dtype=np.uint8)
    right = np.ascontiguousarray(right, dtype=np.uint8)
    H, W = left.shape
    D = params.numDisparities
    if W - D <= 0:
        raise ValueError("cost_volume: invalid argument (cv2 would raise here)")
    Cb = np.empty((H, W - D, D), np.uint16)
    _check(lib().orc_cost_volume(_p(left, C.c_uint8), _p(right, C.c_uint8), W, H, C.byref(params),
                                 _p(Cb, C.c_uint16)), "cost_volume")
    return Cb


def aggregate(Cb, params, unsaturated=False):
    Cb = np.ascontiguousarray(Cb, dtype=np.uint16)
    H, W1, D = Cb.shape
    S = np.empty_like(Cb)
    Su = np.empty_like(Cb) if unsaturated else None
    _check(lib().orc_aggregate(_p(Cb, C.c_uint16), W1, H, C.byref(params), _p(S, C.c_uint16),
                               _p(Su, C.c_uint16)), "aggregate")
    return (S, Su) if unsaturated else S


def aggregate_one(Cb, params, dir_index):
    Cb = np.ascontiguousarray(Cb, dtype=np.uint16)
    H, W1, D = Cb.shape
    L = np.empty_like(Cb)
    _check(lib().orc_aggregate_one(_p(Cb, C.c_uint16), W1, H, C.byref(params), dir_index,
                                   _p(L, C.c_uint16)), "aggregate_one")
    return L


def select(S, W, params):
    S = np.ascontiguousarray(S, dtype=np.uint16)
    H = S.shape[0]
    disp = np.empty((H, W), np.int16)
    _check(lib().orc_select(_p(S, C.c_uint16), W, H, C.byref(params), _p(disp, C.c_int16)), "select")
    return disp


def median3(disp):
    disp = np.ascontiguousarray(disp, dtype=np.int16)
    H, W = disp.shape
    out = np.empty_like(disp)
    lib().orc_median3(_p(disp, C.c_int16), W, H, _p(out, C.c_int16))
    return out


def filter_speckles(disp, newVal, maxSize, maxDiff):
    out = np.array(disp, dtype=np.int16, order="C", copy=True)
    H, W = out.shape
    _check(lib().orc_filter_speckles(_p(out, C.c_int16), W, H, newVal, maxSize, maxDiff), "filter_speckles")
    return out


def sgbm_compute(left, right, params, taps=False):
    """cv2.StereoSGBM.compute restated.  With taps=True also returns the
    intermediate volumes/maps that cv2 does not expose."""
    left = np.ascontiguousarray(left, dtype=np.uint8)
    right = np.ascontiguousarray(right, dtype=np.uint8)
    H, W = left.shape
    D = params.numDisparities
    if W - D <= params.blockSize // 2:
        raise ValueError("sgbm_compute: width - numDisparities must exceed blockSize/2 (cv2.error)")
    disp = np.empty((H, W), np.int16)
    tC = tS = tR = tM = None
    if taps:
        tC = np.empty((H, W - D, D), np.uint16)
        tS = np.empty((H, W - D, D), np.uint16)
        tR = np.empty((H, W), np.int16)
        tM = np.empty((H, W), np.int16)
    _check(lib().orc_sgbm_compute(_p(left, C.c_uint8), _p(right, C.c_uint8), W, H, C.byref(params),
                                  _p(disp, C.c_int16), _p(tC, C.c_uint16), _p(tS, C.c_uint16),
                                  _p(tR, C.c_int16), _p(tM, C.c_int16)), "sgbm_compute")
    if taps:
        return disp, dict(C=tC, S=tS, raw=tR, median=tM)
    return disp


def disp_to_float(disp):
    disp = np.ascontiguousarray(disp, dtype=np.int16)
    out = np.empty(disp.shape, np.float32)
    lib().orc_disp_to_float(_p(disp, C.c_int16), C.c_long(disp.size), _p(out, C.c_float))
    return out


def normalize_u16(depth):
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    out = np.empty(depth.shape, np.uint16)
    lib().orc_normalize_u16(_p(depth, C.c_float), C.c_long(depth.size), _p(out, C.c_uint16))
    return out
