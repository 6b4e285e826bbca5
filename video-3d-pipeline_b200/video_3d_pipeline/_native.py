"""ctypes binding of libv3d.so (include/v3d.h) over PyTorch CUDA tensors.

PyTorch is only the plumbing here (device memory, pinned host memory, streams);
every pixel is computed by the hand-written sm_100a kernels behind the C ABI.
There is NO CPU fallback: if the library is missing or no CUDA device is present
the calls raise, loudly.
"""
import ctypes as C
from pathlib import Path

import torch

import os

# V3D_LIB points at another build of the same library (A/B measurements of kernel variants); default: in-tree
_LIB_PATH = Path(os.environ.get("V3D_LIB") or (Path(__file__).resolve().parent / "libv3d.so"))
_lib = None

V3D_EINVAL, V3D_ENOMEM, V3D_ECUDA, V3D_ESTATE = -1, -2, -3, -4
INVALID_DISP = -16
MODE_SGBM, MODE_HH = 0, 1

STAGES = ("split_gray", "prefilter", "cost", "vertical", "lr", "wta", "select", "median", "speckle", "post", "guided_coeff", "guided_apply", "copy")


class SgbmParams(C.Structure):
    """v3d_sgbm_params; defaults are the reference's literals (depth.py:315-325)."""
    _fields_ = [(n, C.c_int32) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]

    def __init__(self, numDisparities=64, mode=MODE_SGBM, blockSize=5, P1=8 * 3 * 5 ** 2, P2=32 * 3 * 5 ** 2,
                 disp12MaxDiff=1, preFilterCap=0, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32,
                 minDisparity=0):
        super().__init__(minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap,
                         uniquenessRatio, speckleWindowSize, speckleRange, mode)


def lib():
    """Load libv3d.so and declare every prototype of include/v3d.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python video-3d-pipeline_b200/build.py` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    L = C.CDLL(str(_LIB_PATH))
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.v3d_version.restype = C.c_char_p
    L.v3d_last_error.restype = C.c_char_p
    L.v3d_default_params.argtypes = [C.POINTER(SgbmParams)]
    L.v3d_default_params.restype = None
    L.v3d_create.argtypes = [i32, C.POINTER(SgbmParams), i32, i32, i32, C.POINTER(vp)]
    L.v3d_destroy.argtypes = [vp]
    L.v3d_workspace_bytes.argtypes = [vp]
    L.v3d_workspace_bytes.restype = sz
    L.v3d_split_gray.argtypes = [vp, vp, sz, sz, i32, i32, i32, i32, vp, vp, sz, sz, vp]
    L.v3d_bgr_to_gray.argtypes = [vp, vp, sz, sz, i32, i32, i32, vp, sz, sz, vp]
    L.v3d_unsqueeze_bgr.argtypes = [i32, vp, sz, sz, i32, i32, i32, vp, sz, sz, vp]
    L.v3d_sgbm_compute.argtypes = [vp, vp, vp, sz, sz, i32, vp, sz, sz, vp]
    L.v3d_set_debug_taps.argtypes = [vp, i32]
    L.v3d_debug_tap.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(sz)]
    L.v3d_debug_tap_copy.argtypes = [vp, i32, vp, sz, vp]
    L.v3d_postprocess.argtypes = [vp, vp, sz, sz, i32, vp, vp, vp]
    L.v3d_normalize_u16.argtypes = [vp, vp, sz, i32, vp, vp]
    L.v3d_guided_upscale.argtypes = [vp, vp, i32, i32, vp, i32, i32, i32, i32, C.c_float, vp, vp, vp]
    L.v3d_depth_frames.argtypes = [vp, vp, sz, sz, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, i32, C.c_float, vp, vp]
    L.v3d_depth_frames_host.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, i32, C.c_float, vp, vp]
    L.v3d_depth_frames_host_async.argtypes = L.v3d_depth_frames_host.argtypes
    L.v3d_guided_upscale_host_async.argtypes = [vp, vp, vp, i32, i32, i32, i32, C.c_float, vp, vp]
    L.v3d_host_copy_only_async.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, i32, vp, vp]
    L.v3d_host_wait.argtypes = [vp]
    L.v3d_host_wait_oldest.argtypes = [vp]
    L.v3d_host_pending.argtypes = [vp]
    L.v3d_host_pending.restype = C.c_int
    L.v3d_fused_sweep_clusters.argtypes = [vp]
    L.v3d_fused_sweep_clusters.restype = i32
    L.v3d_launch_count.argtypes = [vp]
    L.v3d_launch_count.restype = C.c_ulonglong
    L.v3d_set_depth_scale.argtypes = [vp, i32, C.c_float, C.c_float]
    L.v3d_png16_payload_bytes.argtypes = [i32, i32]
    L.v3d_png16_payload_bytes.restype = sz
    L.v3d_png16_pack.argtypes = [vp, vp, i32, i32, i32, vp, sz, vp]
    L.v3d_set_timing.argtypes = [vp, i32]
    L.v3d_reset_timing.argtypes = [vp]
    L.v3d_stage_ms.argtypes = [vp, i32, C.POINTER(C.c_char_p)]
    L.v3d_stage_ms.restype = C.c_double
    L.v3d_probe_int_throughput.argtypes = [i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.v3d_probe_int_throughput.restype = i32
    for name in ("v3d_create", "v3d_destroy", "v3d_split_gray", "v3d_bgr_to_gray", "v3d_unsqueeze_bgr",
                 "v3d_sgbm_compute", "v3d_set_debug_taps", "v3d_debug_tap", "v3d_debug_tap_copy", "v3d_postprocess", "v3d_normalize_u16",
                 "v3d_guided_upscale", "v3d_depth_frames", "v3d_depth_frames_host", "v3d_set_depth_scale", "v3d_png16_pack", "v3d_set_timing",
                 "v3d_reset_timing", "v3d_depth_frames_host_async", "v3d_guided_upscale_host_async", "v3d_host_copy_only_async",
                 "v3d_host_wait", "v3d_host_wait_oldest"):
        getattr(L, name).restype = i32
    _lib = L
    return L


def probe_int_throughput(device=0):
    """Measured integer add/min rates of `device` (v3d_probe_int_throughput): {mix: (lane-instructions/s,
    algorithmic cell operations/s)} for the three instruction mixes of the path recurrence."""
    out = {}
    for kind, name in enumerate(("int32_add_min", "u16x2_add_then_min", "u16x2_fused_add_min")):
        a, b = C.c_double(), C.c_double()
        _check(lib().v3d_probe_int_throughput(device, kind, C.byref(a), C.byref(b)), "v3d_probe_int_throughput")
        out[name] = (a.value, b.value)
    return out


def _raise(rc, what):
    msg = lib().v3d_last_error().decode(errors="replace")
    text = f"{what}: {msg}"
    if rc == V3D_EINVAL:
        raise ValueError(text)
    if rc == V3D_ENOMEM:
        raise MemoryError(text)
    raise RuntimeError(text)


def _check(rc, what):
    if rc != 0:
        _raise(rc, what)


_PNG_SIG = b"\x89PNG\r\n\x1a\n"


def _png_chunk(tag: bytes, body) -> list:
    import struct
    import zlib
    crc = zlib.crc32(body, zlib.crc32(tag)) & 0xFFFFFFFF
    return [struct.pack(">I", len(body)), tag, body, struct.pack(">I", crc)]


def png16_file_chunks(payload, w: int, h: int) -> list:
    """The byte strings of a complete 16-bit grayscale PNG whose IDAT content is `payload` (a zlib stream of
    the filter-0 scanlines, e.g. one row of Context.png16_pack brought to the host).  The only per-pixel work
    left on the CPU is the CRC-32 of the IDAT chunk (zlib.crc32 releases the GIL)."""
    import struct
    ihdr = struct.pack(">IIBBBBB", int(w), int(h), 16, 0, 0, 0, 0)      # 16 bit, gray, deflate, adaptive, no interlace
    return [_PNG_SIG] + _png_chunk(b"IHDR", ihdr) + _png_chunk(b"IDAT", payload) + _png_chunk(b"IEND", b"")


def write_png16(path, payload, w: int, h: int):
    with open(path, "wb") as f:
        f.writelines(png16_file_chunks(payload, w, h))


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA not available but requested")   # message of depth.py:44
    lib()


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_u8(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA tensor")
    return t


def unsqueeze_bgr(bgr):
    """split_sbs_frame's Lanczos x2 on a CUDA BGR batch [B,H,W,3] -> [B,H,2W,3] (depth.py:263-266)."""
    _dev_u8(bgr, "bgr")
    B, H, W, _ = bgr.shape
    out = torch.empty((B, H, 2 * W, 3), dtype=torch.uint8, device=bgr.device)
    dev = bgr.device.index or 0
    _check(lib().v3d_unsqueeze_bgr(dev, bgr.data_ptr(), W * 3, H * W * 3, W, H, B, out.data_ptr(),
                                   W * 6, H * W * 6, _stream(bgr.device)), "v3d_unsqueeze_bgr")
    return out


class Context:
    """One v3d_ctx: fixed eye size, SGBM parameters and maximum batch."""

    def __init__(self, eye_w, eye_h, params=None, max_batch=1, device=0):
        require_cuda()
        self.params = params or SgbmParams()
        self.W, self.H, self.D = int(eye_w), int(eye_h), int(self.params.numDisparities)
        self.Dk = 64 if self.D <= 64 else (128 if self.D <= 128 else 256)    # kernel disparity count (padded)
        self.max_batch = int(max_batch)
        self.device = torch.device("cuda", int(device))
        self._h = C.c_void_p()
        _check(lib().v3d_create(int(device), C.byref(self.params), self.W, self.H, self.max_batch,
                                C.byref(self._h)), "v3d_create")

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().v3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def workspace_bytes(self):
        return int(lib().v3d_workspace_bytes(self._h))

    @property
    def launch_count(self):
        return int(lib().v3d_launch_count(self._h))

    @property
    def fused_sweep_clusters(self):
        return int(lib().v3d_fused_sweep_clusters(self._h))

    def png16_pack(self, img_u16, out=None):
        """uint16 [B,h,w] CUDA tensor -> uint8 [B,P] IDAT payloads (stored-deflate zlib streams, Adler-32
        included); png16_file_chunks() turns one payload into a .png file."""
        if not (img_u16.is_cuda and img_u16.is_contiguous() and img_u16.dtype == torch.uint16 and img_u16.dim() == 3):
            raise ValueError("img_u16 must be a contiguous uint16 CUDA tensor [B, h, w]")
        B, h, w = img_u16.shape
        P = int(lib().v3d_png16_payload_bytes(w, h))
        if out is None:
            out = torch.empty((B, P), dtype=torch.uint8, device=img_u16.device)
        _check(lib().v3d_png16_pack(self._h, img_u16.data_ptr(), w, h, B, out.data_ptr(), out.stride(0),
                                    _stream(img_u16.device)), "v3d_png16_pack")
        return out

    def set_depth_scale(self, fixed, lo=0.0, hi=1.0):
        """Opt-in clip-level uint16 scale: u16 = trunc(clip((d - lo) / (hi - lo), 0, 1) * 65535).
        fixed=False restores the reference's per-frame min-max (depth.py:400-401)."""
        _check(lib().v3d_set_depth_scale(self._h, int(bool(fixed)), C.c_float(lo), C.c_float(hi)), "v3d_set_depth_scale")

    def set_timing(self, on):
        _check(lib().v3d_set_timing(self._h, int(bool(on))), "v3d_set_timing")

    def reset_timing(self):
        _check(lib().v3d_reset_timing(self._h), "v3d_reset_timing")

    def stage_ms(self):
        out = {}
        for i, name in enumerate(STAGES):
            out[name] = float(lib().v3d_stage_ms(self._h, i, None))
        return out

    def set_debug_taps(self, on):
        _check(lib().v3d_set_debug_taps(self._h, int(bool(on))), "v3d_set_debug_taps")

    # -- stages ------------------------------------------------------------
    def split_gray(self, sbs_bgr, unsqueeze):
        """[B,H,Wsbs,3] uint8 CUDA -> (left, right) gray [B,H,We] (depth.py:250-268, 274-275, 337-338)."""
        _dev_u8(sbs_bgr, "sbs_bgr")
        B, H, Ws, _ = sbs_bgr.shape
        We = Ws if unsqueeze else Ws // 2
        left = torch.empty((B, H, We), dtype=torch.uint8, device=sbs_bgr.device)
        right = torch.empty_like(left)
        _check(lib().v3d_split_gray(self._h, sbs_bgr.data_ptr(), Ws * 3, H * Ws * 3, Ws, H, B, int(bool(unsqueeze)),
                                    left.data_ptr(), right.data_ptr(), We, H * We, _stream(sbs_bgr.device)),
               "v3d_split_gray")
        return left, right

    def bgr_to_gray(self, bgr):
        _dev_u8(bgr, "bgr")
        B, H, W, _ = bgr.shape
        gray = torch.empty((B, H, W), dtype=torch.uint8, device=bgr.device)
        _check(lib().v3d_bgr_to_gray(self._h, bgr.data_ptr(), W * 3, H * W * 3, W, H, B, gray.data_ptr(), W, H * W,
                                     _stream(bgr.device)), "v3d_bgr_to_gray")
        return gray

    def sgbm_compute(self, left_gray, right_gray):
        """stereo.compute (depth.py:341): gray [B,H,W] pairs -> int16 disparity x16 [B,H,W]."""
        _dev_u8(left_gray, "left_gray")
        _dev_u8(right_gray, "right_gray")
        B, H, W = left_gray.shape
        if (W, H) != (self.W, self.H) or right_gray.shape != left_gray.shape:
            raise ValueError(f"expected eyes of {self.W}x{self.H}, got {W}x{H}")
        disp = torch.empty((B, H, W), dtype=torch.int16, device=left_gray.device)
        _check(lib().v3d_sgbm_compute(self._h, left_gray.data_ptr(), right_gray.data_ptr(), W, H * W, B,
                                      disp.data_ptr(), W * 2, H * W * 2, _stream(left_gray.device)),
               "v3d_sgbm_compute")
        return disp

    def debug_tap(self, which, batch):
        """Copy of a workspace volume of the last compute call (parity tests only)."""
        md = int(self.params.minDisparity)
        W1 = (self.W + min(md, 0)) - max(md + self.D, 0)          # OpenCV: maxX1 - minX1
        shape, dt = {0: ((batch, self.H, W1, self.Dk), torch.int16), 1: ((batch, self.H, W1, self.Dk), torch.int16),
                     2: ((batch, self.H, self.W), torch.int16), 3: ((batch, self.H, self.W), torch.int16)}[which]
        out = torch.empty(shape, dtype=dt, device=self.device)
        _check(lib().v3d_debug_tap_copy(self._h, which, out.data_ptr(), out.numel() * out.element_size(),
                                        _stream(self.device)), "v3d_debug_tap_copy")
        return out

    def postprocess(self, disp, want_f32=True, want_u16=True):
        """depth.py:341 (/16), :374 (<=0 -> 0) and :400-403 (min-max -> uint16)."""
        B, H, W = disp.shape
        f32 = torch.empty((B, H, W), dtype=torch.float32, device=disp.device) if want_f32 else None
        u16 = torch.empty((B, H, W), dtype=torch.uint16, device=disp.device) if want_u16 else None
        _check(lib().v3d_postprocess(self._h, disp.data_ptr(), W * 2, H * W * 2, B,
                                     f32.data_ptr() if want_f32 else None, u16.data_ptr() if want_u16 else None,
                                     _stream(disp.device)), "v3d_postprocess")
        return f32, u16

    def normalize_u16(self, depth_f32):
        """save_depth_map's normalisation of float maps [B, ...] (depth.py:400-403)."""
        B = depth_f32.shape[0]
        n = depth_f32[0].numel()
        out = torch.empty(depth_f32.shape, dtype=torch.uint16, device=depth_f32.device)
        _check(lib().v3d_normalize_u16(self._h, depth_f32.data_ptr(), n, B, out.data_ptr(),
                                       _stream(depth_f32.device)), "v3d_normalize_u16")
        return out

    def guided_upscale(self, depth_u16, guide_rgb, r=8, eps=1e-3, want_q=False, out=None):
        """uint16 depth [B,h,w] + RGB guide [B,gh,gw,3] -> uint16 [B,gh,gw] (definition: oracle/guided.py)."""
        B, h, w = depth_u16.shape
        _, gh, gw, _ = guide_rgb.shape
        if out is None:
            out = torch.empty((B, gh, gw), dtype=torch.uint16, device=depth_u16.device)
        q = torch.empty((B, gh, gw), dtype=torch.float32, device=depth_u16.device) if want_q else None
        _check(lib().v3d_guided_upscale(self._h, depth_u16.data_ptr(), w, h, guide_rgb.data_ptr(), gw, gh, B, int(r),
                                        C.c_float(eps), out.data_ptr(), q.data_ptr() if want_q else None,
                                        _stream(depth_u16.device)), "v3d_guided_upscale")
        return (out, q) if want_q else out

    def depth_frames(self, sbs_bgr, unsqueeze, guide_rgb=None, r=8, eps=1e-3, want=("disp", "f32", "u16")):
        """Whole device-resident frame path.  Returns a dict of the requested CUDA tensors."""
        _dev_u8(sbs_bgr, "sbs_bgr")
        B, H, Ws, _ = sbs_bgr.shape
        dev = sbs_bgr.device
        res = {}
        if "disp" in want:
            res["disp"] = torch.empty((B, self.H, self.W), dtype=torch.int16, device=dev)
        if "f32" in want:
            res["f32"] = torch.empty((B, self.H, self.W), dtype=torch.float32, device=dev)
        if "u16" in want:
            res["u16"] = torch.empty((B, self.H, self.W), dtype=torch.uint16, device=dev)
        gw = gh = 0
        if guide_rgb is not None:
            _dev_u8(guide_rgb, "guide_rgb")
            _, gh, gw, _ = guide_rgb.shape
            res["out4k"] = torch.empty((B, gh, gw), dtype=torch.uint16, device=dev)

        def ptr(k):
            return res[k].data_ptr() if k in res else None

        _check(lib().v3d_depth_frames(self._h, sbs_bgr.data_ptr(), Ws * 3, H * Ws * 3, Ws, H, B, int(bool(unsqueeze)),
                                      ptr("disp"), ptr("f32"), ptr("u16"),
                                      guide_rgb.data_ptr() if guide_rgb is not None else None, gw, gh, int(r),
                                      C.c_float(eps), ptr("out4k"), _stream(dev)), "v3d_depth_frames")
        return res

    def depth_frames_host(self, sbs_bgr, unsqueeze, guide_rgb=None, r=8, eps=1e-3, out=None, wait=True):
        """Same path from HOST tensors (pinned recommended); `out` maps 'disp'/'f32'/'u16'/'out4k' to
        preallocated host tensors that receive the results.  wait=False returns as soon as the copies and
        kernels are enqueued (v3d_depth_frames_host_async); call host_wait() before touching the buffers."""
        if sbs_bgr.is_cuda or not sbs_bgr.is_contiguous():
            raise ValueError("sbs_bgr must be a contiguous host tensor")
        B, H, Ws, _ = sbs_bgr.shape
        out = out or {}
        gw = gh = 0
        if guide_rgb is not None:
            if guide_rgb.is_cuda or not guide_rgb.is_contiguous():
                raise ValueError("guide_rgb must be a contiguous host tensor")
            _, gh, gw, _ = guide_rgb.shape

        def ptr(k):
            return out[k].data_ptr() if k in out else None

        fn = lib().v3d_depth_frames_host if wait else lib().v3d_depth_frames_host_async
        with torch.cuda.device(self.device):
            _check(fn(self._h, sbs_bgr.data_ptr(), Ws, H, B, int(bool(unsqueeze)), ptr("disp"), ptr("f32"), ptr("u16"),
                      guide_rgb.data_ptr() if guide_rgb is not None else None, gw, gh,
                      int(r), C.c_float(eps), ptr("out4k"), _stream(self.device)), "v3d_depth_frames_host")
        return out

    def guided_upscale_host(self, depth_u16, guide_rgb, out_u16, r=8, eps=1e-3, wait=True):
        """The upscale step alone from HOST tensors: uint16 depth [B,H,W] (the context's eye size) + RGB guide
        [B,gh,gw,3] -> out_u16 [B,gh,gw] (v3d_guided_upscale_host_async)."""
        B, H, W = depth_u16.shape
        if (W, H) != (self.W, self.H):
            raise ValueError(f"expected depth maps of {self.W}x{self.H}, got {W}x{H}")
        _, gh, gw, _ = guide_rgb.shape
        with torch.cuda.device(self.device):
            _check(lib().v3d_guided_upscale_host_async(self._h, depth_u16.data_ptr(), guide_rgb.data_ptr(), gw, gh, B, int(r),
                                                       C.c_float(eps), out_u16.data_ptr(), _stream(self.device)),
                   "v3d_guided_upscale_host_async")
            if wait:
                self.host_wait()
        return out_u16

    def host_copy_only(self, sbs_bgr=None, guide_rgb=None, out=None):
        """Measurement aid: the host<->device copies of depth_frames_host(wait=False) without any kernel."""
        out = out or {}
        B = (sbs_bgr if sbs_bgr is not None else guide_rgb).shape[0]
        H = Ws = gw = gh = 0
        if sbs_bgr is not None:
            _, H, Ws, _ = sbs_bgr.shape
        if guide_rgb is not None:
            _, gh, gw, _ = guide_rgb.shape
        with torch.cuda.device(self.device):
            _check(lib().v3d_host_copy_only_async(self._h, sbs_bgr.data_ptr() if sbs_bgr is not None else None, Ws, H, B,
                                                  out["disp"].data_ptr() if "disp" in out else None,
                                                  guide_rgb.data_ptr() if guide_rgb is not None else None, gw, gh,
                                                  out["out4k"].data_ptr() if "out4k" in out else None,
                                                  _stream(self.device)), "v3d_host_copy_only_async")

    def host_wait(self):
        """Sleep until every asynchronous host call of this context has delivered its outputs."""
        _check(lib().v3d_host_wait(self._h), "v3d_host_wait")

    def host_wait_oldest(self):
        """Sleep until the OLDEST asynchronous host call in flight has delivered its outputs (two may be in flight:
        submit k+1, then wait for k)."""
        _check(lib().v3d_host_wait_oldest(self._h), "v3d_host_wait_oldest")

    @property
    def host_pending(self) -> int:
        """Asynchronous host calls in flight (0, 1 or 2)."""
        return int(lib().v3d_host_pending(self._h))
