"""Guided depth upscaling on B200 -- drop-in for the reference's upscale.py.

Same class / method surface as /root/reference/src/video_3d_pipeline/upscale.py.  The reference
lets ffmpeg `scale` the 16-bit depth PNGs to the 4K size and encodes 8-bit H.264
(upscale.py:47-59) without ever reading a 4K pixel.  Here the depth maps are upsampled with the
4K frame itself as guide (colour guided filter, radius 8, eps 1e-3 -- the filter the reference's
readme.md:97,119 promises) in the sm_100a kernels of libv3d.so, and written as a 16-bit PNG
sequence (plus an 8-bit mp4 preview when an encoder is available).
"""
import argparse
import glob
import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import cv2
import numpy as np
import torch

from . import _native
from .utils import get_video_info, guide_start_frame


class SimpleDepthUpscaler:
    """Depth upscaling to the 4K frame size (reference class: upscale.py:12)."""

    def __init__(self, use_nvenc: bool = True, radius: int = 8, eps: float = 1e-3, batch_size: int = 4,
                 gpu_index: int = 0, png_compression: int = 1, png_threads: int = 6, preview: bool = True,
                 decode_threads: int = 4, video16: bool = False):
        self.use_nvenc = use_nvenc          # kept for signature compatibility (upscale.py:15-16)
        self.radius = int(radius)
        self.eps = float(eps)
        self.batch_size = int(batch_size)
        self.gpu_index = int(gpu_index)
        # zlib level of the 16-bit 4K PNGs; 0 = stored blocks packed on the GPU (v3d_png16_pack): ~16.6 MB files,
        # but no libpng / deflate on the host (cv2.imwrite of a 4K uint16 frame costs ~0.2 s of one core)
        self.png_compression = int(png_compression)
        self.png_threads = max(1, int(png_threads))
        self.preview = bool(preview)        # also write the 8-bit mp4 preview at output_path
        # also write ONE 16-bit video file next to output_path: <stem>_16bit.mkv, FFV1 (lossless) gray16le -- the
        # single-file 16-bit form of step 3's output (the reference's H.264, upscale.py:53-59, keeps 8 bits)
        self.video16 = bool(video16)
        self.decode_threads = max(1, int(decode_threads))   # guide-video readers, each on a contiguous slice of the clip
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA not available but requested")
        _native.lib()
        print("Initializing guided depth upscaler (B200 native)...")
        self._ctx = None

    def _context(self, w, h, batch):
        c = self._ctx
        if c is None or c.max_batch < batch:
            if c is not None:
                c.close()
            # the upscale kernels only use the context for its coefficient workspace; the SGBM part
            # is sized minimally (width must exceed numDisparities + blockSize/2)
            self._ctx = c = _native.Context(72, 8, _native.SgbmParams(), max_batch=batch, device=self.gpu_index)
        return c

    def upscale_depth_maps_ffmpeg(self, depth_dir: str, target_width: int, target_height: int, output_path: str,
                                  fps: float = 23.976, guide_video: str = None, guide_start_frame: int = 0):
        """upscale.py:21-73.  depth_%06d.png -> target size.  With `guide_video` (what
        process_depth_upscaling passes) frame i of that video guides depth map i; without it the
        upsampled depth guides itself."""
        print("Processing depth upscaling on GPU...")
        print(f"Input: {depth_dir}")
        print(f"Output: {output_path}")
        print(f"Target: {target_width}x{target_height} @ {fps}fps")
        depth_files = sorted(glob.glob(os.path.join(depth_dir, "depth_*.png")))
        if not depth_files:
            raise ValueError(f"No depth maps found in {depth_dir}")           # upscale.py:37-38
        print(f"Found {len(depth_files)} depth maps")

        out_path = Path(output_path)
        png_dir = out_path.with_suffix("")
        png_dir = png_dir.parent / (png_dir.name + "_png16")
        png_dir.mkdir(parents=True, exist_ok=True)
        writer = None
        if self.preview:
            writer = cv2.VideoWriter(str(out_path), cv2.VideoWriter_fourcc(*"mp4v"), float(fps),
                                     (int(target_width), int(target_height)), False)
            if not writer.isOpened():
                writer = None
        writer16, path16 = None, out_path.with_name(out_path.stem + "_16bit.mkv")
        if self.video16:
            writer16 = cv2.VideoWriter(str(path16), cv2.CAP_FFMPEG, cv2.VideoWriter_fourcc(*"FFV1"), float(fps),
                                       (int(target_width), int(target_height)),
                                       [cv2.VIDEOWRITER_PROP_DEPTH, cv2.CV_16U, cv2.VIDEOWRITER_PROP_IS_COLOR, 0])
            if not writer16.isOpened():
                raise RuntimeError("this OpenCV build cannot write 16-bit FFV1 video (video16=True)")

        if guide_video is not None:
            probe = cv2.VideoCapture(str(guide_video))
            ok = probe.isOpened()
            probe.release()
            if not ok:
                raise ValueError(f"Could not open video file: {guide_video}")
        dev = torch.device("cuda", self.gpu_index)
        import queue
        import threading
        pool = ThreadPoolExecutor(max_workers=self.png_threads)
        pending = []
        pinned = {}
        gpu_png = self.png_compression == 0
        png_args = [cv2.IMWRITE_PNG_COMPRESSION, self.png_compression]
        n_batches = (len(depth_files) + self.batch_size - 1) // self.batch_size
        # the preview video needs frames in order, so it keeps one reader; otherwise the clip is cut into
        # contiguous slices, one reader (own VideoCapture, seeked to its first frame) each
        # ... and only where a seek lands exactly on the requested frame (intra-only codecs); long-GOP guides are
        # read by one sequential reader
        from .depth import HybridStereoDepthExtractor
        exact = guide_video is None or HybridStereoDepthExtractor.seek_is_frame_exact(str(guide_video))
        n_readers = 1 if (writer is not None or writer16 is not None or not exact) else max(1, min(self.decode_threads, n_batches))
        batches: "queue.Queue" = queue.Queue(maxsize=2 * n_readers + 1)
        stop = threading.Event()

        def read_depth(f):
            d = cv2.imread(f, cv2.IMREAD_UNCHANGED)
            if d is None or d.ndim != 2:
                raise ValueError(f"Unreadable depth map: {f}")
            return d.astype(np.uint16) if d.dtype == np.uint16 else (d.astype(np.uint16) * 257)

        def producer(b0: int, b1: int):
            # depth PNGs are read by the pool, the guide video is decoded sequentially: both overlap the GPU work
            cap = None
            try:
                if guide_video is not None:
                    cap = cv2.VideoCapture(str(guide_video))
                    first = int(guide_start_frame) + b0 * self.batch_size   # 4K frame of depth map 0 = audio alignment offset (align.py:65-76)
                    if first > 0:
                        cap.set(cv2.CAP_PROP_POS_FRAMES, first)
                for bi in range(b0, b1):
                    if stop.is_set():
                        return
                    s0 = bi * self.batch_size
                    files = depth_files[s0:s0 + self.batch_size]
                    maps = list(pool.map(read_depth, files))
                    guides = []
                    for d in maps:
                        frame = None
                        if cap is not None:
                            ok, frame = cap.read()
                            frame = frame if ok else None
                        if frame is None:       # self-guided
                            g8 = (cv2.resize(d, (target_width, target_height), interpolation=cv2.INTER_LINEAR) >> 8).astype(np.uint8)
                            frame = np.repeat(g8[..., None], 3, axis=2)
                        elif frame.shape[1] != target_width or frame.shape[0] != target_height:
                            frame = cv2.resize(frame, (target_width, target_height), interpolation=cv2.INTER_AREA)
                        guides.append(frame)                                    # BGR; swapped to RGB on the GPU
                    batches.put((s0, maps, guides))
            except BaseException as e:
                batches.put(e)
            finally:
                if cap is not None:
                    cap.release()
                batches.put(None)

        base, extra = divmod(n_batches, n_readers)
        prods, at = [], 0
        for k in range(n_readers):
            nb = base + (1 if k < extra else 0)
            prods.append(threading.Thread(target=producer, args=(at, at + nb), daemon=True))
            at += nb
        for t in prods:
            t.start()
        live = len(prods)
        try:
            while live:
                item = batches.get()
                if item is None:
                    live -= 1
                    continue
                if isinstance(item, BaseException):
                    raise item
                s0, maps, guides = item
                ctx = self._context(target_width, target_height, len(maps))
                n = len(maps)
                if pinned.get("shape") != (self.batch_size, target_height, target_width):
                    pinned["shape"] = (self.batch_size, target_height, target_width)
                    pinned["g"] = torch.empty((self.batch_size, target_height, target_width, 3), dtype=torch.uint8).pin_memory()
                g_h = pinned["g"][:n]
                g_np = g_h.numpy()
                for i, g in enumerate(guides):
                    np.copyto(g_np[i], g)
                d_t = torch.from_numpy(np.stack(maps).view(np.int16)).to(dev).view(torch.uint16)
                g_t = g_h.to(dev, non_blocking=True).flip(-1).contiguous()       # BGR -> RGB guide
                out_dev = ctx.guided_upscale(d_t, g_t, self.radius, self.eps)
                if gpu_png:
                    pay = ctx.png16_pack(out_dev).cpu().numpy()                  # also orders the reuse of the pinned guide buffer
                    for i in range(n):
                        # the row view keeps `pay` alive until the file is written: no per-frame copy
                        pending.append(pool.submit(_native.write_png16, str(png_dir / f"depth4k_{s0 + i:06d}.png"),
                                                   memoryview(pay[i]), target_width, target_height))
                    prev, out = None, None
                    if writer is not None:     # high bytes for the 8-bit preview (torch has no uint16 shifts)
                        prev = ((out_dev.view(torch.int16).to(torch.int32) & 0xFFFF) >> 8).to(torch.uint8).cpu().numpy()
                else:
                    prev = None
                    out = out_dev.cpu().numpy().view(np.uint16)
                    for i in range(n):
                        pending.append(pool.submit(cv2.imwrite, str(png_dir / f"depth4k_{s0 + i:06d}.png"), out[i], png_args))
                if writer is not None:
                    for i in range(len(maps)):
                        writer.write(prev[i] if prev is not None else (out[i] >> 8).astype(np.uint8))
                if writer16 is not None:
                    full = out if out is not None else out_dev.cpu().numpy().view(np.uint16)
                    for i in range(len(maps)):
                        writer16.write(np.ascontiguousarray(full[i]))
            for f in pending:
                f.result()
        finally:
            stop.set()                                  # after an error the producers stop at their next batch
            while any(t.is_alive() for t in prods):     # unblock producers stuck on a full queue
                try:
                    batches.get_nowait()
                except queue.Empty:
                    pass
                for t in prods:
                    t.join(timeout=0.01)
            pool.shutdown(wait=True)
            if writer is not None:
                writer.release()
            if writer16 is not None:
                writer16.release()
        # what was produced is recorded next to the requested path; the 16-bit PNG sequence is the product, the mp4
        # only an 8-bit preview (and absent when preview=False or this OpenCV build has no encoder)
        import json
        have_video = out_path.exists() and out_path.stat().st_size > 0
        self._sidecar(out_path).write_text(json.dumps({"png16_dir": str(png_dir), "frames": len(depth_files),
                                                       "width": int(target_width), "height": int(target_height),
                                                       "preview_video": str(out_path) if have_video else None,
                                                       "video16": str(path16) if writer16 is not None else None}) + "\n")
        print(f"✓ Depth frames saved: {png_dir}" + (f"  (8-bit preview: {output_path})" if have_video else ""))
        return output_path if have_video else str(png_dir)

    @staticmethod
    def _sidecar(out_path: Path) -> Path:
        return out_path.with_suffix(out_path.suffix + ".json")

    def _already_processed(self, out_path: Path, n_expected: int):
        """The finished result of an earlier run: its sidecar exists and the PNG sequence is complete."""
        import json
        try:
            meta = json.loads(self._sidecar(out_path).read_text())
        except (OSError, ValueError):
            return None
        png_dir = Path(meta.get("png16_dir", ""))
        if meta.get("frames") != n_expected or not png_dir.is_dir():
            return None
        if len(list(png_dir.glob("depth4k_*.png"))) < n_expected:
            return None
        pv = meta.get("preview_video")
        return pv if pv and Path(pv).exists() else str(png_dir)

    def process_depth_upscaling(self, depth_dir: str, video_4k_path: str, output_path: str = None,
                                force_reprocess: bool = False) -> str:
        """upscale.py:75-123."""
        print("Processing depth upscaling...")
        print(f"Depth maps: {depth_dir}")
        print(f"4K video: {video_4k_path}")
        info = get_video_info(video_4k_path)
        if not info:
            raise ValueError(f"Could not read video info: {video_4k_path}")   # upscale.py:88-89
        tw, th, fps = info["width"], info["height"], info["fps"]
        print(f"Target resolution: {tw}x{th} @ {fps}fps")
        if output_path is None:
            output_path = f"depth_4k_{Path(depth_dir).name}.mp4"               # upscale.py:98-100
        output_path = Path(output_path)
        if not force_reprocess:                                               # upscale.py:104-107
            n_maps = len(glob.glob(os.path.join(str(depth_dir), "depth_*.png")))
            done = self._already_processed(output_path, n_maps) if n_maps else None
            if done:
                print(f"✓ Using existing depth video: {done}")
                return done
        # the alignment step leaves alignment_data.json in the work dir, next to the depth cache (run_pipeline.py:53)
        start = 0
        align_json = Path(depth_dir).parent / "alignment_data.json"
        if align_json.exists():
            try:
                start = guide_start_frame(str(align_json), str(video_4k_path), fps or 23.976)
                print(f"Guide frames start at 4K frame {start} (audio alignment)")
            except (ValueError, KeyError) as e:
                print(f"Warning: alignment data not applicable ({e}); guide starts at frame 0")
        result = self.upscale_depth_maps_ffmpeg(depth_dir=str(depth_dir), target_width=tw, target_height=th,
                                                output_path=str(output_path), fps=fps or 23.976,
                                                guide_video=video_4k_path, guide_start_frame=start)
        print("✓ Depth upscaling complete!")
        print(f"  Resolution: {tw}x{th}")
        return result


def main():
    """Command line interface (upscale.py:126-158)."""
    parser = argparse.ArgumentParser(description="Guided depth upscaling on GPU")
    parser.add_argument("depth_dir", help="Directory containing depth maps")
    parser.add_argument("video_4k", help="Path to 4K 2D video (dimensions and guide frames)")
    parser.add_argument("--output", help="Output path for 4K depth video")
    parser.add_argument("--no-nvenc", action="store_true")
    parser.add_argument("--force", action="store_true")
    args = parser.parse_args()
    try:
        upscaler = SimpleDepthUpscaler(use_nvenc=not args.no_nvenc)
        out = upscaler.process_depth_upscaling(depth_dir=args.depth_dir, video_4k_path=args.video_4k,
                                               output_path=args.output, force_reprocess=args.force)
        print(f"\n✓ Success! 4K depth video: {out}")
    except Exception as e:
        print(f"Error: {e}")
        return 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
