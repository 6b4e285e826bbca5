"""Wall-clock of the reference-facing upscale step (depth PNGs + 4K guide video -> 16-bit 4K PNGs), SURVEY 8(f) items 2-3.

    python tools/bench_upscale_path.py [frames] [png compression; 0 = GPU-side stored PNG] [batch]

Writes `frames` synthetic 1080p uint16 depth PNGs and a 3840x2160 MJPG guide clip, then runs
SimpleDepthUpscaler.process_depth_upscaling on them (guided filter r=8, eps=1e-3) and reports frames/s.
"""
import sys, time, tempfile, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "video-3d-pipeline_b200"))
import cv2, numpy as np


def main(n=48, png_compression=1, batch=4):
    from video_3d_pipeline import synthetic
    from video_3d_pipeline.upscale import SimpleDepthUpscaler
    tmp = Path(tempfile.mkdtemp(prefix="v3d_up_"))
    ddir = tmp / "w" / "depth_x"
    ddir.mkdir(parents=True)
    base_d = [synthetic.depth_u16(5, t, 1920, 1080) for t in range(2)]
    base_g = [np.ascontiguousarray(synthetic.guide_frame(5, t, 3840, 2160)[..., ::-1]) for t in range(2)]
    for t in range(n):
        cv2.imwrite(str(ddir / f"depth_{t:06d}.png"), base_d[t % 2], [cv2.IMWRITE_PNG_COMPRESSION, 1])
    clip = tmp / "g4k.avi"
    vw = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (3840, 2160))
    assert vw.isOpened()
    for t in range(n):
        vw.write(base_g[t % 2])
    vw.release()
    up = SimpleDepthUpscaler(png_compression=png_compression, batch_size=batch, preview=False)
    up.upscale_depth_maps_ffmpeg(str(ddir), 3840, 2160, str(tmp / "warm.mp4"), guide_video=str(clip))     # warm-up
    t0 = time.perf_counter()
    up.upscale_depth_maps_ffmpeg(str(ddir), 3840, 2160, str(tmp / "out.mp4"), guide_video=str(clip))
    wall = time.perf_counter() - t0
    files = len(list((tmp / "out_png16").glob("*.png")))
    print(json.dumps({"frames": n, "png_compression": png_compression, "batch": batch,
                      "upscale_path_fps": round(n / wall, 1), "out_files": files}))


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
