"""Wall-clock of the reference-facing module path (decode -> GPU -> PNG16), SURVEY 8(f) item 1.

    python tools/bench_module_path.py [frames] [decode threads] [png compression; 0 = GPU-side stored PNG] [gpu lanes]

Writes a synthetic full-SBS 3840x1080 MJPG clip, runs IGEVStereoDepthExtractor.process_video_sbs on it
(unsqueeze off, D=128) and reports frames/s with the time split into decode, GPU call and PNG encode.
"""
import sys, time, tempfile, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "video-3d-pipeline_b200"))
import cv2, numpy as np

def main(n=64, decode_threads=4, png_compression=1, gpu_lanes=2):
    from video_3d_pipeline import synthetic
    from video_3d_pipeline.depth import IGEVStereoDepthExtractor
    tmp = Path(tempfile.mkdtemp(prefix="v3d_mod_"))
    clip = tmp / "sbs.avi"
    base = [synthetic.sbs_frame(31, t, 1920, 1080, 128) for t in range(4)]
    vw = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (3840, 1080))
    assert vw.isOpened()
    for t in range(n):
        vw.write(base[t % 4])
    vw.release()
    # decode-only baseline
    t0 = time.perf_counter(); cap = cv2.VideoCapture(str(clip)); k = 0
    while True:
        ok, f = cap.read()
        if not ok: break
        k += 1
    cap.release(); t_dec = time.perf_counter() - t0
    # PNG16-only baseline (one core)
    img = (np.random.default_rng(0).integers(0, 65535, (1080, 1920))).astype(np.uint16)
    t0 = time.perf_counter()
    for i in range(8): cv2.imwrite(str(tmp / f"p{i}.png"), img)
    t_png = (time.perf_counter() - t0) / 8
    ex = IGEVStereoDepthExtractor(work_dir=str(tmp / "w"), cache_dir=str(tmp / "w"), unsqueeze_sbs=False,
                                  batch_size=16, stereo_only=True, num_disparities=128,
                                  decode_threads=decode_threads, png_compression=png_compression, gpu_lanes=gpu_lanes)
    ex.process_video_sbs(str(clip), max_frames=min(n, 32 * gpu_lanes), force_reprocess=True)   # warm-up: every lane's context
    t0 = time.perf_counter()
    out = ex.process_video_sbs(str(clip), force_reprocess=True)
    wall = time.perf_counter() - t0
    print(json.dumps({"decode_threads": decode_threads, "png_compression": png_compression, "gpu_lanes": gpu_lanes, "frames": k, "module_path_fps": round(k / wall, 1), "decode_only_fps": round(k / t_dec, 1),
                      "png16_encode_ms_per_frame_one_core": round(t_png * 1000, 1), "out_dir_files": len(list(out.glob('*.png')))}))

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 4,
         int(sys.argv[3]) if len(sys.argv) > 3 else 1, int(sys.argv[4]) if len(sys.argv) > 4 else 2)
