"""Checks the frame-range sharded module path (process_video_sbs with num_gpus > 1) against a 1-GPU run."""
import sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "video-3d-pipeline_b200"))
import cv2, numpy as np, torch

def main():
    from video_3d_pipeline import synthetic
    from video_3d_pipeline.depth import IGEVStereoDepthExtractor
    n_gpus = torch.cuda.device_count()
    tmp = Path(tempfile.mkdtemp(prefix="v3d_mg_"))
    clip = tmp / "sbs.avi"
    vw = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (1280, 360))
    for t in range(37):
        vw.write(synthetic.sbs_frame(51, t, 640, 360, 64))
    vw.release()
    outs = {}
    for g in (1, n_gpus):
        ex = IGEVStereoDepthExtractor(work_dir=str(tmp / f"w{g}"), cache_dir=str(tmp / f"w{g}"), unsqueeze_sbs=False,
                                      batch_size=4, stereo_only=True, num_gpus=g)
        t0 = time.time()
        outs[g] = ex.process_video_sbs(str(clip), start_frame=2, max_frames=33)
        print(f"num_gpus={g}: {time.time() - t0:.1f}s -> {len(list(outs[g].glob('*.png')))} files")
    a, b = outs[1], outs[n_gpus]
    for i in range(33):
        x = cv2.imread(str(a / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        y = cv2.imread(str(b / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        assert x is not None and y is not None and np.array_equal(x, y), i
    print(f"MULTIGPU_MODULE_OK gpus={n_gpus}")

if __name__ == "__main__":
    main()
