"""Generates the committed golden vectors from the LIVE cv2 (the reference's arithmetic engine).

    python tests/golden/make_golden.py

Inputs are stored next to the outputs so the fixtures do not depend on the synthetic generator
being bit-stable across machines.  cv2 version used is recorded in each file.
"""
import sys
from pathlib import Path

import cv2
import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "video-3d-pipeline_b200"))

from oracle import cv2_chain, guided          # noqa: E402
from video_3d_pipeline import synthetic       # noqa: E402


def main():
    ver = np.array(cv2.__version__)
    # 1) SGBM: the reference's parameters (depth.py:315-325) at D=64 MODE_SGBM, and D=64 MODE_HH
    for name, (W, H, D, mode) in {"sgbm_d64_mode0": (192, 72, 64, 0), "sgbm_d64_mode1": (160, 48, 64, 1),
                                  "sgbm_d128_mode0": (224, 40, 128, 0)}.items():
        left, right, _ = synthetic.stereo_pair(101, 0, W, H, D)
        disp = cv2_chain.make_matcher(D, mode).compute(left, right)
        np.savez_compressed(HERE / f"{name}.npz", left=left, right=right, disp=disp, D=D, mode=mode, cv2_version=ver)
    # 2) the whole per-frame chain from an SBS BGR frame, with and without unsqueeze
    frame = synthetic.sbs_frame(102, 0, 160, 60, 64)
    out = {}
    for uns in (False, True):
        m = cv2_chain.make_matcher(64, 0)
        l, r = cv2_chain.split_sbs_frame(frame, uns)
        depth = cv2_chain.depth_from_sbs(frame, m, uns)
        out[f"gray_left_{int(uns)}"] = cv2_chain.to_gray(l)
        out[f"depth_f32_{int(uns)}"] = depth
        out[f"depth_u16_{int(uns)}"] = cv2_chain.normalize_u16(depth)
    np.savez_compressed(HERE / "chain_sbs.npz", frame=frame, cv2_version=ver, **out)
    # 3) guided upscale: the float64 definition (no reference implementation exists)
    d = synthetic.depth_u16(103, 0, 80, 45)
    g = synthetic.guide_frame(103, 0, 160, 90)
    q, o = guided.guided_upscale(d, g, 8, 1e-3)
    np.savez_compressed(HERE / "guided_r8.npz", depth=d, guide=g, q=q.astype(np.float64), out=o)
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
