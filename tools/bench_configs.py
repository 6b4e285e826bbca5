"""Device-resident timing of the other BASELINE.json configs (development aid, not the contract bench).

    python tools/bench_configs.py [cfg1|cfg2|cfg5 ...]
"""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "video-3d-pipeline_b200"))
import numpy as np, torch
from video_3d_pipeline import _native as nv, synthetic

CFG = {  # eye_w, eye_h, D, mode, guided, batch  (a trailing True = half-SBS input, Lanczos-unsqueezed on the GPU)
    "cfg1": (960, 1080, 64, 0, False, 30),
    "cfg2": (1920, 1080, 128, 0, False, 15),
    "cfg2g": (1920, 1080, 128, 0, True, 15),
    "cfg5": (1920, 1080, 256, 1, True, 8),
    "cfg5s": (1920, 1080, 256, 1, False, 9),
    "rp_default": (1920, 1080, 64, 0, False, 15, True),   # run_pipeline.py:63-68: 1920x1080 SBS, unsqueeze_sbs=True
}

def run(name, steps=4, warmup=2):
    W, H, D, mode, guided, B = CFG[name][:6]
    unsq = len(CFG[name]) > 6 and CFG[name][6]
    frames = np.stack([synthetic.sbs_frame(3, t % 2, W // 2 if unsq else W, H, D // 2 if unsq else D) for t in range(B)])
    sbs = torch.from_numpy(frames).cuda()
    guide = None
    if guided:
        g = synthetic.guide_frame(3, 0, 2 * W, 2 * H)
        guide = torch.from_numpy(np.stack([g] * B)).cuda()
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=mode), max_batch=B) as ctx:
        for _ in range(warmup):
            ctx.depth_frames(sbs, unsq, guide, want=())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ctx.depth_frames(sbs, unsq, guide, want=())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ctx.set_timing(True); ctx.reset_timing()
        ctx.depth_frames(sbs, unsq, guide, want=())
        torch.cuda.synchronize()
        st = {k: round(v, 2) for k, v in ctx.stage_ms().items() if v > 0.005}
        print(json.dumps({"config": name, "eye": [W, H], "D": D, "mode": mode, "guided": guided, "batch": B,
                          "fps": round(B / ms * 1000, 1), "ms_per_frame": round(ms / B, 3), "stages_ms": st,
                          "workspace_gb": round(ctx.workspace_bytes / 1e9, 2)}), flush=True)

if __name__ == "__main__":
    for n in (sys.argv[1:] or ["cfg1", "cfg2", "cfg5s"]):
        run(n)
