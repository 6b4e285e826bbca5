"""The reference-facing module API (video_3d_pipeline.depth / .upscale) on a GPU box, read like
the tests the reference never had: same calls, same shapes, same values as the cv2 path."""
import cv2
import numpy as np
import pytest

from oracle import cv2_chain
from video_3d_pipeline import synthetic

pytestmark = pytest.mark.gpu

W, H, D = 320, 180, 64


def _write_clip(path, frames, fps=24.0):
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), fps, (w, h))
    if not vw.isOpened():
        pytest.skip("this OpenCV build cannot write MJPG/AVI")
    for f in frames:
        vw.write(f)
    vw.release()


def test_extractor_surface_and_frame_batch(tmp_path):
    from video_3d_pipeline.depth import HybridStereoDepthExtractor, IGEVStereoDepthExtractor
    assert IGEVStereoDepthExtractor is HybridStereoDepthExtractor
    ex = IGEVStereoDepthExtractor(work_dir=str(tmp_path / "w"), cache_dir=str(tmp_path / "w"), unsqueeze_sbs=False,
                                  batch_size=2, stereo_only=True)
    for attr in ("device", "work_dir", "cache_dir", "batch_size", "model_checkpoint", "use_neural_guidance",
                 "stereo_only", "unsqueeze_sbs", "model", "model_loaded", "max_vram_usage", "memory_stats"):
        assert hasattr(ex, attr)
    frames = [synthetic.sbs_frame(21, t, W, H, D) for t in range(3)]
    pairs = [ex.split_sbs_frame(f, unsqueeze=False) for f in frames]
    maps = ex.process_frame_batch(pairs)
    m = cv2_chain.make_matcher(D, 0)
    assert len(maps) == 3
    for f, got in zip(frames, maps):
        ref = cv2_chain.depth_from_sbs(f, m, False)
        assert got.dtype == np.float32 and np.array_equal(got, ref)
    # split with unsqueeze == cv2.resize(INTER_LANCZOS4)
    l, r = ex.split_sbs_frame(frames[0], unsqueeze=True)
    cl, cr = cv2_chain.split_sbs_frame(frames[0], True)
    assert np.array_equal(l, cl) and np.array_equal(r, cr)
    with pytest.raises(ValueError):
        ex.split_sbs_frame(np.zeros((4, 7, 3), np.uint8))
    # save_depth_map == reference normalisation + PNG16
    ex.save_depth_map(maps[0], tmp_path / "d.png")
    back = cv2.imread(str(tmp_path / "d.png"), cv2.IMREAD_UNCHANGED)
    assert back.dtype == np.uint16 and np.array_equal(back, cv2_chain.normalize_u16(maps[0]))


def test_process_video_sbs_and_upscale(tmp_path):
    from video_3d_pipeline.depth import IGEVStereoDepthExtractor
    from video_3d_pipeline.upscale import SimpleDepthUpscaler
    frames = [synthetic.sbs_frame(22, t, W, H, D) for t in range(5)]
    clip = tmp_path / "sbs.avi"
    _write_clip(clip, frames)
    cap = cv2.VideoCapture(str(clip))
    decoded = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        decoded.append(f)
    cap.release()
    assert len(decoded) == 5
    ex = IGEVStereoDepthExtractor(work_dir=str(tmp_path / "w"), cache_dir=str(tmp_path / "w"), unsqueeze_sbs=True,
                                  batch_size=2)
    out_dir = ex.process_video_sbs(str(clip), start_frame=1, max_frames=3)
    files = sorted(p.name for p in out_dir.glob("depth_*.png"))
    assert files == [f"depth_{i:06d}.png" for i in range(3)]
    m = cv2_chain.make_matcher(D, 0)
    for i in range(3):
        ref = cv2_chain.normalize_u16(cv2_chain.depth_from_sbs(decoded[1 + i], m, True))
        got = cv2.imread(str(out_dir / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(got, ref)
    # second call hits the cache (depth.py:435-437)
    assert ex.process_video_sbs(str(clip), start_frame=1, max_frames=3) == out_dir
    # GPU-side PNG writer (png_compression=0): different bytes on disk, identical pixels
    ex0 = IGEVStereoDepthExtractor(work_dir=str(tmp_path / "w0"), cache_dir=str(tmp_path / "w0"), unsqueeze_sbs=True,
                                   batch_size=2, png_compression=0)
    out0 = ex0.process_video_sbs(str(clip), start_frame=1, max_frames=3)
    for i in range(3):
        a = cv2.imread(str(out_dir / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(str(out0 / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        assert b.dtype == np.uint16 and np.array_equal(a, b)

    guide_frames = [synthetic.guide_frame(22, t, 4 * W, 2 * H)[..., ::-1].copy() for t in range(3)]
    gclip = tmp_path / "g4k.avi"
    _write_clip(gclip, guide_frames)
    up = SimpleDepthUpscaler(use_nvenc=True)
    res = up.process_depth_upscaling(str(out_dir), str(gclip), output_path=str(tmp_path / "depth_4k_final.mp4"))
    assert isinstance(res, str)
    pngs = sorted((tmp_path / "depth_4k_final_png16").glob("*.png"))
    assert len(pngs) == 3
    img = cv2.imread(str(pngs[0]), cv2.IMREAD_UNCHANGED)
    assert img.dtype == np.uint16 and img.shape == (2 * H, 4 * W)
    # GPU-side PNG writer for the 4K frames: identical pixels
    up0 = SimpleDepthUpscaler(use_nvenc=True, png_compression=0, preview=False)
    up0.process_depth_upscaling(str(out_dir), str(gclip), output_path=str(tmp_path / "depth_4k_fast.mp4"))
    fast = sorted((tmp_path / "depth_4k_fast_png16").glob("*.png"))
    assert len(fast) == 3
    for a, b in zip(pngs, fast):
        assert np.array_equal(cv2.imread(str(a), cv2.IMREAD_UNCHANGED), cv2.imread(str(b), cv2.IMREAD_UNCHANGED))


def test_integer_throughput_probe_is_plausible():
    """v3d_probe_int_throughput: the measured denominator of bench.py's ALU roofline."""
    import torch
    from video_3d_pipeline import _native as nv
    r = nv.probe_int_throughput(0)
    assert set(r) == {"int32_add_min", "u16x2_add_then_min", "u16x2_fused_add_min"}
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for name, (instr, ops) in r.items():
        # between 8 and 128 lanes per SM and clock at 1-2.2 GHz
        assert 8 * sms * 1.0e9 < instr < 128 * sms * 2.2e9, (name, instr)
        assert ops >= instr
    # measured on B200: the fused add-min (VIADDMNMX) issues at 64 lanes per SM and clock whether packed or not,
    # the VIADD + VIMNMX.U16x2 pair at ~113 (two pipes) -- so per cell the packed forms tie and both double int32
    assert r["u16x2_fused_add_min"][1] > 0.8 * r["u16x2_add_then_min"][1]
    assert r["u16x2_fused_add_min"][1] > 1.5 * r["int32_add_min"][1]
    assert r["u16x2_add_then_min"][0] > 1.5 * r["u16x2_fused_add_min"][0]
