"""The reference-facing module API (video_3d_pipeline.depth / .upscale) on a GPU box, read like
the tests the reference never had: same calls, same shapes, same values as the cv2 path."""
from pathlib import Path

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


def _decode_all(path):
    cap = cv2.VideoCapture(str(path))
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    cap.release()
    return out


def test_upscale_module_pixels_with_alignment_offset(tmp_path):
    """process_depth_upscaling (upscale.py:75-123 surface) on a GPU box, pixel values checked: depth map i is
    guided by DECODED 4K frame start + i, start = the audio-alignment offset stored next to the depth cache
    (utils.py:299-327), BGR -> RGB on the GPU; output within 1 LSB of oracle.guided on exactly those inputs."""
    import json
    from oracle import guided as og
    from video_3d_pipeline.upscale import SimpleDepthUpscaler
    w, h, n, start = 160, 90, 3, 2
    work = tmp_path / "work"
    ddir = work / "depth_abc"
    ddir.mkdir(parents=True)
    maps = [synthetic.depth_u16(31, t, w, h) for t in range(n)]
    for i, d in enumerate(maps):
        assert cv2.imwrite(str(ddir / f"depth_{i:06d}.png"), d)
    gclip = tmp_path / "uhd.avi"
    _write_clip(gclip, [synthetic.guide_frame(31, t, 2 * w, 2 * h)[..., ::-1].copy() for t in range(n + start + 1)], fps=24.0)
    decoded = _decode_all(gclip)                                  # what the upscaler's reader sees (MJPG is lossy)
    assert len(decoded) == n + start + 1
    (work / "alignment_data.json").write_text(json.dumps({"video1_path": "sbs.mkv", "video2_path": str(gclip),
                                                          "time_offset_seconds": start / 24.0}))
    for kw in (dict(), dict(png_compression=0, preview=False)):   # cv2 PNGs + preview / GPU-packed PNGs, several readers
        out = tmp_path / ("o%d.mp4" % len(kw))
        res = SimpleDepthUpscaler(use_nvenc=True, batch_size=2, **kw).process_depth_upscaling(str(ddir), str(gclip), output_path=str(out))
        assert isinstance(res, str) and Path(res).exists()
        pngs = sorted((tmp_path / (out.stem + "_png16")).glob("depth4k_*.png"))
        assert len(pngs) == n
        for i, p in enumerate(pngs):
            got = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)
            assert got.dtype == np.uint16 and got.shape == (2 * h, 2 * w)
            _, want = og.guided_upscale(maps[i], decoded[start + i][..., ::-1], 8, 1e-3)
            assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1, (kw, i)
            _, wrong = og.guided_upscale(maps[i], decoded[i][..., ::-1], 8, 1e-3)     # the offset matters
            assert np.abs(got.astype(np.int64) - wrong.astype(np.int64)).max() > 1
        # a second call finds the finished PNG sequence (upscale.py:104-107), not a placeholder file
        again = SimpleDepthUpscaler(use_nvenc=True, batch_size=2, **kw).process_depth_upscaling(str(ddir), str(gclip), output_path=str(out))
        assert again == res
        if kw:
            assert not out.exists() and Path(res).is_dir()        # preview=False: no fake .mp4
    # single-file 16-bit output: lossless FFV1 gray16, frame for frame the PNG sequence
    out = tmp_path / "o16.mp4"
    SimpleDepthUpscaler(use_nvenc=True, batch_size=2, preview=False, video16=True).process_depth_upscaling(
        str(ddir), str(gclip), output_path=str(out))
    cap = cv2.VideoCapture(str(tmp_path / "o16_16bit.mkv"), cv2.CAP_FFMPEG, [cv2.CAP_PROP_CONVERT_RGB, 0])
    pngs = sorted((tmp_path / "o16_png16").glob("depth4k_*.png"))
    assert len(pngs) == n
    for p in pngs:
        ok, f = cap.read()
        assert ok and f.dtype == np.uint16
        assert np.array_equal(f.reshape(2 * h, 2 * w), cv2.imread(str(p), cv2.IMREAD_UNCHANGED))
    assert not cap.read()[0]
    cap.release()


def _run_pipeline_calls(sbs_video, video_4k, work_dir, max_frames):
    """The calls run_pipeline.py makes with --skip-alignment, restated with its keyword arguments
    (run_pipeline.py:63-68 constructor, :70-75 process_video_sbs, :92-98 upscale)."""
    from video_3d_pipeline.depth import IGEVStereoDepthExtractor          # run_pipeline.py:12
    from video_3d_pipeline.upscale import SimpleDepthUpscaler              # run_pipeline.py:13
    extractor = IGEVStereoDepthExtractor(work_dir=work_dir, cache_dir=work_dir, unsqueeze_sbs=True, batch_size=8)
    depth_dir = extractor.process_video_sbs(video_path=sbs_video, start_frame=0, max_frames=max_frames, force_reprocess=False)
    upscaler = SimpleDepthUpscaler(use_nvenc=True)
    video = upscaler.process_depth_upscaling(depth_dir=str(depth_dir), video_4k_path=video_4k,
                                             output_path=f"{work_dir}/depth_4k_final.mp4", force_reprocess=False)
    return depth_dir, video


def _check_pipeline_outputs(work, sbs_clip, n):
    m = cv2_chain.make_matcher(64, 0)                                     # depth.py:315-325 literals
    decoded = _decode_all(sbs_clip)
    dirs = [d for d in Path(work).glob("depth_*") if d.is_dir() and not d.name.endswith("_png16")]
    assert len(dirs) == 1
    for i in range(n):
        got = cv2.imread(str(dirs[0] / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        ref = cv2_chain.normalize_u16(cv2_chain.depth_from_sbs(decoded[i], m, True))    # unsqueeze_sbs=True
        assert got.dtype == np.uint16 and np.array_equal(got, ref), i
    assert len(list((Path(work) / "depth_4k_final_png16").glob("depth4k_*.png"))) == n


def test_run_pipeline_call_sequence_end_to_end(tmp_path):
    """SURVEY section 2 row 7, "the drop-in test": what run_pipeline.py does with --skip-alignment --max-frames 4
    on an MJPG clip pair, PNGs compared with the reference's cv2 chain on the decoded frames."""
    sbs = tmp_path / "sbs.avi"
    uhd = tmp_path / "uhd.avi"
    _write_clip(sbs, [synthetic.sbs_frame(23, t, W // 2, H, 32) for t in range(6)])      # half-SBS 320x180, unsqueezed to 320/eye
    _write_clip(uhd, [synthetic.guide_frame(23, t, 2 * W, 2 * H)[..., ::-1].copy() for t in range(6)])
    work = tmp_path / "work"
    _run_pipeline_calls(str(sbs), str(uhd), str(work), 4)
    _check_pipeline_outputs(work, sbs, 4)


def test_unmodified_reference_run_pipeline_if_present(tmp_path):
    """INTEGRATION.md recipe A with the reference's own, unmodified run_pipeline.py: only where a checkout of the
    reference exists next to a GPU (it does not travel to the GPU box: the test skips there, and the restated
    call sequence above covers the same calls)."""
    import os
    import subprocess
    import sys
    ref = Path(os.environ.get("V3D_REFERENCE_DIR", "/root/reference"))
    if not (ref / "run_pipeline.py").exists():
        pytest.skip("no reference checkout on this box")
    root = Path(__file__).resolve().parent.parent
    shadow = tmp_path / "shadow"
    (shadow / "src").mkdir(parents=True)
    (shadow / "run_pipeline.py").symlink_to(ref / "run_pipeline.py")                      # unmodified: a symlink
    (shadow / "src" / "video_3d_pipeline").symlink_to(root / "video-3d-pipeline_b200" / "video_3d_pipeline")
    sbs, uhd = tmp_path / "sbs.avi", tmp_path / "uhd.avi"
    _write_clip(sbs, [synthetic.sbs_frame(23, t, W // 2, H, 32) for t in range(6)])
    _write_clip(uhd, [synthetic.guide_frame(23, t, 2 * W, 2 * H)[..., ::-1].copy() for t in range(6)])
    work = tmp_path / "work"
    r = subprocess.run([sys.executable, str(shadow / "run_pipeline.py"), str(sbs), str(uhd), "--work-dir", str(work),
                        "--skip-alignment", "--max-frames", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _check_pipeline_outputs(work, sbs, 4)


def test_long_gop_clip_reads_sequentially_and_matches_single_reader(tmp_path):
    """seeking is not frame-exact on long-GOP streams, so such clips get ONE sequential reader whatever
    decode_threads says; an intra-only clip split over several readers gives the same files as one reader."""
    from video_3d_pipeline.depth import IGEVStereoDepthExtractor
    frames = [synthetic.sbs_frame(24, t, W, H, D) for t in range(20)]
    gop = tmp_path / "gop.mp4"
    vw = cv2.VideoWriter(str(gop), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (2 * W, H))
    clips = []
    if vw.isOpened():
        for f in frames:
            vw.write(f)
        vw.release()
        if gop.exists() and gop.stat().st_size > 0:
            assert not IGEVStereoDepthExtractor.seek_is_frame_exact(str(gop))
            clips.append(gop)
    intra = tmp_path / "intra.avi"
    _write_clip(intra, frames)
    assert IGEVStereoDepthExtractor.seek_is_frame_exact(str(intra))
    clips.append(intra)
    for clip in clips:
        outs = []
        for k, threads in enumerate((1, 4)):
            ex = IGEVStereoDepthExtractor(work_dir=str(tmp_path / f"w{clip.stem}{k}"), cache_dir=str(tmp_path / f"w{clip.stem}{k}"),
                                          unsqueeze_sbs=False, batch_size=2, stereo_only=True, decode_threads=threads)
            outs.append(ex.process_video_sbs(str(clip), start_frame=3, max_frames=15))
        decoded = _decode_all(clip)
        m = cv2_chain.make_matcher(D, 0)
        for i in range(15):
            a = cv2.imread(str(outs[0] / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
            b = cv2.imread(str(outs[1] / f"depth_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
            assert np.array_equal(a, b), (clip.name, i)
            assert np.array_equal(a, cv2_chain.normalize_u16(cv2_chain.depth_from_sbs(decoded[3 + i], m, False))), (clip.name, i)
