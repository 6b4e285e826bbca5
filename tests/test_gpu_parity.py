"""GPU parity tests proper: the sm_100a kernels, called through the C ABI (libv3d.so via ctypes),
against the oracle -- stage by stage, bit-exact for every integer quantity."""
import cv2
import numpy as np
import pytest
import torch

from oracle import cv2_chain
from oracle import guided as og
from oracle import sgbm as osg
from video_3d_pipeline import _native as nv
from video_3d_pipeline import synthetic

pytestmark = pytest.mark.gpu


def _u16(t):
    return t.cpu().numpy().view(np.uint16)


def _run_stages(W, H, D, mode, B=1, unsq=False, seed=7, **kw):
    src_w = W // 2 if unsq else W
    frames = np.stack([synthetic.sbs_frame(seed, t, src_w, H, D // 2 if unsq else D) for t in range(B)])
    po = osg.Params(numDisparities=D, mode=mode, **kw)
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=mode, **kw), max_batch=B) as ctx:
        ctx.set_debug_taps(True)
        l, r = ctx.split_gray(torch.from_numpy(frames).cuda(), unsq)
        disp = ctx.sgbm_compute(l, r)
        C, S = _u16(ctx.debug_tap(0, B)), _u16(ctx.debug_tap(1, B))
        raw, med = ctx.debug_tap(2, B).cpu().numpy(), ctx.debug_tap(3, B).cpu().numpy()
        f32, u16 = ctx.postprocess(disp)
        l, r, disp, f32, u16 = l.cpu().numpy(), r.cpu().numpy(), disp.cpu().numpy(), f32.cpu().numpy(), _u16(u16)
    for b in range(B):
        ol, orr = osg.split_gray(frames[b], unsq)
        assert np.array_equal(l[b], ol) and np.array_equal(r[b], orr), "gray"
        od, taps = osg.sgbm_compute(ol, orr, po, taps=True)
        _, Su = osg.aggregate(taps["C"], po, unsaturated=True)
        assert np.array_equal(C[b][..., :D], taps["C"]), "cost volume"      # d >= D is kernel padding
        assert np.array_equal(S[b][..., :D], Su), "aggregated S"
        assert np.array_equal(raw[b], taps["raw"]), "raw disparity"
        assert np.array_equal(med[b], taps["median"]), "median"
        assert np.array_equal(disp[b], od), "final disparity vs oracle"
        ref = cv2_chain.make_matcher(D, mode, **kw).compute(ol, orr)
        assert np.array_equal(disp[b], ref), "final disparity vs cv2"
        of = osg.disp_to_float(od)
        assert np.array_equal(f32[b], of), "float depth"
        assert np.array_equal(u16[b], osg.normalize_u16(of)), "uint16 depth"


@pytest.mark.parametrize("W,H,D,mode,B", [
    (200, 120, 64, 0, 1), (200, 120, 64, 1, 1), (331, 77, 128, 0, 2), (400, 50, 256, 1, 1),
    (67, 20, 64, 0, 1), (131, 1, 64, 0, 1), (131, 2, 64, 1, 1), (131, 3, 64, 1, 1), (259, 5, 256, 0, 1),
    (640, 360, 128, 0, 3), (960, 270, 64, 0, 2),
])
def test_sgbm_stages_bit_exact(W, H, D, mode, B):
    _run_stages(W, H, D, mode, B)


@pytest.mark.parametrize("W,H,D,mode", [(120, 30, 16, 0), (150, 24, 32, 1), (200, 40, 48, 0), (260, 30, 96, 0),
                                        (300, 20, 112, 1), (400, 24, 160, 0), (420, 16, 240, 1), (90, 12, 80, 0)])
def test_any_multiple_of_16_disparities(W, H, D, mode):
    """cv2 accepts every positive multiple of 16; the kernels run at 64/128/256 with the rest padded."""
    _run_stages(W, H, D, mode)


def test_unsqueeze_path_bit_exact():
    _run_stages(480, 270, 64, 0, B=2, unsq=True)


@pytest.mark.parametrize("kw", [dict(speckleWindowSize=0), dict(uniquenessRatio=0), dict(uniquenessRatio=25),
                                dict(disp12MaxDiff=3), dict(P1=100, P2=900), dict(blockSize=3), dict(blockSize=7),
                                dict(blockSize=1), dict(preFilterCap=31), dict(speckleWindowSize=400, speckleRange=2)])
def test_parameter_variants(kw):
    _run_stages(230, 60, 64, 0, **kw)


@pytest.mark.parametrize("kind", ["uniform", "binary", "flat"])
def test_degenerate_textures(kind):
    rng = np.random.default_rng(3)
    W, H, D = 150, 40, 64
    if kind == "uniform":
        left, right = rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)
    elif kind == "binary":
        left = (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
        right = np.roll(left, -3, axis=1)
    else:
        left = np.full((H, W), 77, np.uint8)
        right = left.copy()
    for mode in (0, 1):
        with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=mode)) as ctx:
            d = ctx.sgbm_compute(torch.from_numpy(left)[None].cuda(), torch.from_numpy(right)[None].cuda())[0].cpu().numpy()
        assert np.array_equal(d, cv2_chain.make_matcher(D, mode).compute(left, right))


def test_full_size_cfg2_vs_cv2():
    """BASELINE configs[1]: 1920x1080/eye, D=128, against the reference's cv2 call chain."""
    W, H, D = 1920, 1080, 128
    frame = synthetic.sbs_frame(11, 0, W, H, D)
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D)) as ctx:
        res = ctx.depth_frames(torch.from_numpy(frame)[None].cuda(), False, want=("disp", "f32", "u16"))
        disp, f32, u16 = res["disp"][0].cpu().numpy(), res["f32"][0].cpu().numpy(), _u16(res["u16"][0])
    m = cv2_chain.make_matcher(D, 0)
    l, r = cv2_chain.split_sbs_frame(frame, False)
    ref = m.compute(cv2_chain.to_gray(l), cv2_chain.to_gray(r))
    assert np.array_equal(disp, ref)
    depth = cv2_chain.depth_from_sbs(frame, m, False)
    assert np.array_equal(f32, depth)
    assert np.array_equal(u16, cv2_chain.normalize_u16(depth))
    assert (disp != -16).mean() > 0.5          # a real workload, not all-invalid


def test_full_size_properties_cfg5():
    """1920x1080, D=256, MODE_HH (BASELINE configs[4]): size-independent properties."""
    W, H, D = 1920, 1080, 256
    left, right, truth = synthetic.stereo_pair(13, 0, W, H, D)
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=1)) as ctx:
        lt, rt = torch.from_numpy(left)[None].cuda(), torch.from_numpy(right)[None].cuda()
        d1 = ctx.sgbm_compute(lt, rt)[0].cpu().numpy()
        d2 = ctx.sgbm_compute(lt, rt)[0].cpu().numpy()
    assert np.array_equal(d1, d2)                                  # deterministic
    assert (d1[:, :D] == -16).all()                                # columns [0, D) invalid
    valid = d1 != -16
    assert ((d1[valid] >= 0) & (d1[valid] <= 16 * (D - 1) + 8)).all()
    assert valid.mean() > 0.4
    err = np.abs(d1[valid] / 16.0 - truth[valid])
    assert (err <= 1.0).mean() > 0.9                               # recovers the synthetic ground truth


def test_identical_images_give_zero_disparity():
    W, H, D = 300, 64, 64
    left, _, _ = synthetic.stereo_pair(2, 0, W, H, D)
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D, speckleWindowSize=0)) as ctx:
        t = torch.from_numpy(left)[None].cuda()
        d = ctx.sgbm_compute(t, t)[0].cpu().numpy()
    assert np.array_equal(d, cv2_chain.make_matcher(D, 0, speckleWindowSize=0).compute(left, left))
    assert (d[:, D:] <= 8).all()


def test_process_frame_batch_entry_bgr_eyes():
    W, H, D = 256, 96, 64
    frame = synthetic.sbs_frame(9, 0, W, H, D)
    left, right = frame[:, :W], frame[:, W:]
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D)) as ctx:
        lg = ctx.bgr_to_gray(torch.from_numpy(np.ascontiguousarray(left))[None].cuda())
        rg = ctx.bgr_to_gray(torch.from_numpy(np.ascontiguousarray(right))[None].cuda())
        assert np.array_equal(lg[0].cpu().numpy(), cv2_chain.to_gray(np.ascontiguousarray(left)))
        d = ctx.sgbm_compute(lg, rg)[0].cpu().numpy()
    assert np.array_equal(d, cv2_chain.make_matcher(D, 0).compute(cv2_chain.to_gray(np.ascontiguousarray(left)),
                                                                  cv2_chain.to_gray(np.ascontiguousarray(right))))


def test_unsqueeze_bgr_matches_cv2():
    import cv2
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (2, 40, 77, 3), dtype=np.uint8)
    out = nv.unsqueeze_bgr(torch.from_numpy(img).cuda()).cpu().numpy()
    for b in range(2):
        assert np.array_equal(out[b], cv2.resize(img[b], (154, 40), interpolation=cv2.INTER_LANCZOS4))


def test_gray_odd_sizes_and_alignment():
    rng = np.random.default_rng(4)
    for (h, ws) in ((3, 34), (5, 70), (2, 258), (4, 1026)):
        frame = rng.integers(0, 256, (1, h, ws, 3), dtype=np.uint8)
        with nv.Context(max(ws // 2, 80), h, nv.SgbmParams()) as ctx:
            l, r = ctx.split_gray(torch.from_numpy(frame).cuda(), False)
        ol, orr = osg.split_gray(frame[0], False)
        assert np.array_equal(l[0].cpu().numpy(), ol) and np.array_equal(r[0].cpu().numpy(), orr)


def test_normalize_arbitrary_float_maps():
    rng = np.random.default_rng(5)
    maps = np.stack([rng.normal(size=(33, 47)).astype(np.float32) * 40, np.full((33, 47), 2.5, np.float32)])
    with nv.Context(80, 8, nv.SgbmParams(), max_batch=2) as ctx:
        out = _u16(ctx.normalize_u16(torch.from_numpy(maps).cuda()))
    for b in range(2):
        assert np.array_equal(out[b], cv2_chain.normalize_u16(maps[b]))


@pytest.mark.parametrize("w,h,gw,gh,r,B", [(96, 54, 192, 108, 8, 1), (100, 60, 230, 131, 4, 1),
                                           (480, 270, 960, 540, 8, 2), (64, 40, 64, 40, 2, 1)])
def test_guided_upscale_within_half_lsb(w, h, gw, gh, r, B):
    """Tolerance (north_star): 0.5 LSB of the 16-bit output on q; <= 1 LSB after rounding."""
    d = np.stack([synthetic.depth_u16(3, t, w, h) for t in range(B)])
    g = np.stack([synthetic.guide_frame(3, t, gw, gh) for t in range(B)])
    with nv.Context(80, 8, nv.SgbmParams(), max_batch=B) as ctx:
        out, q = ctx.guided_upscale(torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16),
                                    torch.from_numpy(g).cuda(), r, 1e-3, want_q=True)
        out, q = _u16(out), q.cpu().numpy()
    for b in range(B):
        oq, ou = og.guided_upscale(d[b], g[b], r, 1e-3)
        assert np.abs(q[b] - oq).max() < 0.5 / 65535
        assert np.abs(out[b].astype(np.int64) - ou.astype(np.int64)).max() <= 1


def test_guided_upscale_full_4k():
    d = synthetic.depth_u16(4, 0, 1920, 1080)
    g = synthetic.guide_frame(4, 0, 3840, 2160)
    with nv.Context(80, 8, nv.SgbmParams()) as ctx:
        out, q = ctx.guided_upscale(torch.from_numpy(d.view(np.int16))[None].cuda().view(torch.uint16),
                                    torch.from_numpy(g)[None].cuda(), 8, 1e-3, want_q=True)
        out, q = _u16(out)[0], q[0].cpu().numpy()
    oq, ou = og.guided_upscale(d, g, 8, 1e-3)
    assert np.abs(q - oq).max() < 0.5 / 65535
    assert np.abs(out.astype(np.int64) - ou.astype(np.int64)).max() <= 1


@pytest.mark.parametrize("gw,gh", [(960, 540), (958, 541), (452, 300)])
def test_guided_vector_and_scalar_store_paths_agree(gw, gh):
    """Without the float tap the apply kernel packs 8 outputs into one 16-byte store and reads their guide
    bytes as 3 x 8 bytes (taken when gw % 8 == 0 and the pointers are aligned); with the tap, or for other
    widths / a misaligned guide, it stores pixel by pixel.  All of them must give the same uint16 image, and
    interior strips (cp.async-staged guide rows) must match border strips (byte loads)."""
    w, h = 480, 270
    d = synthetic.depth_u16(8, 0, w, h)
    g = synthetic.guide_frame(8, 0, gw, gh)
    dt = torch.from_numpy(d.view(np.int16))[None].cuda().view(torch.uint16)
    with nv.Context(80, 8, nv.SgbmParams()) as ctx:
        fast = _u16(ctx.guided_upscale(dt, torch.from_numpy(g)[None].cuda(), 8, 1e-3))[0]
        slow, q = ctx.guided_upscale(dt, torch.from_numpy(g)[None].cuda(), 8, 1e-3, want_q=True)
        slow = _u16(slow)[0]
        # same guide at an address that is only 1-byte aligned
        buf = torch.empty(g.size + 1, dtype=torch.uint8, device="cuda")
        shifted = buf[1:].view(1, gh, gw, 3)
        shifted.copy_(torch.from_numpy(g)[None])
        mis = _u16(ctx.guided_upscale(dt, shifted, 8, 1e-3))[0]
    assert np.array_equal(fast, slow) and np.array_equal(fast, mis)
    oq, ou = og.guided_upscale(d, g, 8, 1e-3)
    assert np.abs(q[0].cpu().numpy() - oq).max() < 0.5 / 65535
    assert np.abs(fast.astype(np.int64) - ou.astype(np.int64)).max() <= 1


@pytest.mark.parametrize("w,h,B", [(91, 37, 3), (1920, 1080, 2), (32767, 3, 1), (5, 7000, 1)])
def test_gpu_png16_writer_round_trips(w, h, B):
    """v3d_png16_pack: the payload is a valid zlib stream (zlib.decompress checks block headers and the
    Adler-32) of the filter-0 big-endian scanlines, and the assembled file decodes to the input with cv2.
    Shapes include rows that straddle 65535-byte stored-block boundaries mid-sample."""
    import zlib
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 65536, (B, h, w), dtype=np.uint16)
    with nv.Context(80, 8, nv.SgbmParams(), max_batch=B) as ctx:
        pay = ctx.png16_pack(torch.from_numpy(img.view(np.int16)).cuda().view(torch.uint16)).cpu().numpy()
    assert pay.shape[1] == nv.lib().v3d_png16_payload_bytes(w, h)
    for b in range(B):
        raw = b"".join(b"\x00" + img[b, y].astype(">u2").tobytes() for y in range(h))
        assert zlib.decompress(pay[b].tobytes()) == raw
        data = b"".join(nv.png16_file_chunks(pay[b].tobytes(), w, h))
        dec = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
        assert dec.dtype == np.uint16 and np.array_equal(dec, img[b])


def test_fixed_depth_scale_is_opt_in_and_exact():
    """v3d_set_depth_scale(fixed) replaces the per-frame min-max by one clip-level scale; the
    default stays the reference's.  Checked against the same three fp32 operations in numpy."""
    W, H, D, B = 320, 96, 64, 2
    frames = np.stack([synthetic.sbs_frame(9, t, W, H, D) for t in range(B)])
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D), max_batch=B) as ctx:
        ref = ctx.depth_frames(torch.from_numpy(frames).cuda(), False, want=("f32", "u16"))
        ref_u16, f32 = _u16(ref["u16"]), ref["f32"].cpu().numpy()
        ctx.set_depth_scale(True, 0.0, float(D))
        fix = _u16(ctx.depth_frames(torch.from_numpy(frames).cuda(), False, want=("u16",))["u16"])
        fix_n = _u16(ctx.normalize_u16(torch.from_numpy(f32).cuda()))
        with pytest.raises(ValueError):
            ctx.set_depth_scale(True, 4.0, 4.0)
        ctx.set_depth_scale(False)
        again = _u16(ctx.depth_frames(torch.from_numpy(frames).cuda(), False, want=("u16",))["u16"])
    for b in range(B):
        assert np.array_equal(ref_u16[b], cv2_chain.normalize_u16(f32[b]))          # default = reference
    t = (f32 - np.float32(0.0)) / (np.float32(D) - np.float32(0.0))
    want = (np.clip(t, np.float32(0), np.float32(1)) * np.float32(65535.0)).astype(np.uint16)
    assert np.array_equal(fix, want) and np.array_equal(fix_n, want)
    assert np.array_equal(again, ref_u16)
    assert not np.array_equal(fix, ref_u16)


def test_host_entry_point_matches_device_path():
    W, H, D, B = 320, 120, 64, 3
    frames = np.stack([synthetic.sbs_frame(6, t, W, H, D) for t in range(B)])
    guides = np.stack([synthetic.guide_frame(6, t, 2 * W, 2 * H) for t in range(B)])
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D), max_batch=B) as ctx:
        dev = ctx.depth_frames(torch.from_numpy(frames).cuda(), False, torch.from_numpy(guides).cuda())
        host = dict(disp=torch.empty((B, H, W), dtype=torch.int16).pin_memory(),
                    f32=torch.empty((B, H, W), dtype=torch.float32).pin_memory(),
                    u16=torch.empty((B, H, W), dtype=torch.uint16).pin_memory(),
                    out4k=torch.empty((B, 2 * H, 2 * W), dtype=torch.uint16).pin_memory())
        n0 = ctx.launch_count
        ctx.depth_frames_host(torch.from_numpy(frames).pin_memory(), False, torch.from_numpy(guides).pin_memory(), out=host)
        assert ctx.launch_count > n0
        for k in ("disp", "f32", "u16", "out4k"):
            a, b = dev[k].cpu(), host[k]
            assert torch.equal(a.view(torch.int16) if a.dtype == torch.uint16 else a,
                               b.view(torch.int16) if b.dtype == torch.uint16 else b), k


def test_bad_arguments_raise():
    with pytest.raises(ValueError):
        nv.Context(66, 20, nv.SgbmParams(numDisparities=64))       # cv2.error site
    with pytest.raises(ValueError):
        nv.Context(200, 20, nv.SgbmParams(numDisparities=40))
    with nv.Context(200, 20, nv.SgbmParams()) as ctx:
        with pytest.raises(ValueError):
            ctx.split_gray(torch.zeros((1, 20, 401, 3), dtype=torch.uint8).cuda(), False)   # odd SBS width
        with pytest.raises(ValueError):
            ctx.sgbm_compute(torch.zeros((2, 20, 200), dtype=torch.uint8).cuda(),
                             torch.zeros((2, 20, 200), dtype=torch.uint8).cuda())           # batch > max_batch


def test_two_lanes_from_two_host_threads():
    """bench.py's lane model: one context + stream per host thread, synchronous host entry points."""
    import threading
    W, H, D, B = 320, 96, 64, 2
    frames = [np.stack([synthetic.sbs_frame(40 + k, t, W, H, D) for t in range(B)]) for k in range(2)]
    outs = [torch.empty((B, H, W), dtype=torch.int16).pin_memory() for _ in range(2)]
    errs = []

    def lane(k):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st), nv.Context(W, H, nv.SgbmParams(numDisparities=D), max_batch=B) as ctx:
                for _ in range(3):
                    ctx.depth_frames_host(torch.from_numpy(frames[k]).pin_memory(), False, out={"disp": outs[k]})
        except Exception as e:          # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=lane, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    m = cv2_chain.make_matcher(D, 0)
    for k in range(2):
        for b in range(B):
            l, r = cv2_chain.split_sbs_frame(frames[k][b], False)
            assert np.array_equal(outs[k][b].numpy(), m.compute(cv2_chain.to_gray(l), cv2_chain.to_gray(r)))


def test_fused_sweep_reports_clusters_and_matches_unfused_shapes():
    """Shapes one cluster's shared memory can hold take the cluster-fused sweep (9 or 8 CTAs for D <= 128,
    16 CTAs for D = 256); wider ones take the per-direction kernels.  Both must equal cv2.  Includes widths where every
    warp of the cluster gets a balanced share of 3+ columns (CPW or CPW - 1) and widths that fill warps in order."""
    for (W, H, D, expect_fused) in ((400, 40, 128, True), (600, 30, 256, True), (2100, 6, 256, True), (2300, 6, 256, False),
                                    (2200, 6, 128, True), (2500, 6, 128, False), (1100, 9, 64, True), (1300, 7, 128, True),
                                    (1000, 8, 128, True)):
        left, right, _ = synthetic.stereo_pair(8, 0, W, H, D)
        with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=1)) as ctx:
            d = ctx.sgbm_compute(torch.from_numpy(left)[None].cuda(), torch.from_numpy(right)[None].cuda())[0].cpu().numpy()
            assert (ctx.fused_sweep_clusters > 0) == expect_fused, (W, H, D, ctx.fused_sweep_clusters)
        assert np.array_equal(d, cv2_chain.make_matcher(D, 1).compute(left, right))


def test_full_size_cfg5_vs_cv2():
    """BASELINE configs[4] at full size, bit-exact against the reference's matcher: 1920x1080/eye, D=256,
    MODE_HH (8 paths), speckle filter on (depth.py:315-325 with numDisparities / mode opened up, :341)."""
    W, H, D = 1920, 1080, 256
    frame = synthetic.sbs_frame(13, 0, W, H, D)
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=1)) as ctx:
        res = ctx.depth_frames(torch.from_numpy(frame)[None].cuda(), False, want=("disp", "u16"))
        disp, u16 = res["disp"][0].cpu().numpy(), _u16(res["u16"][0])
        assert ctx.fused_sweep_clusters > 0
    m = cv2_chain.make_matcher(D, 1)
    l, r = cv2_chain.split_sbs_frame(frame, False)
    ref = m.compute(cv2_chain.to_gray(l), cv2_chain.to_gray(r))
    assert np.array_equal(disp, ref)
    assert np.array_equal(u16, cv2_chain.normalize_u16(cv2_chain.depth_from_sbs(frame, m, False)))
    assert (disp != -16).mean() > 0.4


def test_bench_call_full_size_host_and_device_vs_oracle():
    """The exact calls bench.py times, at its shapes: depth_frames(sbs, guide) (device-resident) and
    depth_frames_host(sbs, guide) (pinned host buffers) on full-SBS 3840x1080 + a 3840x2160 guide.  The 4K uint16
    output must be within 1 LSB of oracle.guided fed with the cv2 chain's uint16 depth, and both entry points
    must agree bit for bit."""
    W, H, D, B = 1920, 1080, 128, 2
    frames = np.stack([synthetic.sbs_frame(11, t, W, H, D) for t in range(B)])
    guides = np.stack([synthetic.guide_frame(11, t, 2 * W, 2 * H) for t in range(B)])
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D), max_batch=B) as ctx:
        dev = ctx.depth_frames(torch.from_numpy(frames).cuda(), False, torch.from_numpy(guides).cuda(), 8, 1e-3,
                               want=("u16",))
        dev4k, devu16 = _u16(dev["out4k"]), _u16(dev["u16"])
        out_h = torch.empty((B, 2 * H, 2 * W), dtype=torch.uint16).pin_memory()
        ctx.depth_frames_host(torch.from_numpy(frames).pin_memory(), False, torch.from_numpy(guides).pin_memory(), 8, 1e-3,
                              out={"out4k": out_h})
        host4k = out_h.numpy().view(np.uint16)
    assert np.array_equal(dev4k, host4k)
    m = cv2_chain.make_matcher(D, 0)
    for b in range(B):
        ref_u16 = cv2_chain.normalize_u16(cv2_chain.depth_from_sbs(frames[b], m, False))
        assert np.array_equal(devu16[b], ref_u16)
        if b == 0:      # the float64 oracle takes ~20 s per 4K frame
            _, o4 = og.guided_upscale(ref_u16, guides[b], 8, 1e-3)
            assert np.abs(host4k[b].astype(np.int64) - o4.astype(np.int64)).max() <= 1


def test_async_host_calls_from_one_submit_thread():
    """bench.py's end-to-end lane model: ONE host thread submits v3d_depth_frames_host_async on several contexts
    (each on its own stream) and sleeps in v3d_host_wait; results equal the synchronous device path.  The
    copy-only twin moves the same bytes and leaves the kernels out (launch count unchanged)."""
    W, H, D, B, L = 320, 96, 64, 3, 3
    frames = [np.stack([synthetic.sbs_frame(50 + k, t, W, H, D) for t in range(B)]) for k in range(L)]
    guides = [np.stack([synthetic.guide_frame(50 + k, t, 2 * W, 2 * H) for t in range(B)]) for k in range(L)]
    ctxs = [nv.Context(W, H, nv.SgbmParams(numDisparities=D), max_batch=B) for _ in range(L)]
    streams = [torch.cuda.Stream() for _ in range(L)]
    fh = [torch.from_numpy(f).pin_memory() for f in frames]
    gh = [torch.from_numpy(g).pin_memory() for g in guides]
    outs = [dict(disp=torch.empty((B, H, W), dtype=torch.int16).pin_memory(),
                 out4k=torch.zeros((B, 2 * H, 2 * W), dtype=torch.uint16).pin_memory()) for _ in range(L)]
    try:
        for _ in range(3):                   # resubmission after a wait, several steps deep
            for k in range(L):
                ctxs[k].host_wait()
                with torch.cuda.stream(streams[k]):
                    ctxs[k].depth_frames_host(fh[k], False, gh[k], 8, 1e-3, out=outs[k], wait=False)
        for k in range(L):
            ctxs[k].host_wait()
            ctxs[k].host_wait()              # idempotent
        for k in range(L):
            ref = ctxs[k].depth_frames(torch.from_numpy(frames[k]).cuda(), False, torch.from_numpy(guides[k]).cuda(), 8, 1e-3,
                                       want=("disp",))
            torch.cuda.synchronize()
            assert torch.equal(ref["disp"].cpu(), outs[k]["disp"])
            assert torch.equal(ref["out4k"].cpu().view(torch.int16), outs[k]["out4k"].view(torch.int16))
        n0 = ctxs[0].launch_count
        with torch.cuda.stream(streams[0]):
            ctxs[0].host_copy_only(fh[0], gh[0], out={"out4k": outs[0]["out4k"]})
        ctxs[0].host_wait()
        assert ctxs[0].launch_count == n0
        # upscale-only host entry (cfg3): uint16 depth maps of the context's eye size + guide -> 4K
        d = np.stack([synthetic.depth_u16(5, t, W, H) for t in range(B)])
        dh = torch.from_numpy(d.view(np.int16)).view(torch.uint16).pin_memory()
        o = torch.empty((B, 2 * H, 2 * W), dtype=torch.uint16).pin_memory()
        ctxs[1].guided_upscale_host(dh, gh[1], o, 8, 1e-3)
        want = ctxs[1].guided_upscale(dh.cuda(), gh[1].cuda(), 8, 1e-3)
        assert torch.equal(want.cpu().view(torch.int16), o.view(torch.int16))
    finally:
        for c in ctxs:
            c.close()


def test_two_host_calls_in_flight_per_context():
    """A context keeps TWO asynchronous host calls in flight (own staging buffers per call): submit k+1, then wait for
    k.  Every call's outputs equal the synchronous device path on that call's inputs, whatever the order of
    submissions, waits and implicit waits (a third submission first waits for the oldest call)."""
    W, H, D, B, N = 320, 96, 64, 3, 5
    frames = [np.stack([synthetic.sbs_frame(70 + k, t, W, H, D) for t in range(B)]) for k in range(N)]
    guides = [np.stack([synthetic.guide_frame(70 + k, t, 2 * W, 2 * H) for t in range(B)]) for k in range(N)]
    fh = [torch.from_numpy(f).pin_memory() for f in frames]
    gh = [torch.from_numpy(g).pin_memory() for g in guides]
    outs = [dict(disp=torch.full((B, H, W), -7, dtype=torch.int16).pin_memory(),
                 u16=torch.zeros((B, H, W), dtype=torch.uint16).pin_memory(),
                 out4k=torch.zeros((B, 2 * H, 2 * W), dtype=torch.uint16).pin_memory()) for _ in range(N)]
    ctx = nv.Context(W, H, nv.SgbmParams(numDisparities=D), max_batch=B)
    stream = torch.cuda.Stream()
    try:
        assert ctx.host_pending == 0
        with torch.cuda.stream(stream):
            for k in range(N):
                if k % 2:                                    # alternate: with and without the upscale step
                    ctx.depth_frames_host(fh[k], False, gh[k], 8, 1e-3, out=outs[k], wait=False)
                else:
                    ctx.depth_frames_host(fh[k], False, out={"disp": outs[k]["disp"], "u16": outs[k]["u16"]}, wait=False)
                assert ctx.host_pending == min(k + 1, 2)     # the third submission waited for the oldest call
                if k == 1:
                    ctx.host_wait_oldest()                   # explicit: call 0 is complete, call 1 still in flight
                    assert ctx.host_pending == 1
        ctx.host_wait()
        assert ctx.host_pending == 0
        ctx.host_wait_oldest()                               # nothing pending: no-op
        for k in range(N):
            want = ("disp", "u16")
            if k % 2:
                ref = ctx.depth_frames(torch.from_numpy(frames[k]).cuda(), False, torch.from_numpy(guides[k]).cuda(), 8, 1e-3, want=want)
            else:
                ref = ctx.depth_frames(torch.from_numpy(frames[k]).cuda(), False, want=want)
            torch.cuda.synchronize()
            assert torch.equal(ref["disp"].cpu(), outs[k]["disp"]), k
            assert torch.equal(ref["u16"].cpu().view(torch.int16), outs[k]["u16"].view(torch.int16)), k
            if k % 2:
                assert torch.equal(ref["out4k"].cpu().view(torch.int16), outs[k]["out4k"].view(torch.int16)), k
        # upscale-only calls share one workspace buffer for the uploaded depth maps: still exact back to back
        d = [np.stack([synthetic.depth_u16(9 + k, t, W, H) for t in range(B)]) for k in range(3)]
        dh = [torch.from_numpy(x.view(np.int16)).view(torch.uint16).pin_memory() for x in d]
        o = [torch.zeros((B, 2 * H, 2 * W), dtype=torch.uint16).pin_memory() for _ in range(3)]
        with torch.cuda.stream(stream):
            for k in range(3):
                ctx.guided_upscale_host(dh[k], gh[k], o[k], 8, 1e-3, wait=False)
        ctx.host_wait()
        for k in range(3):
            want = ctx.guided_upscale(dh[k].cuda(), gh[k].cuda(), 8, 1e-3)
            torch.cuda.synchronize()
            assert torch.equal(want.cpu().view(torch.int16), o[k].view(torch.int16)), k
    finally:
        ctx.close()


def test_row_checkpoints_cover_every_width_remainder():
    """The left-to-right direction is re-run from a checkpoint per 8-pixel chunk, chunks counted from the right
    end of the row: every W1 % 8 (and rows shorter than one chunk) must give cv2's disparities, S taps included."""
    for W1 in (3, 8, 9, 15, 16, 17, 23, 31, 40, 41):
        _run_stages(W1 + 64, 12, 64, 0)
    _run_stages(128 + 21, 9, 128, 1)
    _run_stages(256 + 13, 6, 256, 0)


@pytest.mark.parametrize("minD,W,H,D,mode", [(16, 300, 40, 64, 0), (5, 260, 30, 64, 1), (48, 420, 24, 128, 0), (-16, 300, 40, 64, 0),
                                             (-5, 231, 17, 48, 1), (-70, 300, 20, 64, 0), (-64, 280, 16, 64, 0), (100, 700, 12, 256, 0),
                                             (1, 150, 9, 16, 0)])
def test_min_disparity_matches_cv2(minD, W, H, D, mode):
    """minDisparity != 0 (outside the reference's literals, inside cv2's contract): window [max(minD + D, 0),
    W + min(minD, 0)), invalid value (minD - 1) * 16, disparities offset by minD -- final map bit-exact vs cv2, with and
    without the speckle filter."""
    left, right, _ = synthetic.stereo_pair(21, 0, W, H, D)
    right = np.roll(right, minD // 2, axis=1)                 # some matches inside the shifted range
    for kw in (dict(), dict(speckleWindowSize=0, uniquenessRatio=0)):
        with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=mode, minDisparity=minD, **kw)) as ctx:
            lt, rt = torch.from_numpy(left)[None].cuda(), torch.from_numpy(right)[None].cuda()
            d = ctx.sgbm_compute(lt, rt)[0].cpu().numpy()
            f32, u16 = ctx.postprocess(torch.from_numpy(d)[None].cuda())
        ref = cv2_chain.make_matcher(D, mode, minDisparity=minD, **kw).compute(left, right)
        assert np.array_equal(d, ref), (minD, kw, float((d != ref).mean()))
        depth = ref.astype(np.float32) / 16.0                  # depth.py:341, 374, 400-403 on that map
        depth[depth <= 0] = 0
        assert np.array_equal(f32[0].cpu().numpy(), depth)
        assert np.array_equal(_u16(u16)[0], cv2_chain.normalize_u16(depth))


def test_very_wide_rows_match_cv2():
    """Eye widths beyond 8192 columns (the disparity-selection kernel keeps a row of votes in shared memory, opted in
    above 48 KB; the speckle filter, the per-direction vertical kernels and the cost strips take their generic paths)."""
    for (W, H, D, mode) in ((9000, 3, 64, 0), (12345, 2, 128, 1)):
        left, right, _ = synthetic.stereo_pair(17, 0, W, H, D)
        with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=mode)) as ctx:
            d = ctx.sgbm_compute(torch.from_numpy(left)[None].cuda(), torch.from_numpy(right)[None].cuda())[0].cpu().numpy()
        assert np.array_equal(d, cv2_chain.make_matcher(D, mode).compute(left, right)), (W, H, D, mode)
    with pytest.raises(ValueError):
        nv.Context(40000, 2, nv.SgbmParams())


def test_context_reuse_with_varying_batches_and_streams():
    """One context, calls with different batch sizes on different streams (workspace buffers, checkpoints, side stream
    and cluster choice are per context and must not leak state between calls)."""
    W, H, D, B = 520, 48, 128, 5
    frames = np.stack([synthetic.sbs_frame(60, t, W, H, D) for t in range(B)])
    m = cv2_chain.make_matcher(D, 1)
    refs = []
    for b in range(B):
        l, r = cv2_chain.split_sbs_frame(frames[b], False)
        refs.append(m.compute(cv2_chain.to_gray(l), cv2_chain.to_gray(r)))
    dev = torch.from_numpy(frames).cuda()
    streams = [torch.cuda.Stream() for _ in range(2)]
    with nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=1), max_batch=B) as ctx:
        for k, (lo, n) in enumerate(((0, 5), (3, 2), (1, 4), (4, 1), (0, 3))):
            with torch.cuda.stream(streams[k % 2]):
                streams[k % 2].wait_stream(torch.cuda.current_stream())
                res = ctx.depth_frames(dev[lo:lo + n].contiguous(), False, want=("disp",))
            streams[k % 2].synchronize()
            got = res["disp"].cpu().numpy()
            for i in range(n):
                assert np.array_equal(got[i], refs[lo + i]), (k, lo, n, i)
        with pytest.raises(ValueError):
            ctx.depth_frames(torch.zeros((B + 1, H, 2 * W, 3), dtype=torch.uint8, device="cuda"), False, want=())
