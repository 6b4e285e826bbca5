"""Pins the oracle: the C restatement (oracle/sgbm_oracle.c) must equal the live cv2 -- the
reference's own arithmetic (depth.py:265-266, 274-275, 315-325, 337-341) -- bit for bit."""
import cv2
import numpy as np
import pytest

from oracle import cv2_chain, guided
from oracle import sgbm as osg
from video_3d_pipeline import synthetic

CASES = [  # W, H, D, mode
    (96, 40, 32, 0), (131, 33, 48, 0), (96, 40, 32, 1), (200, 120, 64, 0), (200, 120, 64, 1),
    (37, 5, 16, 0), (19, 1, 16, 0), (23, 2, 16, 1), (83, 3, 64, 0), (320, 180, 128, 0), (300, 60, 256, 1),
]


@pytest.mark.parametrize("W,H,D,mode", CASES)
def test_sgbm_compute_matches_cv2(W, H, D, mode):
    left, right, _ = synthetic.stereo_pair(1, 0, W, H, D)
    got = osg.sgbm_compute(left, right, osg.Params(numDisparities=D, mode=mode))
    ref = cv2_chain.make_matcher(D, mode).compute(left, right)
    assert got.dtype == np.int16 and np.array_equal(got, ref)
    assert (ref[:, :D] == -16).all()          # columns [0, D) are always invalid


@pytest.mark.parametrize("kind", ["uniform", "binary", "flat"])
def test_sgbm_degenerate_textures(kind):
    rng = np.random.default_rng(3)
    W, H, D = 90, 30, 32
    if kind == "uniform":
        left, right = rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)
    elif kind == "binary":
        left = (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
        right = np.roll(left, -3, axis=1)
    else:
        left = np.full((H, W), 77, np.uint8)
        right = left.copy()
    for mode in (0, 1):
        got = osg.sgbm_compute(left, right, osg.Params(numDisparities=D, mode=mode))
        assert np.array_equal(got, cv2_chain.make_matcher(D, mode).compute(left, right))


def test_speckle_off_and_other_params():
    left, right, _ = synthetic.stereo_pair(2, 1, 160, 60, 48)
    for kw in (dict(speckleWindowSize=0), dict(uniquenessRatio=0), dict(uniquenessRatio=25),
               dict(disp12MaxDiff=3), dict(disp12MaxDiff=-1), dict(P1=100, P2=900), dict(blockSize=3),
               dict(blockSize=7), dict(preFilterCap=31)):
        got = osg.sgbm_compute(left, right, osg.Params(numDisparities=48, **kw))
        assert np.array_equal(got, cv2_chain.make_matcher(48, 0, **kw).compute(left, right)), kw


def test_width_precondition_raises_like_cv2():
    left = np.zeros((10, 18), np.uint8)
    with pytest.raises(ValueError):
        osg.sgbm_compute(left, left, osg.Params(numDisparities=16))
    with pytest.raises(cv2.error):
        cv2_chain.make_matcher(16, 0).compute(left, left)


def test_gray_exhaustive_sample_and_split():
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (64, 4096, 3), dtype=np.uint8)
    assert np.array_equal(osg.bgr_to_gray(bgr), cv2_chain.to_gray(bgr))
    frame = synthetic.sbs_frame(3, 1, 240, 135, 32)
    for uns in (False, True):
        l, r = cv2_chain.split_sbs_frame(frame, uns)
        ol, orr = osg.split_gray(frame, uns)
        assert np.array_equal(ol, cv2_chain.to_gray(l)) and np.array_equal(orr, cv2_chain.to_gray(r))
    with pytest.raises(ValueError):
        cv2_chain.split_sbs_frame(np.zeros((4, 7, 3), np.uint8))
    with pytest.raises(ValueError):
        osg.split_gray(np.zeros((4, 7, 3), np.uint8), False)


def test_lanczos_unsqueeze_matches_cv2():
    rng = np.random.default_rng(1)
    for shape in ((50, 77, 3), (9, 5, 3), (30, 960, 3), (12, 33)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        ref = cv2.resize(img, (2 * shape[1], shape[0]), interpolation=cv2.INTER_LANCZOS4)
        assert np.array_equal(osg.unsqueeze_x2(img), ref)


def test_median_and_speckle_match_cv2():
    rng = np.random.default_rng(2)
    d = rng.integers(-16, 1024, (60, 90)).astype(np.int16)
    d[rng.random((60, 90)) < 0.3] = -16
    assert np.array_equal(osg.median3(d), cv2.medianBlur(d, 3))
    for seed in range(3):
        rng = np.random.default_rng(10 + seed)
        d2 = (rng.integers(0, 4, (80, 120)) * 600).astype(np.int16)
        d2[rng.random((80, 120)) < 0.35] = -16
        ref = d2.copy()
        cv2.filterSpeckles(ref, -16, 20, 512)
        assert np.array_equal(osg.filter_speckles(d2, -16, 20, 512), ref)


def test_epilogue_matches_numpy_chain():
    left, right, _ = synthetic.stereo_pair(4, 0, 200, 90, 64)
    d = cv2_chain.make_matcher(64, 0).compute(left, right)
    ref = d.astype(np.float32) / 16.0          # depth.py:341
    ref[ref <= 0] = 0                          # depth.py:374
    f = osg.disp_to_float(d)
    assert np.array_equal(f, ref)
    assert np.array_equal(osg.normalize_u16(f), cv2_chain.normalize_u16(ref))
    flat = np.full((4, 5), 3.0, np.float32)
    assert (osg.normalize_u16(flat) == 0).all() and (cv2_chain.normalize_u16(flat) == 0).all()


def test_stage_taps_are_consistent():
    left, right, _ = synthetic.stereo_pair(5, 0, 150, 40, 32)
    p = osg.Params(numDisparities=32)
    disp, taps = osg.sgbm_compute(left, right, p, taps=True)
    assert np.array_equal(taps["C"], osg.cost_volume(left, right, p))
    S, Su = osg.aggregate(taps["C"], p, unsaturated=True)
    assert np.array_equal(S, taps["S"]) and np.array_equal(np.minimum(Su, 32767), S)
    assert np.array_equal(sum(osg.aggregate_one(taps["C"], p, k).astype(np.uint32) for k in range(5)), Su)
    assert np.array_equal(osg.select(S, 150, p), taps["raw"])
    assert np.array_equal(osg.median3(taps["raw"]), taps["median"])
    assert np.array_equal(osg.filter_speckles(taps["median"], -16, 100, 512), disp)


def test_guided_oracle_building_blocks():
    rng = np.random.default_rng(0)
    a = rng.random((50, 70))
    assert np.abs(guided.box_mean(a, 8) - cv2.boxFilter(a, -1, (17, 17), borderType=cv2.BORDER_REFLECT)).max() < 1e-12
    d = synthetic.depth_u16(1, 0, 96, 54).astype(np.float64) / 65535
    up = guided.bilinear_upsample(d, 108, 192)
    assert np.abs(up - cv2.resize(d, (192, 108), interpolation=cv2.INTER_LINEAR)).max() < 1e-12
    # constant guide: a = 0 and q = box(box(p))
    g = np.full((108, 192, 3), 90, np.uint8)
    q, out = guided.guided_upscale(synthetic.depth_u16(1, 0, 96, 54), g, 4, 1e-3)
    assert np.abs(q - guided.box_mean(guided.box_mean(up, 4), 4)).max() < 1e-9
    # the fp32 cv2 port used as the timed CPU baseline agrees with the float64 definition
    gd = synthetic.guide_frame(1, 0, 192, 108)
    q64, o64 = guided.guided_upscale(synthetic.depth_u16(1, 0, 96, 54), gd)
    q32, o32 = guided.guided_upscale_cv2(synthetic.depth_u16(1, 0, 96, 54), gd)
    assert np.abs(q64 - q32).max() * 65535 < 0.5
    assert np.abs(o64.astype(np.int64) - o32.astype(np.int64)).max() <= 1


def test_guided_oracle_matches_bruteforce_definition():
    """The guided-upscale oracle has no reference to pin it (upstream ships no guided filter), so it is
    at least checked against an independent, literal evaluation of He/Sun/Tang eq. (19)-(21): per-pixel
    window loops over a symmetric-padded image and np.linalg.solve, no separable / cumulative sums."""
    rng = np.random.default_rng(7)
    h, w, r, eps = 9, 11, 2, 1e-3
    depth = rng.integers(0, 65536, (h, w)).astype(np.uint16)
    guide = rng.integers(0, 256, (2 * h, 2 * w, 3), dtype=np.uint8)
    q, out = guided.guided_upscale(depth, guide, r, eps)

    H, W = guide.shape[:2]
    p = guided.bilinear_upsample(depth.astype(np.float64) / 65535.0, H, W)
    I = guide.astype(np.float64) / 255.0
    pad = lambda a: np.pad(a, [(r, r), (r, r)] + [(0, 0)] * (a.ndim - 2), mode="symmetric")
    Ip, pp = pad(I), pad(p)
    a = np.zeros((H, W, 3))
    b = np.zeros((H, W))
    for y in range(H):
        for x in range(W):
            wi = Ip[y:y + 2 * r + 1, x:x + 2 * r + 1].reshape(-1, 3)
            wp = pp[y:y + 2 * r + 1, x:x + 2 * r + 1].reshape(-1)
            mu, pbar = wi.mean(0), wp.mean()
            sigma = (wi - mu).T @ (wi - mu) / len(wp)
            cov = ((wi - mu) * (wp - pbar)[:, None]).mean(0)
            a[y, x] = np.linalg.solve(sigma + eps * np.eye(3), cov)
            b[y, x] = pbar - a[y, x] @ mu
    ap, bp = pad(a), pad(b)
    q_ref = np.zeros((H, W))
    for y in range(H):
        for x in range(W):
            abar = ap[y:y + 2 * r + 1, x:x + 2 * r + 1].reshape(-1, 3).mean(0)
            q_ref[y, x] = abar @ I[y, x] + bp[y:y + 2 * r + 1, x:x + 2 * r + 1].mean()
    assert np.abs(q - q_ref).max() < 1e-9
    assert np.array_equal(out, np.floor(np.clip(q_ref, 0, 1) * 65535 + 0.5).astype(np.uint16))
