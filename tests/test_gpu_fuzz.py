"""Seeded randomised sweep of shapes and matcher parameters: CUDA path vs the live cv2, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import cv2_chain
from video_3d_pipeline import _native as nv
from video_3d_pipeline import synthetic

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    D = int(rng.choice([16, 32, 48, 64, 80, 96, 128, 144, 192, 256]))
    bs = int(rng.choice([1, 3, 5, 5, 5, 7]))
    mode = int(rng.integers(0, 2))
    W = int(rng.integers(D + bs // 2 + 1, D + 420))
    H = int(rng.integers(1, 48))
    P1 = int(rng.integers(1, 900))
    P2 = int(rng.integers(P1 + 1, 3000))
    kw = dict(blockSize=bs, P1=P1, P2=P2, disp12MaxDiff=int(rng.integers(-1, 4)),
              preFilterCap=int(rng.choice([0, 7, 15, 31, 63])), uniquenessRatio=int(rng.integers(0, 30)),
              speckleWindowSize=int(rng.choice([0, 20, 100, 300])), speckleRange=int(rng.integers(1, 40)))
    kind = int(rng.integers(0, 4))
    return D, mode, W, H, kw, kind, rng


@pytest.mark.parametrize("seed", range(48))
def test_random_shapes_and_parameters_match_cv2(seed):
    D, mode, W, H, kw, kind, rng = _case(seed)
    if kind == 0:      # textured scene with true disparities
        left, right, _ = synthetic.stereo_pair(seed, 0, W, H, D)
    elif kind == 1:    # independent noise: almost everything fails uniqueness / LR
        left = rng.integers(0, 256, (H, W), dtype=np.uint8)
        right = rng.integers(0, 256, (H, W), dtype=np.uint8)
    elif kind == 2:    # low-contrast ramp with a shift: many ties
        base = (np.arange(W + 8)[None, :] // 3 + np.arange(H)[:, None] // 2).astype(np.uint8)
        left, right = np.ascontiguousarray(base[:, :W]), np.ascontiguousarray(base[:, 4:W + 4])
    else:              # saturated blocks
        left = (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
        right = np.roll(left, -int(rng.integers(0, 9)), axis=1)
    ref = cv2_chain.make_matcher(D, mode, **kw).compute(left, right)
    try:
        ctx = nv.Context(W, H, nv.SgbmParams(numDisparities=D, mode=mode, **kw))
    except ValueError as e:            # documented limit: blockSize^2 * cost + P2 must fit the 16-bit state
        assert "too large" in str(e)
        pytest.skip(str(e))
    with ctx:
        got = ctx.sgbm_compute(torch.from_numpy(left)[None].cuda(), torch.from_numpy(right)[None].cuda())[0].cpu().numpy()
    assert np.array_equal(got, ref), dict(D=D, mode=mode, W=W, H=H, kind=kind, **kw)
