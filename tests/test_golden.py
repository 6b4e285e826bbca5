"""Committed golden vectors (tests/golden/*.npz, produced from the live cv2 by make_golden.py):
the oracle must reproduce them on CPU, the CUDA path must reproduce them on the GPU."""
from pathlib import Path

import numpy as np
import pytest

from oracle import guided as og
from oracle import sgbm as osg

G = Path(__file__).resolve().parent / "golden"
SGBM = ["sgbm_d64_mode0", "sgbm_d64_mode1", "sgbm_d128_mode0"]


@pytest.mark.parametrize("name", SGBM)
def test_oracle_reproduces_golden_sgbm(name):
    z = np.load(G / f"{name}.npz")
    got = osg.sgbm_compute(z["left"], z["right"], osg.Params(numDisparities=int(z["D"]), mode=int(z["mode"])))
    assert np.array_equal(got, z["disp"])


def test_oracle_reproduces_golden_chain():
    z = np.load(G / "chain_sbs.npz")
    for uns in (0, 1):
        l, r = osg.split_gray(z["frame"], bool(uns))
        assert np.array_equal(l, z[f"gray_left_{uns}"])
        f = osg.disp_to_float(osg.sgbm_compute(l, r, osg.Params(numDisparities=64)))
        assert np.array_equal(f, z[f"depth_f32_{uns}"])
        assert np.array_equal(osg.normalize_u16(f), z[f"depth_u16_{uns}"])


def test_oracle_reproduces_golden_guided():
    z = np.load(G / "guided_r8.npz")
    q, o = og.guided_upscale(z["depth"], z["guide"], 8, 1e-3)
    assert np.abs(q - z["q"]).max() < 1e-12 and np.array_equal(o, z["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", SGBM)
def test_cuda_reproduces_golden_sgbm(name):
    import torch
    from video_3d_pipeline import _native as nv
    z = np.load(G / f"{name}.npz")
    H, W = z["left"].shape
    with nv.Context(W, H, nv.SgbmParams(numDisparities=int(z["D"]), mode=int(z["mode"]))) as ctx:
        d = ctx.sgbm_compute(torch.from_numpy(z["left"])[None].cuda(), torch.from_numpy(z["right"])[None].cuda())
    assert np.array_equal(d[0].cpu().numpy(), z["disp"])


@pytest.mark.gpu
def test_cuda_reproduces_golden_chain_and_guided():
    import torch
    from video_3d_pipeline import _native as nv
    z = np.load(G / "chain_sbs.npz")
    frame = z["frame"]
    H, Ws = frame.shape[:2]
    for uns in (0, 1):
        We = Ws if uns else Ws // 2
        with nv.Context(We, H, nv.SgbmParams(numDisparities=64)) as ctx:
            res = ctx.depth_frames(torch.from_numpy(frame)[None].cuda(), bool(uns), want=("f32", "u16"))
            assert np.array_equal(res["f32"][0].cpu().numpy(), z[f"depth_f32_{uns}"])
            assert np.array_equal(res["u16"][0].cpu().numpy().view(np.uint16), z[f"depth_u16_{uns}"])
    g = np.load(G / "guided_r8.npz")
    with nv.Context(80, 8, nv.SgbmParams()) as ctx:
        out, q = ctx.guided_upscale(torch.from_numpy(g["depth"].view(np.int16))[None].cuda().view(torch.uint16),
                                    torch.from_numpy(g["guide"])[None].cuda(), 8, 1e-3, want_q=True)
    assert np.abs(q[0].cpu().numpy() - g["q"]).max() < 0.5 / 65535
    assert np.abs(out[0].cpu().numpy().view(np.uint16).astype(np.int64) - g["out"].astype(np.int64)).max() <= 1
