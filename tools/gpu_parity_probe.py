"""Stage-by-stage GPU-vs-oracle probe (development aid; the real tests are tests/test_gpu_*.py)."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "video-3d-pipeline_b200"))
import numpy as np, torch, cv2
from oracle import sgbm as osg, cv2_chain, guided as og
from video_3d_pipeline import _native as nv, synthetic

def run_case(W, H, D, mode, B=1, speckle=100, unsq=False):
    p_o = osg.Params(numDisparities=D, mode=mode, speckleWindowSize=speckle)
    p_n = nv.SgbmParams(numDisparities=D, mode=mode, speckleWindowSize=speckle)
    src_w = W // 2 if unsq else W
    frames = np.stack([synthetic.sbs_frame(7, t, src_w, H, D if not unsq else D // 2) for t in range(B)])
    ctx = nv.Context(W, H, p_n, max_batch=B)
    ctx.set_debug_taps(True)
    sbs = torch.from_numpy(frames).cuda()
    l, r = ctx.split_gray(sbs, unsq)
    res = {}
    ol, orr = zip(*[osg.split_gray(f, unsq) for f in frames])
    res['gray'] = int((l.cpu().numpy() != np.stack(ol)).sum() + (r.cpu().numpy() != np.stack(orr)).sum())
    disp = ctx.sgbm_compute(l, r)
    torch.cuda.synchronize()
    C = ctx.debug_tap(0, B).cpu().numpy().view(np.uint16)
    S = ctx.debug_tap(1, B).cpu().numpy().view(np.uint16)
    raw = ctx.debug_tap(2, B).cpu().numpy()
    med = ctx.debug_tap(3, B).cpu().numpy()
    dn = disp.cpu().numpy()
    for b in range(B):
        od, taps = osg.sgbm_compute(ol[b], orr[b], p_o, taps=True)
        _, Su = osg.aggregate(taps['C'], p_o, unsaturated=True)
        res.setdefault('C', 0); res['C'] += int((C[b] != taps['C']).sum())
        res.setdefault('S', 0); res['S'] += int((S[b] != Su).sum())
        res.setdefault('raw', 0); res['raw'] += int((raw[b] != taps['raw']).sum())
        res.setdefault('med', 0); res['med'] += int((med[b] != taps['median']).sum())
        res.setdefault('disp', 0); res['disp'] += int((dn[b] != od).sum())
        m = cv2_chain.make_matcher(D, mode, speckleWindowSize=speckle)
        res.setdefault('cv2', 0); res['cv2'] += int((dn[b] != m.compute(ol[b], orr[b])).sum())
    f32, u16 = ctx.postprocess(disp)
    of = np.stack([osg.disp_to_float(d) for d in dn])
    res['f32'] = int((f32.cpu().numpy() != of).sum())
    res['u16'] = int((u16.cpu().numpy().view(np.uint16) != np.stack([osg.normalize_u16(f) for f in of])).sum())
    ctx.close()
    print(f"W{W} H{H} D{D} mode{mode} B{B} unsq{unsq}:", json.dumps(res), flush=True)
    return res

def run_guided(w, h, gw, gh, r=8, eps=1e-3, B=1):
    ctx = nv.Context(max(w, 80), h, nv.SgbmParams(), max_batch=B)
    d = np.stack([synthetic.depth_u16(3, t, w, h) for t in range(B)])
    g = np.stack([synthetic.guide_frame(3, t, gw, gh) for t in range(B)])
    out, q = ctx.guided_upscale(torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16), torch.from_numpy(g).cuda(), r, eps, want_q=True)
    out = out.cpu().numpy().view(np.uint16); q = q.cpu().numpy()
    errq = 0.0; erru = 0
    for b in range(B):
        oq, ou = og.guided_upscale(d[b], g[b], r, eps)
        errq = max(errq, float(np.abs(q[b] - oq).max()))
        erru = max(erru, int(np.abs(out[b].astype(np.int64) - ou.astype(np.int64)).max()))
    ctx.close()
    print(f"guided {w}x{h}->{gw}x{gh} r{r}: max|q| err {errq:.3e} ({errq*65535:.3f} LSB16), max u16 diff {erru}", flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), nv.lib().v3d_version())
    run_case(200, 120, 64, 0)
    run_case(200, 120, 64, 1)
    run_case(331, 77, 128, 0, B=2)
    run_case(400, 50, 256, 1)
    run_case(67, 20, 64, 0)
    run_case(131, 1, 64, 0)
    run_case(131, 3, 64, 1)
    run_case(480, 270, 64, 0, unsq=True)
    run_case(640, 360, 128, 0, B=3)
    run_guided(96, 54, 192, 108)
    run_guided(100, 60, 230, 131, r=4)
    run_guided(480, 270, 960, 540, B=2)
    t = time.time(); run_case(1920, 1080, 128, 0); print("full-size case wall", time.time() - t)
