#!/usr/bin/env python
"""Headline benchmark: 1080p-SBS -> 4K depth frames/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg1|cfg2|cfg3|cfg4|cfg5]
                    [--impl ours|reference] [--lanes L] [--batch B] [--reps R]

One "step" = one pass of the depth hot path over one batch of synthetic frames per GPU.  The default
workload (cfg4 = BASELINE.json configs[1] + configs[2] fused, i.e. configs[3]'s per-GPU work):
full-SBS 3840x1080 BGR -> split + gray -> cv2-exact SGBM (D=128, 5 paths, uniqueness, sub-pixel, LR
check, 3x3 median, speckle) -> /16, clamp, per-frame min-max -> uint16 -> guided upscale to 3840x2160
uint16 with a 4K RGB guide (r=8, eps=1e-3).  --config selects the other BASELINE.json configurations
(same JSON line, that configuration's SURVEY 8(d) algorithmic bytes).

Printed JSON line (rank 0):
  value    frames/s, inputs already resident in HBM, CUDA-event timed, max over ranks; median of --reps
           repetitions of K steps (all repetitions listed under "repetitions")
  e2e      the same through the host-buffer C-ABI call (v3d_depth_frames_host_async + v3d_host_wait):
           pinned host inputs, H2D + D2H copies inside the timed region; host_copy_ceiling = the same
           copies with no kernel in between (what the host side of the box allows)
  roofline dominant kernel against the measured HBM peak (MEASURED_PEAKS.json) + the chain's measured DRAM
           traffic per frame (profiles/traffic_per_frame.json, from ncu)
  cpu_baseline  the reference's cv2 chain (+ CPU guided upscale port) on the host cores, bounded sample
  depth_only    (default config) the SGBM chain alone on the GPU and on the reference's cv2 code

Multi-GPU: frames shard by contiguous ranges with no data-path collective; ranks
only meet for the barrier and the max-over-ranks of the elapsed time.  scaling = "weak".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "video-3d-pipeline_b200"))

import numpy as np  # noqa: E402

METRIC = "1080p-SBS->4K depth frames/sec"
UNIT = "frames/s"
RADIUS, EPS = 8, 1e-3
SEED = 11


class Cfg:
    """One BASELINE.json configuration as a concrete run."""

    def __init__(self, name, eye_w, eye_h, D, mode, depth, upscale, workload):
        self.name, self.eye_w, self.eye_h, self.D, self.mode = name, eye_w, eye_h, D, mode
        self.depth, self.upscale, self.workload = depth, upscale, workload
        self.gw, self.gh = 2 * eye_w, 2 * eye_h
        self.sbs_bytes = 2 * eye_w * eye_h * 3 if depth else 0
        self.guide_bytes = self.gw * self.gh * 3 if upscale else 0
        self.out_bytes = self.gw * self.gh * 2 if upscale else 2 * eye_w * eye_h      # uint16 4K / int16 disparity
        self.depth_in_bytes = 0 if depth else 2 * eye_w * eye_h                       # cfg3: uint16 depth in
        # SURVEY 8(d) algorithmic bytes per frame
        if depth and upscale:
            self.alg_bytes = self.sbs_bytes + self.guide_bytes + self.out_bytes           # 53 913 600
        elif depth:
            self.alg_bytes = self.sbs_bytes + 2 * eye_w * eye_h                            # 16 588 800 / 8 294 400
        else:
            self.alg_bytes = 4 * eye_w * eye_h + self.guide_bytes + self.out_bytes         # 49 766 400 (8d counts an f32 map in)
        self.ndirs = 8 if mode == 1 else 5
        self.cells = (eye_w - D) * eye_h * D if depth else 0
        self.alg_iops = (31 + 9 * self.ndirs) * self.cells
        self.h2d = self.sbs_bytes + self.guide_bytes + self.depth_in_bytes
        self.d2h = self.out_bytes


CONFIGS = {
    "cfg1": Cfg("cfg1", 960, 1080, 64, 0, True, False,
                "cfg1: half-SBS 1920x1080 (960x1080/eye) SGBM numDisparities=64 MODE_SGBM (depth.py:315-325 literals), "
                "int16 disparity out"),
    "cfg2": Cfg("cfg2", 1920, 1080, 128, 0, True, False,
                "cfg2: full-SBS 3840x1080 (1920x1080/eye) SGBM numDisparities=128 MODE_SGBM with uniqueness, sub-pixel, "
                "LR check, median, speckle, int16 disparity out"),
    "cfg3": Cfg("cfg3", 1920, 1080, 0, 0, False, True,
                "cfg3: guided upscale 1920x1080 uint16 depth -> 3840x2160 uint16 (4K RGB guide, r=8, eps=1e-3)"),
    "cfg4": Cfg("cfg4", 1920, 1080, 128, 0, True, True,
                "cfg2+cfg3 fused: full-SBS 3840x1080 (1920x1080/eye) SGBM numDisparities=128 MODE_SGBM "
                "with uniqueness, sub-pixel, LR check, median, speckle -> min-max uint16 -> guided upscale "
                "to 3840x2160 uint16 (4K RGB guide, r=8, eps=1e-3)"),
    "cfg5": Cfg("cfg5", 1920, 1080, 256, 1, True, True,
                "cfg5 stress: full-SBS 3840x1080 (1920x1080/eye) SGBM numDisparities=256 MODE_HH (8 paths), speckle filter "
                "-> min-max uint16 -> guided upscale to 3840x2160 uint16 (4K RGB guide, r=8, eps=1e-3)"),
}


def workload_config(cfg, n_gpus):
    """Identical for both arms and every lane/batch choice (the driver compares it verbatim)."""
    return {
        "workload": cfg.workload,
        "name": cfg.name,
        "parallelism": f"frame-range shards x{n_gpus}, no collective",
        "l2": "inputs and (for the SGBM configurations) the per-frame cost volumes exceed the 126 MB L2; "
              "every lane cycles >= 15 distinct frames",
    }


def synthetic_frames(cfg, n, seed=SEED, t0=0):
    """n distinct seeded frames (SURVEY 8d generator): dict of numpy arrays."""
    from concurrent.futures import ThreadPoolExecutor
    from video_3d_pipeline import synthetic
    out = {}
    with ThreadPoolExecutor(max_workers=min(8, host_cores())) as ex:
        if cfg.depth:
            out["sbs"] = np.stack(list(ex.map(lambda t: synthetic.sbs_frame(seed, t0 + t, cfg.eye_w, cfg.eye_h, cfg.D), range(n))))
        else:
            out["depth"] = np.stack(list(ex.map(lambda t: synthetic.depth_u16(seed, t0 + t, cfg.eye_w, cfg.eye_h), range(n))))
        if cfg.upscale:
            out["guide"] = np.stack(list(ex.map(lambda t: synthetic.guide_frame(seed, t0 + t, cfg.gw, cfg.gh), range(n))))
    return out


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One worker: the reference's cv2 chain on `n` frames (+ the CPU guided-upscale port)."""
    cfg_name, seed, t0, n = args
    cfg = CONFIGS[cfg_name]
    import cv2
    cv2.setNumThreads(1)
    from oracle import cv2_chain, guided
    fr = synthetic_frames(cfg, n, seed, t0)
    matcher = cv2_chain.make_matcher(cfg.D, cfg.mode) if cfg.depth else None
    t = time.perf_counter()
    td = 0.0
    for i in range(n):
        if cfg.depth:
            a = time.perf_counter()
            depth = cv2_chain.depth_from_sbs(fr["sbs"][i], matcher, unsqueeze=False)
            u16 = cv2_chain.normalize_u16(depth)
            td += time.perf_counter() - a
        else:
            u16 = fr["depth"][i]
        if cfg.upscale:
            guided.guided_upscale_cv2(u16, fr["guide"][i], RADIUS, EPS)
    return time.perf_counter() - t, td


def cpu_reference(cfg, frames_per_worker, workers, seed=SEED):
    """Frame-sharded pool over the host cores.  Returns (fps, depth_only_fps or None, wall seconds)."""
    import multiprocessing as mp
    jobs = [(cfg.name, seed, k * frames_per_worker, frames_per_worker) for k in range(workers)]
    if workers == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(workers) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = max(r[0] for r in res)
    depth_wall = max(r[1] for r in res)
    total = frames_per_worker * workers
    return total / wall, (total / depth_wall if depth_wall > 0 else None), wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_text(cfg, workers, fpw, wall=None):
    import cv2
    parts = [f"{workers} worker processes x {fpw} frame per step" + (f" (wall {wall:.1f}s)" if wall else "") +
             ", cv2.setNumThreads(1) each"]
    if cfg.depth:
        parts.append(f"depth = the reference's cv2 {cv2.__version__} call chain "
                     "(depth.py:257-266,274-275,315-325,337-341,374,400-403)")
    if cfg.upscale:
        parts.append("upscale = CPU port of the guided filter (cv2.boxFilter fp32; the reference has none: "
                     "upscale.py:47-59 is ffmpeg scale)")
    return "; ".join(parts)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    cores = host_cores()
    workers = max(1, min(cores, 64))
    fpw = 1
    # warm-up pass (W) then K timed passes, each a bounded sample of `workers` frames
    for _ in range(min(args.warmup, 1)):
        cpu_reference(cfg, fpw, workers)
    vals, dvals, t_total = [], [], 0.0
    for _ in range(max(1, args.steps)):
        fps, dfps, wall = cpu_reference(cfg, fpw, workers)
        vals.append(fps); dvals.append(dfps); t_total += wall
        if t_total > 150:
            break
    value = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * workers * fpw / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/f32",
        "data": "synthetic", "config": workload_config(cfg, args.gpus),
        "frames_per_step": workers * fpw,
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": workers,
            "kind": "reference" if cfg.depth else "port",
            "sample": cpu_sample_text(cfg, workers, fpw),
            "depth_only_fps": statistics.median(dvals) if cfg.depth else None,
            "reference_authored_fraction_of_time": (value / statistics.median(dvals)) if cfg.depth else 0.0,
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------- measured traffic (ncu)
def measured_traffic(cfg_name):
    """DRAM bytes per frame and kernel from the committed ncu capture (tools/ncu_traffic.py writes the file)."""
    f = ROOT / "profiles" / "traffic_per_frame.json"
    if not f.exists():
        return None
    try:
        return json.loads(f.read_text()).get(cfg_name)
    except Exception:
        return None


# stage timer -> kernel it brackets (one kernel per stage timer)
STAGE_KERNEL = {
    "prefilter": "k_prefilter_expand", "cost": "k_cost", "vertical": "k_path_vert3", "lr": "k_path_lr_ckpt",
    "wta": "k_path_rl_wta_tma", "guided_coeff": "k_guided_coeff_s", "guided_apply": "k_guided_apply_s",
}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from video_3d_pipeline import _native as nv

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # Host placement: pinned staging buffers on the GPU's NUMA node; with several ranks per box also keep the
    # rank's (launch-and-copy only) host threads on that node's cores.  V3D_NUMA_BIND=0 switches both off.
    from video_3d_pipeline import shard
    numa = {"node": shard.prefer_gpu_numa_memory(local),
            "cpus_bound": len(shard.bind_to_gpu_numa(local)) if world > 1 else 0}

    params = nv.SgbmParams(numDisparities=cfg.D or 64, mode=cfg.mode)
    W, H, GW, GH = cfg.eye_w, cfg.eye_h, cfg.gw, cfg.gh
    n_lanes = max(1, args.lanes)
    if args.batch > 0:
        B = args.batch
        if B % n_lanes:
            raise ValueError("--batch must be a multiple of --lanes")
        Bl = B // n_lanes
    else:
        # The fused vertical sweep keeps one thread-block cluster per frame co-resident; how many fit is a
        # property of the individual GPU (GPC yield).  Size each lane's batch to exactly one such wave.
        Bl = 15
        with nv.Context(W, H, params, max_batch=1, device=local) as probe:
            if cfg.depth:
                z = torch.zeros((1, H, 2 * W, 3), dtype=torch.uint8, device=dev)
                probe.depth_frames(z, False, want=())
                torch.cuda.synchronize(dev)
                Bl = probe.fused_sweep_clusters or 15
            per_frame = probe.workspace_bytes
        # lanes that fit: workspace + guided coefficients + device-resident and staged copies of the inputs / outputs
        per_frame += (16 * GW * GH if cfg.upscale else 0) + 3 * (cfg.h2d + cfg.d2h)
        free_b, _total_b = torch.cuda.mem_get_info(dev)
        n_lanes = max(1, min(n_lanes, int((free_b - 8e9) // (Bl * per_frame))))
        if world > 1:   # every rank must do the same amount of work (weak scaling): agree on the minimum
            t = torch.tensor([Bl, n_lanes], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            Bl, n_lanes = int(t[0].item()), int(t[1].item())
        B = Bl * n_lanes
    # Bl + n_lanes - 1 distinct seeded frames per rank; lane i takes frames [i, i + Bl): every lane cycles Bl distinct
    # frames and no two lanes hold the same set (speckle / uniqueness / LR work depends on the content)
    pool = synthetic_frames(cfg, Bl + n_lanes - 1, SEED, t0=rank * 1000)

    class Lane:
        """One stream + context + buffers; lanes overlap each other's tails, copies and kernels."""
        def __init__(self, i):
            self.stream = torch.cuda.Stream(dev)
            self.ctx = nv.Context(W, H, params, max_batch=Bl, device=local)
            self.h, self.d = {}, {}
            for k, a in pool.items():
                if k == "depth":
                    t = torch.from_numpy(np.ascontiguousarray(a[i:i + Bl]).view(np.int16)).view(torch.uint16)
                else:
                    t = torch.from_numpy(a[i:i + Bl])
                self.h[k] = t.pin_memory()
                self.d[k] = self.h[k].to(dev)
            # two host output buffers per lane: a context keeps two host calls in flight (submit k+1, wait for k)
            if cfg.upscale:
                self.out_hs = [torch.empty((Bl, GH, GW), dtype=torch.uint16).pin_memory() for _ in range(2)]
                self.out_d = torch.empty((Bl, GH, GW), dtype=torch.uint16, device=dev)
            self.disp_hs = [torch.empty((Bl, H, W), dtype=torch.int16).pin_memory() for _ in range(2)] if cfg.depth else None
            self.calls = 0

        # device-resident step
        def step_device(self, depth_only=False):
            with torch.cuda.stream(self.stream):
                if not cfg.depth:
                    self.ctx.guided_upscale(self.d["depth"], self.d["guide"], RADIUS, EPS, out=self.out_d)
                elif cfg.upscale and not depth_only:
                    self.ctx.depth_frames(self.d["sbs"], False, self.d["guide"], RADIUS, EPS, want=())
                else:
                    self.ctx.depth_frames(self.d["sbs"], False, want=())

        # end-to-end step through the host entry point: submit now, wait later
        def submit_host(self, depth_only=False, copy_only=False):
            if self.ctx.host_pending >= args.inflight:
                self.ctx.host_wait_oldest()
            self.out_h = self.out_hs[self.calls & 1] if cfg.upscale else None
            self.disp_h = self.disp_hs[self.calls & 1] if cfg.depth else None
            self.calls += 1
            with torch.cuda.stream(self.stream):
                if copy_only:
                    out = {"out4k": self.out_h} if cfg.upscale else {"disp": self.disp_h}
                    self.ctx.host_copy_only(self.h.get("sbs"), self.h.get("guide"), out)
                    if not cfg.depth:   # cfg3 also uploads the depth maps: same bytes as the disparity download, other direction
                        self.d["depth"].copy_(self.h["depth"], non_blocking=True)
                elif not cfg.depth:
                    self.ctx.guided_upscale_host(self.h["depth"], self.h["guide"], self.out_h, RADIUS, EPS, wait=False)
                elif cfg.upscale and not depth_only:
                    self.ctx.depth_frames_host(self.h["sbs"], False, self.h["guide"], RADIUS, EPS, out={"out4k": self.out_h}, wait=False)
                else:
                    self.ctx.depth_frames_host(self.h["sbs"], False, out={"disp": self.disp_h}, wait=False)

        def wait_host(self):
            self.ctx.host_wait()

    lanes = [Lane(i) for i in range(n_lanes)]
    ctx = lanes[0].ctx

    def launch_count():
        return sum(l.ctx.launch_count for l in lanes)

    def timed(kind, steps, warmup, **kw):
        """K steps of every lane between two events on the current stream; lanes fork from / join into it.
        Everything is submitted by THIS thread: the device-resident calls are asynchronous launches, the host calls
        are v3d_*_host_async submissions and the thread only ever sleeps in v3d_host_wait (blocking-sync event) on
        the lane it is about to resubmit (each lane keeps two calls in flight: the uploads of call k+1 run under the
        kernels of call k), while the other lanes' work is already queued."""
        def all_lanes(n):
            if kind == "device":
                for _ in range(n):
                    for lane in lanes:
                        lane.step_device(**kw)
            else:
                for _ in range(n):
                    for lane in lanes:
                        lane.submit_host(**kw)       # sleeps for the lane's oldest call when two are in flight
                for lane in lanes:
                    lane.wait_host()

        all_lanes(warmup)
        barrier()
        l0 = launch_count()
        cur = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for lane in lanes:
            lane.stream.wait_event(e0)
        all_lanes(steps)
        for lane in lanes:
            ev = torch.cuda.Event()
            ev.record(lane.stream)
            cur.wait_event(ev)
        e1.record(cur)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = launch_count() - l0
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    def fps(ms, steps):
        return world * B * steps / (ms / 1000.0)

    reps = max(1, args.reps)
    sampler = ClockSampler(local) if rank == 0 else None
    dev_runs = [timed("device", args.steps, args.warmup if i == 0 else 1) for i in range(reps)]
    clocks = sampler.stop() if sampler else None
    dev_ms = sorted(r[0] for r in dev_runs)
    ms = statistics.median(dev_ms)
    launches = dev_runs[0][1]
    value = fps(ms, args.steps)

    e2e_runs = [timed("host", args.steps, max(args.warmup, 3) if i == 0 else 1)[0] for i in range(reps)]
    ms_e2e = statistics.median(e2e_runs)
    e2e = fps(ms_e2e, args.steps)
    copy_ms, _ = timed("host", args.steps, 1, copy_only=True)
    copy_ceiling = fps(copy_ms, args.steps)

    depth_only = None
    if cfg.depth and cfg.upscale and not args.no_depth_only:
        d_ms, _ = timed("device", args.steps, 1, depth_only=True)
        dh_ms, _ = timed("host", args.steps, 1, depth_only=True)
        depth_only = {"value": fps(d_ms, args.steps), "unit": UNIT,
                      "e2e": {"value": fps(dh_ms, args.steps), "h2d_bytes_per_step": B * cfg.sbs_bytes,
                              "d2h_bytes_per_step": B * 2 * W * H},
                      "what": "the SGBM chain alone (split+gray ... speckle, int16 disparity out): the part of the step "
                              "the reference implements itself (cv2.StereoSGBM); compare with cpu_baseline.depth_only_fps"}

    # per-stage breakdown (CUDA events on the launching stream) of ONE lane running alone, untimed pass
    ctx.set_timing(True)
    ctx.reset_timing()
    prof_steps = max(1, min(args.steps, 3))
    for _ in range(prof_steps):
        lanes[0].step_device()
    torch.cuda.synchronize(dev)
    stages = {k: v / prof_steps for k, v in ctx.stage_ms().items()}
    ctx.set_timing(False)
    workspace_gb = sum(l.ctx.workspace_bytes for l in lanes) / 1e9
    # measured integer add/min rate of this GPU (register-only probe kernel, ~1 s): the ALU roofline's denominator
    int_probe = nv.probe_int_throughput(local) if rank == 0 and cfg.depth else None
    fused_clusters = ctx.fused_sweep_clusters

    line = None
    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        per_gpu_fps = value / world
        # One stage timer = one kernel launch per step.  The dominant kernel is the largest of them.
        kstages = {k: v for k, v in stages.items() if k in STAGE_KERNEL and v > 0}
        dom = max(kstages, key=kstages.get)
        dom_ms = kstages[dom]
        # algorithmic bytes of the stage the kernel belongs to: the upscale kernels see depth + guide in, uint16 out
        # (cfg3's figure), the SGBM kernels the depth stage's (SBS in, int16 out)
        alg_pf = CONFIGS["cfg3"].alg_bytes if dom.startswith("guided") else (cfg.sbs_bytes + 2 * W * H)
        alg_bytes = alg_pf * Bl
        achieved = alg_bytes / (dom_ms / 1000.0) / 1e9
        traffic = measured_traffic(cfg.name)
        tk = (traffic or {}).get("kernels", {})
        dom_traffic = tk.get(STAGE_KERNEL[dom])
        total_pf = sum(tk.values()) if tk else None
        roofline = {
            "bound": "hbm", "kernel": STAGE_KERNEL[dom],
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": dom_traffic * Bl if dom_traffic else None, "peak_source": peak_src,
            "traffic_source": (traffic or {}).get("source"),
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms, "frames_per_launch": Bl,
            # the whole step against HBM: what the kernels of one frame actually move (ncu dram__bytes) x frames/s
            "traffic_total_per_frame": total_pf,
            "traffic_total_frac_of_hbm": (total_pf * per_gpu_fps / 1e9 / hbm_peak) if total_pf else None,
            "traffic_over_algorithmic": (total_pf / cfg.alg_bytes) if total_pf else None,
            "per_kernel": {STAGE_KERNEL[k]: {"ms_per_launch": v,
                                             "dram_gbs": (tk[STAGE_KERNEL[k]] * Bl / (v / 1000.0) / 1e9) if STAGE_KERNEL[k] in tk else None}
                           for k, v in kstages.items()},
            "whole_step_hbm_frac_algorithmic": cfg.alg_bytes * per_gpu_fps / 1e9 / hbm_peak,
        }
        if int_probe:
            pk_i = max(v[1] for v in int_probe.values())
            roofline["alu"] = {"algorithmic_iops_per_frame": cfg.alg_iops, "achieved_tiops": cfg.alg_iops * per_gpu_fps / 1e12,
                               "nominal_peak_tiops": 148 * 128 * 1.965e9 / 1e12,
                               "measured": {k: {"lane_instr_tps": v[0] / 1e12, "algorithmic_tiops": v[1] / 1e12}
                                            for k, v in int_probe.items()},
                               "measured_peak_tiops": pk_i / 1e12,
                               "frac_of_measured_peak": cfg.alg_iops * per_gpu_fps / pk_i}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            shard.reset_numa_memory_policy()    # the CPU workers allocate wherever they run
            workers = max(1, min(host_cores(), 64))
            cfps, dfps, wall = cpu_reference(cfg, 1, workers)
            cpu = {"value": cfps, "unit": UNIT, "cores": workers, "kind": "reference" if cfg.depth else "port",
                   "sample": cpu_sample_text(cfg, workers, 1, wall), "depth_only_fps": dfps}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16/f32" if cfg.depth else "f32", "data": "synthetic",
            "config": workload_config(cfg, world),
            "frames_per_step": world * B, "frames_per_step_per_gpu": B, "lanes": n_lanes,
            "repetitions": {"n": reps, "device_fps": [fps(m, args.steps) for m in dev_ms][::-1],
                            "e2e_fps": sorted(fps(m, args.steps) for m in e2e_runs),
                            "value_is": "median", "min_device_fps": fps(max(dev_ms), args.steps)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * cfg.h2d,
                    "d2h_bytes_per_step": B * cfg.d2h, "ms_per_step": ms_e2e / args.steps,
                    "host_copy_ceiling": copy_ceiling, "frac_of_host_copy_ceiling": e2e / copy_ceiling,
                    "host_gbs": {"h2d": world * B * cfg.h2d * args.steps / (ms_e2e / 1000.0) / 1e9,
                                 "d2h": world * B * cfg.d2h * args.steps / (ms_e2e / 1000.0) / 1e9},
                    "submit": "one host thread per rank, two calls in flight per lane; v3d_*_host_async + v3d_host_wait_oldest (blocking-sync event), "
                              "per-frame copies on separate upload / download streams"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "depth_only": depth_only,
            "stages_ms_per_step": stages,
            "workspace_gb": workspace_gb,
            "fused_sweep_clusters": fused_clusters,
            "host_numa": numa,
        }
    for lane in lanes:
        lane.ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (default cfg4 = cfg2 + cfg3 fused, the headline)")
    ap.add_argument("--reps", type=int, default=3, help="repetitions of the K timed steps; value = the median")
    ap.add_argument("--batch", type=int, default=0,
                    help="frames per step per GPU (default: lanes x the co-resident clusters of the fused sweep, 90 on most B200s)")
    ap.add_argument("--lanes", type=int, default=6, help="streams/contexts the batch is split over")
    ap.add_argument("--inflight", type=int, default=2, choices=[1, 2],
                    help="end-to-end host calls in flight per lane (2: uploads of call k+1 under the kernels of call k)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-depth-only", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # plain `python bench.py --gpus N`: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
