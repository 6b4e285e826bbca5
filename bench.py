#!/usr/bin/env python
"""Headline benchmark: 1080p-SBS -> 4K depth frames/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one pass of the depth hot path over one batch of B synthetic frames per GPU:
full-SBS 3840x1080 BGR -> split + gray -> cv2-exact SGBM (D=128, 5 paths, uniqueness, sub-pixel,
LR check, 3x3 median, speckle) -> /16, clamp, per-frame min-max -> uint16 -> guided upscale to
3840x2160 uint16 with a 4K RGB guide (r=8, eps=1e-3).  (BASELINE.json configs[1] + configs[2],
i.e. configs[3]'s per-GPU work.)

Printed JSON line (rank 0):
  value    frames/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e      the same through the host-buffer C-ABI call (v3d_depth_frames_host): pinned host
           inputs, H2D + D2H copies inside the timed region
  roofline dominant kernel family against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's cv2 chain (+ CPU guided upscale port) on the host cores, bounded sample

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
EYE_W, EYE_H, D = 1920, 1080, 128
GW, GH = 3840, 2160
RADIUS, EPS = 8, 1e-3
SBS_BYTES = 2 * EYE_W * EYE_H * 3            # 12 441 600
GUIDE_BYTES = GW * GH * 3                    # 24 883 200
OUT_BYTES = GW * GH * 2                      # 16 588 800
ALG_BYTES_DEPTH = SBS_BYTES + 2 * EYE_W * EYE_H          # 16 588 800
ALG_BYTES_FUSED = SBS_BYTES + GUIDE_BYTES + OUT_BYTES    # 53 913 600
W1 = EYE_W - D
CELLS = W1 * EYE_H * D
ALG_IOPS = (31 + 9 * 5) * CELLS                          # 18.8 Gop
# DRAM bytes per frame measured by `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum of a
# 15-frame launch / 15), see profiles/README.md; keyed by bench stage.
NCU_DRAM_BYTES_PER_FRAME = {      # profiles/r01d_ncu_full_top6.csv
    "cost": (1.978912e9 + 7.377555e9) / 15,
    "vertical": (14.863622e9 + 7.381563e9) / 15,
    "lr": (7.432424e9 + 7.389744e9) / 15,
    "wta": (14.864874e9 + 0.236329e9) / 15,
    "guided": (0.452217e9 + 1.937375e9 + 2.407501e9 + 0.245115e9) / 2 / 15,
}


def workload_config(batch, n_gpus, lanes=1):
    return {
        "workload": "cfg2+cfg3 fused: full-SBS 3840x1080 (1920x1080/eye) SGBM numDisparities=128 MODE_SGBM "
                    "with uniqueness, sub-pixel, LR check, median, speckle -> min-max uint16 -> guided upscale "
                    "to 3840x2160 uint16 (4K RGB guide, r=8, eps=1e-3)",
        "frames_per_step_per_gpu": batch,
        "lanes": lanes,
        "global_frames_per_step": batch * n_gpus,
        "parallelism": f"frame-range shards x{n_gpus}, no collective",
        "l2": "inputs (37.3 MB/frame) and the per-frame C/S volumes (0.99 GB/frame) exceed the 126 MB L2",
    }


def synthetic_batch(batch, n_distinct=2, seed=11):
    from video_3d_pipeline import synthetic
    sbs = [synthetic.sbs_frame(seed, t, EYE_W, EYE_H, D) for t in range(n_distinct)]
    guide = [synthetic.guide_frame(seed, t, GW, GH) for t in range(n_distinct)]
    s = np.stack([sbs[i % n_distinct] for i in range(batch)])
    g = np.stack([guide[i % n_distinct] for i in range(batch)])
    return s, g


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One worker: the reference's cv2 chain on `n` frames (+ the CPU guided-upscale port)."""
    seed, t0, n, with_upscale = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import cv2_chain, guided
    from video_3d_pipeline import synthetic
    matcher = cv2_chain.make_matcher(D, 0)
    frames = [synthetic.sbs_frame(seed, t0 + i, EYE_W, EYE_H, D) for i in range(n)]
    guides = [synthetic.guide_frame(seed, t0 + i, GW, GH) for i in range(n)] if with_upscale else None
    t = time.perf_counter()
    td = 0.0
    for i in range(n):
        a = time.perf_counter()
        depth = cv2_chain.depth_from_sbs(frames[i], matcher, unsqueeze=False)
        u16 = cv2_chain.normalize_u16(depth)
        td += time.perf_counter() - a
        if with_upscale:
            guided.guided_upscale_cv2(u16, guides[i], RADIUS, EPS)
    return time.perf_counter() - t, td


def cpu_reference(frames_per_worker, workers, with_upscale=True, seed=11):
    """Frame-sharded pool over the host cores.  Returns (fps, depth_only_fps, wall seconds)."""
    import multiprocessing as mp
    jobs = [(seed, k * frames_per_worker, frames_per_worker, with_upscale) for k in range(workers)]
    if workers == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(workers) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = max(r[0] for r in res)
    depth_wall = max(r[1] for r in res)
    total = frames_per_worker * workers
    return total / wall, total / depth_wall, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import cv2
    cores = host_cores()
    workers = max(1, min(cores, 64))
    fpw = 1
    # warm-up pass (W) then K timed passes, each a bounded sample of `workers` frames
    for _ in range(min(args.warmup, 1)):
        cpu_reference(fpw, workers)
    vals, dvals, t_total = [], [], 0.0
    for _ in range(max(1, args.steps)):
        fps, dfps, wall = cpu_reference(fpw, workers)
        vals.append(fps); dvals.append(dfps); t_total += wall
        if t_total > 150:
            break
    value = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * workers * fpw / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/f32",
        "data": "synthetic", "config": workload_config(workers * fpw, 1),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": workers, "kind": "reference",
            "sample": f"{workers} worker processes x {fpw} frame per step, cv2.setNumThreads(1) each; depth = the "
                      f"reference's cv2 {cv2.__version__} call chain (depth.py:257-266,274-275,315-325,337-341,374,400-403); "
                      "upscale = CPU port of the guided filter (cv2.boxFilter fp32; the reference has none)",
            "depth_only_fps": statistics.median(dvals),
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


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from video_3d_pipeline import _native as nv

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

    params = nv.SgbmParams(numDisparities=D, mode=nv.MODE_SGBM)
    n_lanes = max(1, args.lanes)
    if args.batch > 0:
        B = args.batch
        if B % n_lanes:
            raise ValueError("--batch must be a multiple of --lanes")
        Bl = B // n_lanes
    else:
        # The fused vertical sweep keeps one thread-block cluster per frame co-resident; how many fit is a
        # property of the individual GPU (GPC yield).  Size each lane's batch to exactly one such wave.
        with nv.Context(EYE_W, EYE_H, params, max_batch=1, device=local) as probe:
            z = torch.zeros((1, EYE_H, 2 * EYE_W, 3), dtype=torch.uint8, device=dev)
            probe.depth_frames(z, False, want=())
            torch.cuda.synchronize(dev)
            Bl = probe.fused_sweep_clusters or 15
        # lanes that fit: a lane holds ~1.45 GB per frame (cost volumes, guided coefficients, in/out buffers)
        free_b, _total_b = torch.cuda.mem_get_info(dev)
        n_lanes = max(1, min(n_lanes, int((free_b - 6e9) // (Bl * 1.45e9))))
        if world > 1:   # every rank must do the same amount of work (weak scaling): agree on the minimum
            t = torch.tensor([Bl, n_lanes], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            Bl, n_lanes = int(t[0].item()), int(t[1].item())
        B = Bl * n_lanes
    sbs_np, guide_np = synthetic_batch(Bl)

    class Lane:
        """One stream + context + buffers; lanes overlap each other's tails, copies and kernels."""
        def __init__(self):
            self.stream = torch.cuda.Stream(dev)
            self.ctx = nv.Context(EYE_W, EYE_H, params, max_batch=Bl, device=local)
            self.sbs_d = torch.from_numpy(sbs_np).to(dev)
            self.guide_d = torch.from_numpy(guide_np).to(dev)
            self.sbs_h = torch.from_numpy(sbs_np).pin_memory()
            self.guide_h = torch.from_numpy(guide_np).pin_memory()
            self.out_h = torch.empty((Bl, GH, GW), dtype=torch.uint16).pin_memory()

        def step_device(self):
            with torch.cuda.stream(self.stream):
                self.ctx.depth_frames(self.sbs_d, False, self.guide_d, RADIUS, EPS, want=())

        def step_host(self):
            with torch.cuda.stream(self.stream):
                self.ctx.depth_frames_host(self.sbs_h, False, self.guide_h, RADIUS, EPS, out={"out4k": self.out_h})

    lanes = [Lane() for _ in range(n_lanes)]
    ctx = lanes[0].ctx

    def launch_count():
        return sum(l.ctx.launch_count for l in lanes)

    def timed(kind, steps, warmup):
        """K steps of every lane between two events on the default stream; lanes fork from / join into it."""
        import threading

        def run(lane, n):
            torch.cuda.set_device(local)
            for _ in range(n):
                (lane.step_device if kind == "device" else lane.step_host)()

        def all_lanes(n):
            if kind == "device" or n_lanes == 1:
                for _ in range(n):
                    for lane in lanes:
                        (lane.step_device if kind == "device" else lane.step_host)()
            else:   # the host entry point is synchronous: one host thread per lane (ctypes drops the GIL)
                th = [threading.Thread(target=run, args=(lane, n)) for lane in lanes]
                for t in th:
                    t.start()
                for t in th:
                    t.join()

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

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = timed("device", args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms / 1000.0)

    ms_e2e, _ = timed("host", args.steps, max(args.warmup, 3) if args.warmup else 0)
    e2e = world * B * args.steps / (ms_e2e / 1000.0)

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
    int_probe = nv.probe_int_throughput(local) if rank == 0 else None
    fused_clusters = ctx.fused_sweep_clusters

    line = None
    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # One stage timer = one kernel (guided = two).  The dominant kernel is the largest of them.
        vol = 2.0 * CELLS                 # one uint16 cost volume, bytes per frame
        kernels = {   # stage -> (kernel name, launches in the stage, design bytes per frame)
            "cost": ("k_cost", 1, vol + 0.15e9),
            "vertical": ("k_path_vert3", 1, 3 * vol),      # C read + S read-modify-write (L2 reductions)
            "lr": ("k_path_lr_tma", 1, 2 * vol),
            "wta": ("k_path_rl_wta_tma", 1, 2 * vol),
            "guided": ("k_guided_coeff_s + k_guided_apply_s (avg of 2)", 2, (GUIDE_BYTES + 16 * GW * GH) / 2 + GW * GH * 9.0),
        }
        dom_name = max(kernels, key=lambda k: stages.get(k, 0.0) / kernels[k][1])
        kname, nl, design_pf = kernels[dom_name]
        dom_ms = stages[dom_name] / nl
        alg_bytes = (ALG_BYTES_FUSED if dom_name == "guided" else ALG_BYTES_DEPTH) * Bl
        achieved = alg_bytes / (dom_ms / 1000.0) / 1e9
        design_bytes = design_pf * Bl
        ncu_pf = NCU_DRAM_BYTES_PER_FRAME.get(dom_name)
        roofline = {
            "bound": "hbm", "kernel": kname,
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": ncu_pf * Bl if ncu_pf else None, "peak_source": peak_src,
            "traffic_source": "profiles/ (ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, per frame x frames per launch)",
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
            "design_bytes_per_launch": design_bytes,
            "design_gbs": design_bytes / (dom_ms / 1000.0) / 1e9,
            "design_frac": design_bytes / (dom_ms / 1000.0) / 1e9 / hbm_peak,
            "per_kernel": {k: {"kernel": v[0], "ms_per_launch": stages.get(k, 0.0) / v[1],
                               "design_gbs": v[2] * Bl / max(stages.get(k, 1e-9) / v[1] / 1000.0, 1e-12) / 1e9}
                           for k, v in kernels.items()},
            "alu": {"algorithmic_iops_per_frame": ALG_IOPS, "achieved_tiops": ALG_IOPS * (value / world) / 1e12,
                    "nominal_peak_tiops": 148 * 128 * 1.965e9 / 1e12,
                    "measured": {k: {"lane_instr_tps": v[0] / 1e12, "algorithmic_tiops": v[1] / 1e12}
                                 for k, v in int_probe.items()},
                    "measured_peak_tiops": max(v[1] for v in int_probe.values()) / 1e12,
                    "frac_of_measured_peak": ALG_IOPS * (value / world) / max(v[1] for v in int_probe.values())},
            "whole_step_hbm_frac_algorithmic": ALG_BYTES_FUSED * (value / world) / 1e9 / hbm_peak,
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import cv2
            shard.reset_numa_memory_policy()    # the CPU workers allocate wherever they run
            workers = max(1, min(host_cores(), 64))
            fps, dfps, wall = cpu_reference(1, workers)
            cpu = {"value": fps, "unit": UNIT, "cores": workers, "kind": "reference",
                   "sample": f"{workers} worker processes x 1 frame (wall {wall:.1f}s), cv2.setNumThreads(1) each; depth = "
                             f"the reference's cv2 {cv2.__version__} call chain; upscale = CPU port of the guided filter "
                             "(the reference has none)",
                   "depth_only_fps": dfps}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16/f32", "data": "synthetic", "config": workload_config(B, world, n_lanes),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * (SBS_BYTES + GUIDE_BYTES),
                    "d2h_bytes_per_step": B * OUT_BYTES, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "stages_ms_per_step": stages,
            "workspace_gb": workspace_gb,
            "fused_sweep_clusters": fused_clusters,
            "lanes": n_lanes,
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
    ap.add_argument("--batch", type=int, default=0,
                    help="frames per step per GPU (default: lanes x the co-resident clusters of the fused sweep, 90 on most B200s)")
    ap.add_argument("--lanes", type=int, default=6, help="streams/contexts the batch is split over")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
