"""Frame-range sharding across the GPUs of one box.

Frames are independent -- the matcher is stateless per call (depth.py:315-341), normalisation is
per frame (depth.py:400-401) and every frame is its own file (depth.py:466) -- so N GPUs are N
independent workers over contiguous frame ranges with a host-side gather by file index.  There is
no collective on the data path and therefore no NCCL here.
"""
import sys
from pathlib import Path
from typing import List, Tuple

import torch


def frame_ranges(first: int, count: int, parts: int) -> List[Tuple[int, int]]:
    """Split [first, first+count) into `parts` contiguous (start, n) ranges, sizes differing by <= 1.

    Empty ranges are kept (n = 0) so that rank k always owns entry k.
    """
    if parts <= 0:
        raise ValueError("parts must be positive")
    base, extra = divmod(max(count, 0), parts)
    out, s = [], first
    for k in range(parts):
        n = base + (1 if k < extra else 0)
        out.append((s, n))
        s += n
    return out


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (sysfs cpulist format) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def _gpu_sysfs(gpu_index: int) -> Path:
    p = torch.cuda.get_device_properties(gpu_index)
    return Path(f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0")


def gpu_local_cpus(gpu_index: int) -> List[int]:
    """CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs `local_cpulist`); [] when unknown."""
    try:
        return parse_cpulist((_gpu_sysfs(gpu_index) / "local_cpulist").read_text())
    except Exception:
        return []


def gpu_numa_node(gpu_index: int) -> int:
    """NUMA node of the GPU (sysfs `numa_node`); -1 when unknown or the box has a single node."""
    try:
        return int((_gpu_sysfs(gpu_index) / "numa_node").read_text())
    except Exception:
        return -1


def prefer_gpu_numa_memory(gpu_index: int) -> int:
    """Make the GPU's NUMA node the preferred node for this process's future allocations.

    With several workers per box every frame crosses the host twice (37 MB in, 17 MB out at the headline
    shape).  Pinned staging buffers live where they are first touched; a worker whose buffers sit on the other
    socket sends all of that over the inter-socket link.  `set_mempolicy(MPOL_PREFERRED, node)` keeps the
    threads free to run anywhere (decoding wants every core) and still falls back to other nodes when the
    preferred one is full.  Call before allocating pinned memory.  Returns the node, or -1 when nothing was
    done (unknown topology, not Linux x86-64/aarch64, or V3D_NUMA_BIND=0).
    """
    import ctypes
    import os
    import platform
    if os.environ.get("V3D_NUMA_BIND", "1") == "0":
        return -1
    node = gpu_numa_node(gpu_index)
    nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
    if node < 0 or node >= 1024 or nr is None or not sys.platform.startswith("linux"):
        return -1
    MPOL_PREFERRED = 1
    mask = (ctypes.c_ulong * 16)()                      # 1024 node bits
    bits = 8 * ctypes.sizeof(ctypes.c_ulong)
    mask[node // bits] = 1 << (node % bits)
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(MPOL_PREFERRED), ctypes.byref(mask), ctypes.c_ulong(1024))
    return node if rc == 0 else -1


def reset_numa_memory_policy() -> None:
    """Back to the default (local-node) allocation policy, e.g. before forking CPU-only helpers."""
    import ctypes
    import platform
    nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
    if nr is None or not sys.platform.startswith("linux"):
        return
    ctypes.CDLL(None).syscall(ctypes.c_long(nr), ctypes.c_int(0), None, ctypes.c_ulong(0))


def bind_to_gpu_numa(gpu_index: int, min_cpus: int = 4) -> List[int]:
    """Pin the calling process to the allowed CPUs that are local to `gpu_index` (for workers whose host side
    only launches and copies, like bench.py's ranks; decode-bound workers use prefer_gpu_numa_memory alone).

    Does nothing (returns []) when the topology is unknown, when fewer than `min_cpus` local CPUs are allowed,
    or with V3D_NUMA_BIND=0.
    """
    import os
    if os.environ.get("V3D_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return []
    local = set(gpu_local_cpus(gpu_index)) & set(os.sched_getaffinity(0))
    if len(local) < min_cpus:
        return []
    try:
        os.sched_setaffinity(0, local)
    except OSError:
        return []
    return sorted(local)


def _worker(rank: int, cfg: dict, ret):
    # imported here: the child must not inherit an initialised CUDA context
    from .depth import HybridStereoDepthExtractor
    prefer_gpu_numa_memory(rank)
    ex = HybridStereoDepthExtractor(
        model_checkpoint=cfg["model_checkpoint"], work_dir=cfg["work_dir"], cache_dir=cfg["cache_dir"],
        device="cuda", batch_size=cfg["batch_size"], use_neural_guidance=False, stereo_only=True,
        unsqueeze_sbs=cfg["unsqueeze_sbs"], num_disparities=cfg["num_disparities"], sgbm_mode=cfg["sgbm_mode"],
        num_gpus=1, gpu_index=rank, decode_threads=cfg["decode_threads"], png_threads=cfg["png_threads"],
        png_compression=cfg["png_compression"], depth_scale=cfg["depth_scale"], gpu_lanes=cfg["gpu_lanes"])
    ex.load_model()
    start, n = cfg["ranges"][rank]
    done = 0
    if n > 0:
        # global file index = offset of this range inside the whole job (depth.py:464-466 numbering)
        done = ex._process_range(cfg["video_path"], start, n, Path(cfg["cache_path"]), start - cfg["first"])
    ret[rank] = done


def run_sharded(extractor, video_path: str, first: int, count: int, cache_path: Path) -> int:
    """Run `extractor`'s frame range on extractor.num_gpus GPUs; returns frames written."""
    n_gpus = min(extractor.num_gpus, torch.cuda.device_count())
    if n_gpus < extractor.num_gpus:
        print(f"Only {n_gpus} GPUs visible; sharding over those", file=sys.stderr)
    ranges = frame_ranges(first, count, n_gpus)
    cfg = dict(model_checkpoint=extractor.model_checkpoint, work_dir=str(extractor.work_dir),
               cache_dir=str(extractor.cache_dir), batch_size=extractor.batch_size,
               unsqueeze_sbs=extractor.unsqueeze_sbs, num_disparities=extractor.num_disparities,
               sgbm_mode=extractor.sgbm_mode, decode_threads=extractor.decode_threads,
               png_threads=extractor.png_threads, png_compression=extractor.png_compression,
               depth_scale=extractor.depth_scale, gpu_lanes=extractor.gpu_lanes, ranges=ranges, first=first, video_path=video_path,
               cache_path=str(cache_path))
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, cfg, ret)) for r in range(n_gpus)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        bad = [r for r, p in enumerate(procs) if p.exitcode != 0]
        if bad:
            raise RuntimeError(f"depth shard worker(s) {bad} failed; rerun those frame ranges (outputs are idempotent)")
        return int(sum(ret.get(r, 0) for r in range(n_gpus)))
