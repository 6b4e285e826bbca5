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


def _worker(rank: int, cfg: dict, ret):
    # imported here: the child must not inherit an initialised CUDA context
    from .depth import HybridStereoDepthExtractor
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
