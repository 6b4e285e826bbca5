"""CPU-side checks: the C-ABI library loads and exports every symbol include/v3d.h declares, the
error paths that need no GPU behave, and the host-side logic (module surface, cache keys, frame
sharding incl. a world_size-2 gloo run) is right."""
import ctypes as C
import hashlib
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "v3d.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(v3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from video_3d_pipeline import _native
    lib = _native.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libv3d.so does not export {n}"
    assert b"sm_100a" in lib.v3d_version()


def test_params_struct_matches_header_and_reference_literals():
    from video_3d_pipeline import _native
    p = _native.SgbmParams()
    _native.lib().v3d_default_params(C.byref(p))
    # depth.py:315-325
    assert (p.minDisparity, p.numDisparities, p.blockSize, p.P1, p.P2) == (0, 64, 5, 600, 2400)
    assert (p.disp12MaxDiff, p.uniquenessRatio, p.speckleWindowSize, p.speckleRange, p.mode) == (1, 10, 100, 32, 0)
    assert C.sizeof(p) == 44


def test_create_validates_before_touching_cuda():
    from video_3d_pipeline import _native
    L = _native.lib()
    h = C.c_void_p()
    bad = [
        (_native.SgbmParams(numDisparities=64), 66, 20),        # W - D <= blockSize/2 : cv2.error site
        (_native.SgbmParams(numDisparities=40), 400, 20),       # not a multiple of 16 (cv2 rejects it too)
        (_native.SgbmParams(numDisparities=272), 600, 20),      # beyond the largest kernel instantiation
        (_native.SgbmParams(minDisparity=2000), 4000, 20),      # (minDisparity + numDisparities) * 16 must fit int16
        (_native.SgbmParams(minDisparity=20), 86, 20),          # window [84, 86) not wider than blockSize/2: cv2.error site
        (_native.SgbmParams(blockSize=4), 400, 20),
        (_native.SgbmParams(mode=2), 400, 20),
        (_native.SgbmParams(P2=60000), 400, 20),                # packed 16-bit state would overflow
    ]
    for p, w, hh in bad:
        assert L.v3d_create(0, C.byref(p), w, hh, 1, C.byref(h)) == _native.V3D_EINVAL
        assert L.v3d_last_error()
    assert L.v3d_create(0, None, 400, 20, 1, C.byref(h)) == _native.V3D_EINVAL


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from video_3d_pipeline import _native
    h = C.c_void_p()
    p = _native.SgbmParams()
    assert _native.lib().v3d_create(0, C.byref(p), 400, 20, 1, C.byref(h)) == _native.V3D_ECUDA
    assert b"no CPU fallback" in _native.lib().v3d_last_error()
    with pytest.raises(RuntimeError):
        _native.Context(400, 20)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.probe_int_throughput(0)
    a = C.c_double()
    assert _native.lib().v3d_probe_int_throughput(0, 7, C.byref(a), C.byref(a)) == _native.V3D_EINVAL
    from video_3d_pipeline.depth import IGEVStereoDepthExtractor
    with pytest.raises(RuntimeError, match="CUDA not available"):
        IGEVStereoDepthExtractor(work_dir="/tmp/v3d_t", cache_dir="/tmp/v3d_t")


def test_product_never_imports_the_oracle():
    pkg = ROOT / "video-3d-pipeline_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")):
        text = f.read_text()
        assert "oracle" not in re.sub(r"#.*|//.*|\"\"\".*?\"\"\"", "", text, flags=re.S).replace("oracle/guided.py", ""), f
    assert not re.search(r"^\s*(from|import)\s+oracle", (pkg / "video_3d_pipeline" / "_native.py").read_text(), re.M)


def test_module_surface_matches_reference():
    import video_3d_pipeline as v
    from video_3d_pipeline import depth, upscale, utils
    # reference __init__.py:5-16 + the name run_pipeline.py:12 imports
    for n in ("VideoAligner", "IGEVStereoDepthExtractor", "SimpleDepthUpscaler", "get_video_info", "extract_audio",
              "verify_video_compatibility"):
        assert hasattr(v, n)
    assert depth.IGEVStereoDepthExtractor is depth.HybridStereoDepthExtractor
    import inspect
    sig = inspect.signature(depth.HybridStereoDepthExtractor.__init__)
    want = ["model_checkpoint", "work_dir", "cache_dir", "device", "batch_size", "use_neural_guidance", "stereo_only",
            "unsqueeze_sbs"]                                             # depth.py:23-31, same order
    assert list(sig.parameters)[1:1 + len(want)] == want
    assert sig.parameters["batch_size"].default == 8 and sig.parameters["unsqueeze_sbs"].default is True
    for m in ("load_model", "get_cache_path", "is_cached", "extract_frames_opencv", "extract_frames_ffmpeg",
              "split_sbs_frame", "preprocess_frame_pair", "process_frame_batch", "save_depth_map", "process_video_sbs"):
        assert callable(getattr(depth.HybridStereoDepthExtractor, m))
    s = inspect.signature(depth.HybridStereoDepthExtractor.process_video_sbs)
    assert list(s.parameters)[1:] == ["video_path", "start_frame", "max_frames", "force_reprocess"]
    u = inspect.signature(upscale.SimpleDepthUpscaler.process_depth_upscaling)
    assert list(u.parameters)[1:] == ["depth_dir", "video_4k_path", "output_path", "force_reprocess"]
    assert list(inspect.signature(upscale.SimpleDepthUpscaler.upscale_depth_maps_ffmpeg).parameters)[1:6] == \
        ["depth_dir", "target_width", "target_height", "output_path", "fps"]
    assert callable(depth.main) and callable(upscale.main) and callable(utils.create_work_directory)


def test_cache_key_is_the_references(tmp_path):
    # depth.py:116-125: md5(f"{video}_{start}_{count}_{model}_{unsqueeze}")[:16]
    from video_3d_pipeline.depth import HybridStereoDepthExtractor as E
    ex = E.__new__(E)
    ex.cache_dir = tmp_path
    ex.model_checkpoint = "Intel/dpt-large"
    ex.unsqueeze_sbs = True
    p = ex.get_cache_path("clip.mkv", 5, 100)
    key = hashlib.md5("clip.mkv_5_100_Intel/dpt-large_True".encode()).hexdigest()[:16]
    assert p == tmp_path / f"depth_{key}" and p.is_dir()
    assert not ex.is_cached(p, 2)
    for i in range(2):
        (p / f"depth_{i:06d}.png").write_bytes(b"x")
    assert ex.is_cached(p, 2) and not ex.is_cached(p, 3)
    ex.depth_scale, ex.num_disparities = "fixed", 64     # opt-in output format gets its own cache entry
    assert ex.get_cache_path("clip.mkv", 5, 100) != p
    # every knob that changes the pixels gets its own entry; reference-equivalent settings keep the reference hash
    ex.depth_scale = "frame"
    assert ex.get_cache_path("clip.mkv", 5, 100) == p
    seen = {p}
    for nd, mode, ign in ((128, 0, False), (64, 1, False), (128, 1, False), (64, 0, True)):
        ex.num_disparities, ex.sgbm_mode, ex._guidance_ignored = nd, mode, ign
        q = ex.get_cache_path("clip.mkv", 5, 100)
        assert q not in seen
        seen.add(q)


def test_reader_plan_and_exact_seek_detection(tmp_path):
    """Several decode threads only where a seek is frame-exact (intra-only codecs); slices are whole batches,
    contiguous, and cover the range exactly once."""
    import cv2
    from video_3d_pipeline.depth import HybridStereoDepthExtractor as E
    for first, count, bs, readers in ((0, 100, 8, 4), (10, 37, 8, 4), (3, 5, 8, 4), (0, 64, 8, 3), (0, 1, 1, 9)):
        sl = E.plan_reader_slices(first, count, bs, readers, 1000)
        assert len(sl) <= readers and sum(n for _, n, _ in sl) == count
        at = first
        for s0, n, idx in sl:
            assert s0 == at and idx == 1000 + (s0 - first) and n > 0
            assert (s0 - first) % bs == 0          # every slice starts on a batch boundary
            at += n
    frames = [np.full((32, 64, 3), 20 * i, np.uint8) for i in range(6)]

    def write(path, fourcc):
        vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), 24.0, (64, 32))
        if not vw.isOpened():
            return False
        for f in frames:
            vw.write(f)
        vw.release()
        return path.exists() and path.stat().st_size > 0

    if write(tmp_path / "intra.avi", "MJPG"):
        assert E.seek_is_frame_exact(str(tmp_path / "intra.avi"))
    if write(tmp_path / "gop.mp4", "mp4v"):
        assert not E.seek_is_frame_exact(str(tmp_path / "gop.mp4"))      # long-GOP: one sequential reader
    assert not E.seek_is_frame_exact(str(tmp_path / "missing.mkv"))


def test_video_info_fallback_and_work_dir(tmp_path):
    import cv2
    from video_3d_pipeline.utils import create_work_directory, get_video_info
    assert get_video_info(str(tmp_path / "missing.mp4")) is None
    clip = tmp_path / "c.avi"
    vw = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (64, 32))
    if not vw.isOpened():
        pytest.skip("no MJPG writer")
    for i in range(7):
        vw.write(np.full((32, 64, 3), i * 10, np.uint8))
    vw.release()
    info = get_video_info(str(clip))
    assert info and (info["width"], info["height"], info["frames"]) == (64, 32, 7)
    assert abs(info["fps"] - 24.0) < 1e-3 and abs(info["duration"] - 7 / 24.0) < 1e-3
    d = create_work_directory(str(tmp_path / "wd"))
    assert d.is_dir() and create_work_directory(str(tmp_path / "wd")) == d


def test_frame_ranges_partition():
    from video_3d_pipeline.shard import frame_ranges
    for first, count, parts in ((0, 10000, 8), (10, 23, 4), (0, 3, 8), (7, 0, 2), (0, 1, 1)):
        r = frame_ranges(first, count, parts)
        assert len(r) == parts and sum(n for _, n in r) == count
        assert r[0][0] == first
        for (s0, n0), (s1, _) in zip(r, r[1:]):
            assert s1 == s0 + n0
        sizes = [n for _, n in r]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        frame_ranges(0, 10, 0)


def test_numa_placement_helpers(monkeypatch):
    """Host placement of shard workers: cpulist parsing, no-ops without topology, preferred-node policy."""
    from video_3d_pipeline import shard
    assert shard.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert shard.parse_cpulist("") == [] and shard.parse_cpulist("5") == [5]
    before = os.sched_getaffinity(0)
    # no GPU here: topology unknown -> nothing is changed
    assert shard.gpu_local_cpus(0) == [] and shard.gpu_numa_node(0) == -1
    assert shard.prefer_gpu_numa_memory(0) == -1 and shard.bind_to_gpu_numa(0) == []
    # fewer local CPUs than the floor -> no binding either
    monkeypatch.setattr(shard, "gpu_local_cpus", lambda i: sorted(before)[:1])
    assert shard.bind_to_gpu_numa(0, min_cpus=2) == [] and os.sched_getaffinity(0) == before
    # the switch
    monkeypatch.setattr(shard, "gpu_numa_node", lambda i: 0)
    monkeypatch.setenv("V3D_NUMA_BIND", "0")
    assert shard.prefer_gpu_numa_memory(0) == -1
    monkeypatch.delenv("V3D_NUMA_BIND")
    try:
        assert shard.prefer_gpu_numa_memory(0) in (0, -1)      # -1 where the sandbox forbids set_mempolicy
    finally:
        shard.reset_numa_memory_policy()


_GLOO_SCRIPT = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from video_3d_pipeline.shard import frame_ranges
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
start, n = frame_ranges(100, 37, world)[rank]
# every rank "processes" its range; gather = host side, by global file index
mine = torch.zeros(37, dtype=torch.int64)
mine[start - 100:start - 100 + n] = rank + 1
dist.all_reduce(mine)                      # only the test gathers; the product writes files instead
ms = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing
if rank == 0:
    assert (mine > 0).all() and int((mine == 1).sum()) == 19 and int((mine == 2).sum()) == 18, mine
    assert ms.item() == 11.0
    print("GLOO_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "gloo_shard.py"
    script.write_text(_GLOO_SCRIPT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script),
                          str(ROOT / "video-3d-pipeline_b200")], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_bench_configs_and_reference_arm_schema():
    """bench.py: every BASELINE.json configuration carries SURVEY 8(d)'s algorithmic figures, and both arms print
    the same `config` object for a configuration whatever the lane / batch choice (the driver compares it)."""
    sys.path.insert(0, str(ROOT))
    import bench
    assert sorted(bench.CONFIGS) == ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    c = bench.CONFIGS
    assert c["cfg1"].alg_bytes == 8294400 and c["cfg2"].alg_bytes == 16588800
    assert c["cfg3"].alg_bytes == 49766400 and c["cfg4"].alg_bytes == c["cfg5"].alg_bytes == 53913600
    assert c["cfg1"].alg_iops == 76 * 896 * 1080 * 64 and c["cfg2"].alg_iops == 76 * 1792 * 1080 * 128
    assert c["cfg5"].alg_iops == 103 * 1664 * 1080 * 256
    assert c["cfg4"].h2d == 12441600 + 24883200 and c["cfg4"].d2h == 16588800
    for name, cfg in c.items():
        a, b = bench.workload_config(cfg, 1), bench.workload_config(cfg, 1)
        assert a == b and a["name"] == name and "workload" in a
        assert not any(k in a for k in ("lanes", "frames_per_step_per_gpu", "global_frames_per_step"))


def test_bench_reference_arm_runs_a_small_config(monkeypatch, capsys):
    """`bench.py --impl reference` end to end on a shrunken configuration: one JSON line with the contract keys."""
    import json
    sys.path.insert(0, str(ROOT))
    import bench
    small = bench.Cfg("cfg4", 160, 48, 64, 0, True, True, bench.CONFIGS["cfg4"].workload)
    monkeypatch.setitem(bench.CONFIGS, "cfg4", small)
    monkeypatch.setattr(bench, "host_cores", lambda: 1)
    monkeypatch.setenv("RANK", "0")
    args = type("A", (), dict(config="cfg4", warmup=0, steps=1, gpus=1))()
    assert bench.run_reference_arm(args) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["config"] == bench.workload_config(small, 1)
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_alignment_offset_semantics(tmp_path):
    """utils.py:299-327: the SBS clip is the time reference, the 4K clip is shifted, never below zero."""
    import json
    from video_3d_pipeline.utils import apply_alignment_offset, guide_start_frame
    f = tmp_path / "alignment_data.json"
    f.write_text(json.dumps({"video1_path": "sbs.mkv", "video2_path": "uhd.mkv", "time_offset_seconds": 1.5}))
    assert apply_alignment_offset(str(f), "sbs.mkv", 10.0) == 10.0
    assert apply_alignment_offset(str(f), "uhd.mkv", 10.0) == 11.5
    assert guide_start_frame(str(f), "uhd.mkv", 24.0) == 36
    with pytest.raises(ValueError):
        apply_alignment_offset(str(f), "other.mkv")
    f.write_text(json.dumps({"video1_path": "sbs.mkv", "video2_path": "uhd.mkv", "time_offset_seconds": -2.0}))
    assert apply_alignment_offset(str(f), "uhd.mkv", 0.5) == 0.0
    assert guide_start_frame(str(f), "uhd.mkv", 24.0) == 0


def test_png16_container_writer_without_gpu(tmp_path):
    """_native.png16_file_chunks / write_png16 (the host half of the GPU PNG writer) on a CPU-made payload."""
    import zlib
    import cv2
    from video_3d_pipeline import _native as nv
    img = np.random.default_rng(3).integers(0, 65536, (37, 91), dtype=np.uint16)
    raw = b"".join(b"\x00" + img[y].astype(">u2").tobytes() for y in range(37))
    nv.write_png16(tmp_path / "a.png", zlib.compress(raw, 0), 91, 37)
    dec = cv2.imread(str(tmp_path / "a.png"), cv2.IMREAD_UNCHANGED)
    assert dec.dtype == np.uint16 and np.array_equal(dec, img)
