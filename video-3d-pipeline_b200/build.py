"""Builds libv3d.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

    python video-3d-pipeline_b200/build.py [--force] [--verbose]

Every translation unit is compiled to its own object (in parallel, only when it or a header changed) and
the objects are linked into video_3d_pipeline/libv3d.so, next to the Python package, so that the library
travels with the source tree; objects and library are git-ignored.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
OUT = Path(os.environ.get("V3D_LIB_OUT") or (HERE / "video_3d_pipeline" / "libv3d.so"))   # V3D_LIB_OUT: build a variant elsewhere
SOURCES = ["v3d_api.cu", "k_gray.cu", "k_cost.cu", "k_paths.cu", "k_paths_h.cu", "k_post.cu", "k_guided.cu", "k_png.cu",
           "k_probe.cu"]
HEADERS = [CSRC / "v3d_internal.h", CSRC / "path_common.cuh", CSRC / "tma.cuh", HERE.parent / "include" / "v3d.h"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def _flags():
    return NVCC_FLAGS + os.environ.get("V3D_NVCC_EXTRA", "").split()      # e.g. -DV3D_COST_HN=4 for tuning experiments


def _obj(src, tag):
    return OBJ / f"{Path(src).stem}.{tag}.o"


def build(force=False, verbose=False):
    flags = _flags() + (["-Xptxas", "-v"] if verbose else [])
    tag = hashlib.sha1(" ".join(flags).encode()).hexdigest()[:8]           # different flags -> different objects
    sources = [s for s in SOURCES if (CSRC / s).exists()]
    hdr_t = max(h.stat().st_mtime for h in HEADERS + [Path(__file__)])
    OBJ.mkdir(exist_ok=True)
    todo = []
    for s in sources:
        o = _obj(s, tag)
        if force or not o.exists() or o.stat().st_mtime < max(hdr_t, (CSRC / s).stat().st_mtime):
            todo.append(s)

    def compile_one(s):
        cmd = ["nvcc"] + flags + ["-c", "-o", str(_obj(s, tag)), str(CSRC / s)]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, todo))
    objs = [_obj(s, tag) for s in sources]
    if todo or not OUT.exists() or OUT.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        cmd = ["nvcc", "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(OUT)] + [str(o) for o in objs] + ["-ldl"]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
