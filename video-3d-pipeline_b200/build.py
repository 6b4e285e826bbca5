"""Builds libv3d.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

    python video-3d-pipeline_b200/build.py [--force] [--verbose]

The .so lands next to the Python package (video_3d_pipeline/libv3d.so) so that it
travels with the source tree; it is git-ignored.
"""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "video_3d_pipeline" / "libv3d.so"
SOURCES = ["v3d_api.cu", "k_gray.cu", "k_cost.cu", "k_paths.cu", "k_paths_h.cu", "k_post.cu", "k_guided.cu", "k_png.cu", "k_probe.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--shared",
]


def _stale():
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "v3d_internal.h", CSRC / "path_common.cuh", CSRC / "tma.cuh", HERE.parent / "include" / "v3d.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    extra = os.environ.get("V3D_NVCC_EXTRA", "").split()      # e.g. -DV3D_COST_RPB=5 for tuning experiments
    cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", str(OUT)] + [str(CSRC / s) for s in SOURCES]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
