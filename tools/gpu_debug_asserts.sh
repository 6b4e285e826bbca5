# Parity tests against a DEBUG build of libv3d.so (-DV3D_DEBUG_ASSERTS=1: index / protocol checks in the fused vertical
# sweep's inbox ring and the speckle filter's union-find trap on violation, mbarrier waits time out instead of hanging).
# Build here (nvcc cross-compiles), run on the GPU box:
#   V3D_LIB_OUT=$PWD/variants/debug.so V3D_NVCC_EXTRA=-DV3D_DEBUG_ASSERTS=1 python video-3d-pipeline_b200/build.py
#   gpurun -- bash tools/gpu_debug_asserts.sh
export V3D_LIB=$PWD/variants/debug.so
test -f $V3D_LIB || { echo "build variants/debug.so first"; exit 1; }
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu -k "not guided and not png" 2>&1 | grep -E "V3D_DASSERT|passed|failed|Error" | sort | uniq -c | tail -8
