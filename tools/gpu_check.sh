# GPU-box check used during development: full GPU test suite, then a short default bench with stage times.
#   gpurun -- bash tools/gpu_check.sh [tag]
TAG=${1:-cur}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_${TAG}.log 2> gpurun_out/b_${TAG}.err
tail -3 gpurun_out/b_${TAG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b_${TAG}.log").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "copy ceiling", round(d["e2e"]["host_copy_ceiling"],1),
      "depth-only", d["depth_only"] and round(d["depth_only"]["value"],1))
print({k:round(v,3) for k,v in d["stages_ms_per_step"].items()})
print(d["repetitions"])
PY
