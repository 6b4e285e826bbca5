# GPU-box check used during development: full GPU test suite, then a short default bench with stage times.
#   gpurun -- bash tools/gpu_check.sh
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_cur.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/b_cur.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], {k:round(v,3) for k,v in d["stages_ms_per_step"].items()})
PY
