# A/B of kernel variants on one box: tools/ab.sh [--config cfgN] <lib.so|default> ...   (bench stage times)
CFG=cfg4; if [ "$1" = "--config" ]; then CFG=$2; shift 2; fi
for v in "$@"; do
  if [ "$v" = default ]; then unset V3D_LIB; else export V3D_LIB=$PWD/$v; fi
  timeout 300 python bench.py --config $CFG --steps 5 --warmup 2 --reps 1 --no-cpu-baseline --no-depth-only > gpurun_out/ab_$(basename $v .so).json 2> gpurun_out/ab_$(basename $v .so).err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$(basename $v .so).json").read().strip().splitlines()[-1])
    print("$v".ljust(28), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k:round(x,3) for k,x in d["stages_ms_per_step"].items() if x>0.2})
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/ab_$(basename $v .so).err").read()[-800:])
PY
done
