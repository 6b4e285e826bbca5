# End-to-end number with one or two host calls in flight per lane, for several lane counts (one box):
#   gpurun -- bash tools/ab_inflight.sh
for lanes in 2 6; do for inf in 1 2; do
  timeout 300 python bench.py --steps 12 --warmup 3 --reps 1 --lanes $lanes --inflight $inf --no-cpu-baseline --no-depth-only > gpurun_out/inf_${lanes}_${inf}.json 2> gpurun_out/inf_${lanes}_${inf}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/inf_${lanes}_${inf}.json").read().strip().splitlines()[-1])
print("lanes", $lanes, "inflight", $inf, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ceiling", round(d["e2e"]["host_copy_ceiling"],1))
PY
done; done
