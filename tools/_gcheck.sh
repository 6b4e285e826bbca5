python -m pytest tests -m gpu -x -q -k "guided or upscale or smoke or module" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_stream.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/b_stream.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stages_ms_per_step"]["guided"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_guided" -c 4 --csv --log-file gpurun_out/g_launch.csv python bench.py --steps 1 --warmup 1 --batch 15 --lanes 1 --no-cpu-baseline > /dev/null 2>&1
tail -2 gpurun_out/g_launch.csv | awk -F'","' '{print substr($5,1,40), $NF}'
