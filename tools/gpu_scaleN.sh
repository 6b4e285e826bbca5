# N-GPU line of the headline configuration (N = 2 or 4):   gpurun --gpus N -- bash tools/gpu_scaleN.sh N [tag]
N=$1; TAG=${2:-r02}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-depth-only > gpurun_out/${TAG}_cfg4_${N}gpu.json 2> gpurun_out/${TAG}_cfg4_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_cfg4_${N}gpu.json").read().strip().splitlines()[-1])
print("N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ceiling", round(d["e2e"]["host_copy_ceiling"],1), d["e2e"]["host_gbs"])
PY
