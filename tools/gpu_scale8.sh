# 8-GPU evidence (one box, one rank per GPU under torchrun): the headline configuration and the stress configuration,
# each with its end-to-end number and the copy-only ceiling of the box.   gpurun --gpus 8 -- bash tools/gpu_scale8.sh [tag]
TAG=${1:-r02}
nvidia-smi topo -m > gpurun_out/${TAG}_topology_8gpu_box.txt 2>&1; nproc >> gpurun_out/${TAG}_topology_8gpu_box.txt
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/${TAG}_topology_8gpu_box.txt; free -g | head -2 >> gpurun_out/${TAG}_topology_8gpu_box.txt
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@" > gpurun_out/${TAG}_${tag}_8gpu.json 2> gpurun_out/${TAG}_${tag}_8gpu.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${tag}_8gpu.json").read().strip().splitlines()[-1])
    print("${tag}", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ceiling", round(d["e2e"]["host_copy_ceiling"],1), d["e2e"]["host_gbs"])
except Exception as e:
    print("${tag} failed", e); print(open("gpurun_out/${TAG}_${tag}_8gpu.err").read()[-1500:])
PY
}
run cfg4 --no-depth-only
run cfg5 --config cfg5 --no-depth-only
