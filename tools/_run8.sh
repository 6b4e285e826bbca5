nvidia-smi topo -m > gpurun_out/r02a_topo8.txt 2>&1; nproc >> gpurun_out/r02a_topo8.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r02a_topo8.txt; free -g | head -2 >> gpurun_out/r02a_topo8.txt
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@" > gpurun_out/r02a_n8_${tag}.json 2> gpurun_out/r02a_n8_${tag}.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02a_n8_${tag}.json").read().strip().splitlines()[-1])
    print("${tag}", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ceiling", round(d["e2e"]["host_copy_ceiling"],1), d["e2e"]["host_gbs"], d["repetitions"]["e2e_fps"])
except Exception as e:
    print("${tag} failed", e); print(open("gpurun_out/r02a_n8_${tag}.err").read()[-1500:])
PY
}
run cfg4 --no-depth-only
run cfg4_3lanes --no-depth-only --lanes 3
run cfg5 --config cfg5 --no-depth-only
