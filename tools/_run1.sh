bash tools/gpu_check.sh r2a
V3D_NO_SIDE_STREAM=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-depth-only --reps 1 > gpurun_out/b_r2a_noside.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/b_r2a_noside.log').read().strip().splitlines()[-1]);print('noside value',d['value'],'e2e',d['e2e']['value'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-depth-only --reps 1 --lanes 1 > gpurun_out/b_r2a_1lane.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/b_r2a_1lane.log').read().strip().splitlines()[-1]);print('1lane value',d['value'],'e2e',d['e2e']['value'])"
ls /usr/lib/x86_64-linux-gnu/ | grep -i -E "nvcuvid|nvidia-encode|libcuda" ; ldconfig -p | grep -i -E "nvcuvid|nvenc|nvidia-encode"; nproc; lscpu | grep -E "Model name|Socket|NUMA"
