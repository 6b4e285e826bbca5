# End-of-round evidence run on a GPU box: tests, smoke, default bench, reference arm, ncu launch list of the
# default command, ncu --set full of the six largest kernels.  Outputs land in gpurun_out/ and are summarised into
# profiles/ by hand.   gpurun -- bash tools/gpu_profile.sh [tag]      (tools/bench_configs.py covers the other configs)
TAG=${1:-r01}
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_ref.json
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_default_cmd.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --batch 15 --lanes 1 --no-cpu-baseline > /dev/null 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_cost|k_path_vert3|k_path_lr_tma|k_path_rl_wta_tma|k_guided_coeff_s|k_guided_apply_s" -s 6 -c 6 -o gpurun_out/${TAG}_top6 -f python bench.py --steps 1 --warmup 1 --batch 15 --lanes 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
