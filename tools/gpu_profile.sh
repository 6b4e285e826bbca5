# End-of-round evidence run on a GPU box: tests, smoke, default bench, reference arm, ncu launch list of the
# default command, ncu --set full of the six largest kernels, the other BASELINE configs.  Outputs land in
# gpurun_out/ and are summarised into profiles/ by hand.   gpurun -- bash tools/gpu_profile.sh
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; tail -c 400 gpurun_out/r01_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_ref.json 2>/dev/null; cat gpurun_out/r01_ref.json | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches_default_cmd.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --batch 15 --lanes 1 --no-cpu-baseline > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_cost|k_path_vert3|k_path_lr_tma|k_path_rl_wta_tma|k_guided_coeff_s|k_guided_apply_s" -s 6 -c 6 -o gpurun_out/r01_top6 -f python bench.py --steps 1 --warmup 1 --batch 15 --lanes 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python tools/bench_configs.py cfg1 rp_default cfg2 cfg2g cfg5s cfg5 2>/dev/null | cut -c1-200
ls -la gpurun_out
