# Evidence run on a GPU box: default bench, reference arm, the other BASELINE configs, ncu launch list of the
# default command, ncu --set full of the largest kernels.  Outputs land in gpurun_out/ and are summarised into
# profiles/ (tools/ncu_traffic.py turns the capture into profiles/traffic_per_frame.json).
#   gpurun -- bash tools/gpu_profile.sh [tag] [configs...]      (tools/ab.sh: A/B of library variants on one box)
#   CFG_FLAGS="--no-cpu-baseline" shortens the per-configuration lines (the CPU arm does not depend on the build)
TAG=${1:-r02}; shift
CFGS=${@:-cfg1 cfg2 cfg3 cfg5}
K='k_cost|k_path_vert3|k_path_lr_ckpt|k_path_rl_wta_tma|k_guided_coeff_s|k_guided_apply_s|k_prefilter_expand'
set -x
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_ref.json
for c in $CFGS; do
  timeout 600 python bench.py --config $c $CFG_FLAGS > gpurun_out/${TAG}_${c}.json 2> gpurun_out/${TAG}_${c}.err; cut -c1-200 gpurun_out/${TAG}_${c}.json
done
timeout 300 python bench.py --steps 2 --warmup 1 --reps 1 --no-cpu-baseline --no-depth-only > /dev/null 2>&1 && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_default_cmd.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-cpu-baseline --no-depth-only > /dev/null 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --reps 1 --batch 15 --lanes 1 --no-cpu-baseline --no-depth-only > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 7 -c 7 -o gpurun_out/${TAG}_top7 -f python bench.py --steps 1 --warmup 1 --reps 1 --batch 15 --lanes 1 --no-cpu-baseline --no-depth-only > /dev/null 2>&1
ls -la gpurun_out | tail -12
