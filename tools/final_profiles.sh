# Profiles of the current code, run on the GPU box after the bench and the tests have exited 0 without ncu:
#   bash tools/final_profiles.sh r02w
set -x
T=${1:-rXX}
ncu --set full --clock-control none --kernel-name regex:aggregate -c 6 -o gpurun_out/${T}_agg python tools/stage_bench.py --batch 64 --once > gpurun_out/ncu_agg.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"sp_(costs|relax_exact)" --launch-skip 14 --launch-count 2 -o gpurun_out/${T}_sp python tools/stage_bench.py --batch 16 --once > gpurun_out/ncu_sp.log 2>&1
ncu --set full --clock-control none --kernel-name regex:"wta_walk" -c 1 -o gpurun_out/${T}_wta python tools/stage_bench.py --batch 64 --once > gpurun_out/ncu_wta.log 2>&1
# launch list: the first 1600 launches (five 64-frame batches) of the default 1000-frame sequence
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 0 --frames 1000 --extra 0 --cpu-sample 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
python tools/stage_bench.py --batch 64 > gpurun_out/${T}_stage_times_b64.txt 2>&1; cat gpurun_out/${T}_stage_times_b64.txt
