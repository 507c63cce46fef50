set -x
ncu --set full --clock-control none --kernel-name regex:aggregate -c 6 -o gpurun_out/r02s_agg python tools/stage_bench.py --batch 64 --once > gpurun_out/ncu_agg.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"sp_(costs|relax_exact)" --launch-skip 14 --launch-count 2 -o gpurun_out/r02s_sp python tools/stage_bench.py --batch 16 --once > gpurun_out/ncu_sp.log 2>&1
ncu --set full --clock-control none --kernel-name regex:"wta_walk" -c 1 -o gpurun_out/r02s_wta python tools/stage_bench.py --batch 64 --once > gpurun_out/ncu_wta.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02s_launches.csv python bench.py --steps 1 --warmup 0 --frames 1000 --extra 0 --cpu-sample 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
python tools/stage_bench.py --batch 64 > gpurun_out/r02s_stage_times_b64.txt 2>&1; cat gpurun_out/r02s_stage_times_b64.txt
