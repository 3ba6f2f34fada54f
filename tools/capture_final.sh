set -x
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_default.csv python bench.py > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_mirror -c 1 -o gpurun_out/mirror_warp python tools/ab_kernels.py --n 1024 --only 8192 --rows 8 --reps 1 > gpurun_out/ncu_mirror.log 2>&1
tail -3 gpurun_out/ncu_mirror.log
cat gpurun_out/bench_final.json | cut -c1-400
