#!/bin/bash
# round 2: smoke, launch list of the default bench command, ncu --set full of the timed launch
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
python bench.py --steps 2 --warmup 3 --cpu-seconds 3 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_c5_default.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 3 > gpurun_out/ncu_launches.log 2>&1
grep -c "rk4_" gpurun_out/r2_launches_c5_default.csv
python tools/ab_kernels.py --workload c5 --rows 128 --reps 1 --only 0 > gpurun_out/plain_ab.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_horner -s 1 -c 1 -f -o gpurun_out/r2_cycle_horner_bench_share python tools/ab_kernels.py --workload c5 --rows 128 --reps 1 --only 0 > gpurun_out/ncu_hot.log 2>&1
tail -1 gpurun_out/ncu_hot.log
