#!/bin/bash
# round 2: ncu --set full of the packed half-ring kernel on BASELINE config 4 (N=256, 1e5 points)
python tools/ab_kernels.py --workload c4 --rows 100 --reps 2 --only 0,8192,8448,8704 > gpurun_out/r2_ab_c4_mirror.txt 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_mirror_packed -s 1 -c 1 -f -o gpurun_out/r2_c4_mirror_packed python tools/ab_kernels.py --workload c4 --rows 100 --reps 1 --only 8192 > gpurun_out/ncu_c4_mirror.log 2>&1
cat gpurun_out/r2_ab_c4_mirror.txt; tail -2 gpurun_out/ncu_c4_mirror.log
