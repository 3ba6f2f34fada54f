#!/bin/bash
# round 2 (closing state): smoke, FP64 opcode counters of the hot kernel, launch list of the default
# bench command, ncu --set full of the launch bench.py times
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum
python tools/ab_kernels.py --workload c5 --rows 16 --reps 1 --only 0 > gpurun_out/plain_ab16.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:rk4_cycle_horner -c 1 --csv --log-file gpurun_out/r2_cycle_horner_opcounts.csv python tools/ab_kernels.py --workload c5 --rows 16 --reps 1 --only 0 > gpurun_out/ncu_op.log 2>&1
tail -4 gpurun_out/r2_cycle_horner_opcounts.csv | cut -c1-50,200-400
python bench.py --steps 2 --warmup 3 --cpu-seconds 3 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_c5_default.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 3 > gpurun_out/ncu_launches.log 2>&1
grep -c "rk4_" gpurun_out/r2_launches_c5_default.csv
python tools/ab_kernels.py --workload c5 --rows 128 --reps 1 --only 0 > gpurun_out/plain_ab.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_horner -s 1 -c 1 -f -o gpurun_out/r2_cycle_horner_bench_share python tools/ab_kernels.py --workload c5 --rows 128 --reps 1 --only 0 > gpurun_out/ncu_hot.log 2>&1
tail -1 gpurun_out/ncu_hot.log; cat gpurun_out/plain_ab.log
