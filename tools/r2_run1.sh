#!/bin/bash
# round 2, GPU call 1: full-size parity tests + FP64 opcode counters of the hot kernel
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_fullsize.py -x -q --durations=12 ) > gpurun_out/r2_fullsize_tests.log 2>&1
tail -30 gpurun_out/r2_fullsize_tests.log
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum
python tools/ab_kernels.py --workload c5 --rows 16 --reps 1 --only 0 > gpurun_out/plain.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:rk4_cycle_horner -c 1 --csv --log-file gpurun_out/r2_hot_opcounts.csv python tools/ab_kernels.py --workload c5 --rows 16 --reps 1 --only 0 > gpurun_out/ncu1.log 2>&1
tail -5 gpurun_out/r2_hot_opcounts.csv
