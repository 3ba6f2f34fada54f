#!/bin/bash
# ncu --set full of the bulk-copy streaming kernel, KB = 8 (default) and KB = 1 (one pass each)
python tools/run_stream.py 8 1024 8 > gpurun_out/plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_stream_tma -c 1 -f -o gpurun_out/r2_stream_tma_kb8_final python tools/run_stream.py 8 1024 8 > gpurun_out/ncu_kb8.log 2>&1
python tools/run_stream.py 1 1024 8 r8 > gpurun_out/plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_stream_tma -s 2 -c 1 -f -o gpurun_out/r2_stream_tma_kb1_final python tools/run_stream.py 1 1024 8 r8 > gpurun_out/ncu_kb1.log 2>&1
cat gpurun_out/plain8.log gpurun_out/plain1.log; tail -1 gpurun_out/ncu_kb8.log gpurun_out/ncu_kb1.log
