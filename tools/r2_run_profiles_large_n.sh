#!/bin/bash
# round 2: ncu --set full of the two kernels above one thread block's registers
# (complete graph split over blocks, new block rule; cycle graph across a cluster), N = 8192
python tools/ab_large_n.py --topology complete --sizes 8192 --flags 0 --points 2368 > gpurun_out/plain_split.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rk4_complete_folded -s 1 -c 1 -f -o gpurun_out/r2_complete_split_n8192 python tools/ab_large_n.py --topology complete --sizes 8192 --flags 0 --points 2368 > gpurun_out/ncu_split.log 2>&1
tail -2 gpurun_out/ncu_split.log
python tools/ab_large_n.py --topology circular --sizes 8192 --flags 0 --points 2368 > gpurun_out/plain_cluster.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_cluster -s 1 -c 1 -f -o gpurun_out/r2_cycle_cluster_n8192 python tools/ab_large_n.py --topology circular --sizes 8192 --flags 0 --points 2368 > gpurun_out/ncu_cluster.log 2>&1
tail -2 gpurun_out/ncu_cluster.log
