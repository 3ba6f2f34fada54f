#!/bin/bash
# nested cycle kernel at several ring sizes (block shapes of 1 .. 8 warps, ragged sizes) and the
# rings above one block (cluster / streaming); QW_B200_LIB=<other build> for the A/B partner
for n in 512 1001 1024 2048 3000 4096; do python tools/ab_kernels.py --workload c5 --n $n --rows 16 --reps 2 --only 0 | sed "s/^/N=$n /"; done
python tools/ab_large_n.py --topology circular --sizes 5000,8192 --flags 0 --points 2368
