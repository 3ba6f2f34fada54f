#!/bin/bash
# run ablation binaries of tools/hot_probe.cu (built into tools/hp_bin/) on the GPU box
for b in ${2:-tools/hp_bin/*}; do timeout 25 $b ${1:-16} || echo "$b: timeout / error"; done 2>&1 | tee gpurun_out/hp_run.txt
