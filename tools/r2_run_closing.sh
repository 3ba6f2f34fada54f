#!/bin/bash
# round 2, closing state of the library: split-rule A/B, GPU suite, bench lines, config table
python tools/ab_large_n.py --topology complete --sizes 4097,5000,6144,7000,8192,9000,10000,12289,16384,40000 --flags 0,128 --points 2368 > gpurun_out/r2_ab_complete_split_rule.txt 2>&1
python tools/ab_large_n.py --topology circular --sizes 5000,8192,12289,16384 --flags 0,8192 --points 2368 > gpurun_out/r2_ab_cycle_large_mirror.txt 2>&1
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests_closing.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_closing.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke_closing.txt
bash tools/r2_run_final.sh > gpurun_out/r2_run_final.log 2>&1
cat gpurun_out/r2_ab_complete_split_rule.txt gpurun_out/r2_ab_cycle_large_mirror.txt gpurun_out/r2_gpu_tests_closing.txt; tail -3 gpurun_out/r2_smoke_closing.txt; head -12 gpurun_out/r2_run_final.log
