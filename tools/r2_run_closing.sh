#!/bin/bash
# round 2, closing state of the library: GPU suite, smoke, bench lines of every config, config table
# (the A/B records of the last changes: profiles/r2_ab_complete_split_rule.txt,
#  r2_ab_mirror_packed.txt, r2_ab_mirror_survey.txt, r2_ab_cycle_large_mirror.txt)
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests_closing.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_closing.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke_closing.txt
bash tools/r2_run_final.sh > gpurun_out/r2_run_final.log 2>&1
cat gpurun_out/r2_gpu_tests_closing.txt; tail -3 gpurun_out/r2_smoke_closing.txt; head -12 gpurun_out/r2_run_final.log
