# Round-1 closing capture (one GPU): full GPU test suite, smoke(), default bench, C3 bench,
# config table, ncu --set full of the complete-graph kernels.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err || exit 1
cut -c1-300 gpurun_out/bench_final3.json
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-stream > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
cut -c1-300 gpurun_out/bench_c3_final.json
timeout 600 python tests/run_configs.py > gpurun_out/configs11.md 2> gpurun_out/configs11.err; tail -3 gpurun_out/configs11.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk4_complete_folded -s 1 -c 1 -o gpurun_out/complete_folded_n1024 python tools/ab_kernels.py --workload c3 --rows 8 --n 1024 --reps 1 --only 0 > gpurun_out/ncu_complete_folded.log 2>&1
tail -1 gpurun_out/ncu_complete_folded.log
