# Round-1 closing capture (one GPU): full GPU test suite, smoke(), default bench, reference arm,
# launch list of the default bench command, ncu --set full of the complete-graph kernel.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err || exit 1
cut -c1-400 gpurun_out/bench_final3.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk4_complete_horner -s 1 -c 1 -o gpurun_out/complete_dadd python tools/ab_kernels.py --workload c3 --rows 8 --reps 1 --only 0 > gpurun_out/ncu_complete_dadd.log 2>&1
tail -2 gpurun_out/ncu_complete_dadd.log
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-stream > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
cut -c1-300 gpurun_out/bench_c3_final.json
