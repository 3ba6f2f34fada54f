set -x
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk4_cycle_stream -c 1 -o gpurun_out/stream_kb8 python tools/run_stream.py 8 1024 8 > gpurun_out/ncu_kb8.log 2>&1
tail -2 gpurun_out/ncu_kb8.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
