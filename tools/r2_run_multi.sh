#!/bin/bash
# round 2, two-GPU call: NCCL parallel_routine check, weak and strong scaling bench lines
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tests/run_parallel_routine.py 2>&1 | grep -v "^W0\|^\*\*\*\*\|Setting OMP" | tee gpurun_out/r2_parallel_routine_n$N.txt
timeout 300 $TR --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_n$N.err
timeout 300 $TR --master-port 29535 bench.py --gpus $N --steps 2 --warmup 3 --scaling strong > gpurun_out/r2_bench_c5_strong_n$N.json 2>> gpurun_out/r2_bench_n$N.err
timeout 200 $TR --master-port 29536 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2>> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
for f in ["gpurun_out/r2_bench_c5_n$N.json", "gpurun_out/r2_bench_c5_strong_n$N.json", "gpurun_out/r2_bench_ref_n$N.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["scaling"], d["e2e"], d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
