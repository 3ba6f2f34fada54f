#!/bin/bash
# round 2 closing measurements on one B200
python bench.py > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err
python bench.py --scaling strong --steps 3 --warmup 3 --no-stream --no-cpu-baseline > gpurun_out/r2_bench_c5_strong_n1.json 2>> gpurun_out/r2_bench_c5_n1.err
for w in c2 c3 c4; do python bench.py --workload $w --no-stream --cpu-seconds 5 > gpurun_out/r2_bench_${w}_n1.json 2>> gpurun_out/r2_bench_c5_n1.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n1.json 2>> gpurun_out/r2_bench_c5_n1.err
python tests/run_configs.py > gpurun_out/r2_configs.md 2> gpurun_out/r2_configs.err
tail -n 3 gpurun_out/r2_bench_c5_n1.err; tail -n 3 gpurun_out/r2_configs.err
python - <<PY
import json
for f in ["r2_bench_c5_n1", "r2_bench_c5_strong_n1", "r2_bench_c2_n1", "r2_bench_c3_n1", "r2_bench_c4_n1", "r2_bench_ref_n1"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.4e" % d["value"], "%.1f ms" % d["ms_per_step"], d.get("roofline", {}).get("frac"), d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
