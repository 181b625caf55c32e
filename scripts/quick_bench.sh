#!/bin/bash
# usage: quick_bench.sh tag [bench args]  -- runs gpu tests (brief) and the Sod bench, prints the kernel split
tag=$1; shift
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_$tag.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    print("$tag ms/step", round(d["ms_per_step"],2), {k:round(v["ms"],2) for k,v in d["kernels"].items() if v["ms"]>1})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.log").read()[-2000:])
PY
