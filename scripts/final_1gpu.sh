#!/bin/bash
# round-end measurement batch on ONE B200 (run under gpurun): plain bench runs first, profiler passes after them
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
tail -2 gpurun_out/r02_bench_1gpu.err
timeout 300 python bench.py --steps 10 --warmup 3 --warm-state-steps 60 --dt-scale 4 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_1gpu_developed.json 2> gpurun_out/r02_bench_1gpu_developed.err
timeout 900 python bench.py --workload plummer --steps 5 --warmup 3 --no-e2e --cpu-seconds 10 > gpurun_out/r02_bench_plummer64M_1gpu.json 2> gpurun_out/r02_bench_plummer64M_1gpu.err
tail -3 gpurun_out/r02_bench_plummer64M_1gpu.err
# profiler passes (numbers printed under ncu are never bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --nx 192 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_nl_(force|newton_h|build|density_omega)' --launch-skip 24 -c 10 \
    -o gpurun_out/r02_final -f python bench.py --nx 192 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_final.ncu-rep
