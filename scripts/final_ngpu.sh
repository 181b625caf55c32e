#!/bin/bash
# round-end measurement batch on N GPUs of one box: bash scripts/final_ngpu.sh N [tests]
N=$1
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
grep -v "^\[W\|^$\|frame #\|^\*\*\*\|OMP_NUM" gpurun_out/r02_bench_${N}gpu.err | tail -3
timeout 900 $TR --master-port 29562 bench.py --gpus $N --workload plummer --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_bench_plummer64M_${N}gpu.json 2> gpurun_out/r02_bench_plummer64M_${N}gpu.err
grep -v "^\[W\|^$\|frame #\|^\*\*\*\|OMP_NUM" gpurun_out/r02_bench_plummer64M_${N}gpu.err | tail -3
if [ "$2" = "tests" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "range or gravity or (lists and args0)" > gpurun_out/r02_multi_tests_${N}gpu.log 2>&1
  tail -3 gpurun_out/r02_multi_tests_${N}gpu.log
fi
