#!/bin/bash
# compile-time parameter sweep on the GPU box: scripts/sweep.sh "<flags1>" "<flags2>" ...
for flags in "$@"; do
  SPHB_NVCC_EXTRA="$flags" python python-sph_b200/build_native.py --force > /dev/null 2>&1 || { echo "BUILD FAILED $flags"; continue; }
  python bench.py --nx 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sweep_tmp.log 2>&1
  python - "$flags" <<'PY'
import json, sys
try:
    d=json.loads(open("gpurun_out/sweep_tmp.log").read().strip().splitlines()[-1])
    k=d["kernels"]
    print(f"{sys.argv[1]:60s} step {d['ms_per_step']:7.2f}  force {k['force']['ms']/2:6.2f} newton {k['newton_h']['ms']/2:6.2f} omega {k['density_omega']['ms']/2:6.2f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/sweep_tmp.log").read()[-500:])
PY
done
python python-sph_b200/build_native.py --force > /dev/null 2>&1
