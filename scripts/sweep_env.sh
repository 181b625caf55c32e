#!/bin/bash
# runtime-parameter sweep: scripts/sweep_env.sh "VAR=val ..." ...
for envs in "$@"; do
  env $envs python bench.py --nx 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sweep_tmp.log 2>&1
  python - "$envs" <<'PY'
import json, sys
try:
    d=json.loads(open("gpurun_out/sweep_tmp.log").read().strip().splitlines()[-1])
    k=d["kernels"]
    print(f"{sys.argv[1]:40s} step {d['ms_per_step']:7.2f}  force {k['force']['ms']/2:6.2f} newton {k['newton_h']['ms']/2:6.2f} omega {k['density_omega']['ms']/2:6.2f} build {k['grid_build']['ms']/3:5.2f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/sweep_tmp.log").read()[-300:])
PY
done
