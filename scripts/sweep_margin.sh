#!/bin/bash
# compile-time sweep of the Newton list margin: scripts/sweep_margin.sh 2e-3 4e-3 ...
for m in "$@"; do
  SPHB_NVCC_EXTRA="-DSPHB_NEWTON_MARGIN=$m" python -c "
import sys; sys.path.insert(0,'python-sph_b200'); import build_native; build_native.build(force=True)" > /dev/null 2>&1
  echo "margin $m"
  python scripts/prof_simloop.py 2>&1 | grep "newton calls:\|ms/step"
done
