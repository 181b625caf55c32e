#!/bin/bash
# compile-time sweep: scripts/sweep_def.sh "-DA=1 -DB=2" "-DA=2" ...  (prints per-kernel times of plain / carry steps)
for defs in "$@"; do
  SPHB_NVCC_EXTRA="$defs" python -c "
import sys; sys.path.insert(0,'python-sph_b200'); import build_native; build_native.build(force=True)" > /dev/null 2>&1
  echo "== $defs"
  python scripts/prof_simloop.py 2>&1 | grep "newton calls:\|ms/step\|force  \|density_omega"
done
