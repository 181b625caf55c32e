#!/bin/bash
# compile-time sweep on the GPU box: scripts/sweep_variants.sh <script.py args...> -- "<flags1>" "<flags2>" ...
# Every variant is built into build/variants/ and selected with SPHB_LIB; the default libsphb200.so is never touched.
set -u
cmd=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do cmd+=("$1"); shift; done
shift
mkdir -p build/variants
i=0
for flags in "$@"; do
  i=$((i+1))
  out=build/variants/v$i.so
  python python-sph_b200/build_native.py --out $out $flags > /dev/null 2>&1 || { echo "BUILD FAILED: $flags"; continue; }
  echo "=== $flags"
  SPHB_LIB=$PWD/$out python "${cmd[@]}" 2>&1 | tail -3
done
