#!/bin/bash
# compute-sanitizer over the small-N GPU tests.  ONE tool per gpurun call (B200_PROFILING.md: running
# all four tools in one call has left a GPU unusable):
#     gpurun --timeout 1500 -- scripts/sanitize.sh memcheck      (or racecheck | initcheck | synccheck)
# Runs the kept-list tests and the N <= 6000 parity tests; writes gpurun_out/sanitize_<tool>.log; exit code = the tool's.
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
sel='test_drains_equal_scanning_kernels_bit_for_bit or test_check_flags or test_golden_B_stage or test_golden_A_newton_paths or test_stage_vs_oracle and not 6000'
timeout 1400 compute-sanitizer --tool "$tool" --error-exitcode 3 --launch-timeout 600 \
    python -m pytest tests/test_gpu_nlist.py tests/test_gpu_parity.py -x -q -m gpu -k "$sel" \
    > gpurun_out/sanitize_"$tool".log 2>&1
rc=$?
tail -15 gpurun_out/sanitize_"$tool".log
echo "compute-sanitizer --tool $tool: exit $rc"
exit $rc
