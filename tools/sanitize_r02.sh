#!/bin/bash
# compute-sanitizer passes over what round 2 added (run on a B200 box):
#   memcheck  -- the v3 8-bit kernel (TMA boxes, tcgen05.ld/st accumulators), its prefilter variant, the 16-bit v3w kernel,
#                the multi-device halo copy path, wrapped external frames, detelecine device copies
#   racecheck -- shared-memory hazards of the v3 kernels (LUT fill vs first use, tile reuse between frames)
#   initcheck -- uninitialised global reads of the halo / prefilter planes (the one-box mismatch seen before the
#                same-device halo copy became a plain stream-ordered copy)
# usage: bash tools/sanitize_r02.sh > gpurun_out/sanitize_r02.txt 2>&1
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
CS="compute-sanitizer --error-exitcode 99 --print-limit 5"
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error|error|Invalid|hazard|Uninitialized" | head -12; echo "exit=${PIPESTATUS[0]}"; }
run $CS --tool memcheck python -m pytest tests/test_nlmeans_gpu.py -q -m gpu -x -k "config1 or ragged or test_10bit or prefilter_modes"
run $CS --tool memcheck python -m pytest tests/test_nlmeans_multi_gpu.py -q -m gpu -x -k "two_handles"
run $CS --tool memcheck python -m pytest tests/test_wrap_gpu.py tests/test_detelecine_gpu.py -q -m gpu -x -k "wrap or device_resident"
run $CS --tool racecheck python -m pytest tests/test_nlmeans_gpu.py -q -m gpu -x -k "ragged or test_10bit"
run $CS --tool initcheck python -m pytest tests/test_nlmeans_multi_gpu.py -q -m gpu -x -k "two_handles"
