#!/bin/bash
# compute-sanitizer memcheck over the kernels added last in round 1 (detelecine metrics / reductions, EEDI2 postproc 2/3);
# sized for a few seconds of box time.  usage: bash tools/sanitize_r01o.sh > gpurun_out/sanitize_r01o.txt 2>&1
set -u
CS="compute-sanitizer --error-exitcode 99 --print-limit 5 --tool memcheck"
run() { echo "=== $*"; "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Error|error" | head -8; echo "exit=${PIPESTATUS[0]}"; }
run timeout ${1:-16} $CS python -m pytest tests/test_detelecine_gpu.py -q -m gpu -x -p no:cacheprovider -k "matches_reference and 10-328-122"
run timeout ${2:-16} $CS python -m pytest tests/test_decomb_gpu.py -q -m gpu -x -p no:cacheprovider -k "postproc_matches_port and 10-208-120-24-3"
