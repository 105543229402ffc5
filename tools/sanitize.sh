#!/bin/bash
# compute-sanitizer passes over the kernels added in round 1 (run on a B200 box):
#   memcheck  -- out-of-bounds / misaligned accesses of the vectorised EEDI2 stages, the bit-packed comb-detect path,
#                unsharp, the NLMeans 10-bit kernel + prefilters, device frames
#   racecheck -- shared-memory hazards of block_deal / lattice pass B / the comb filter chain
# usage: bash tools/sanitize.sh > gpurun_out/sanitize.txt 2>&1
set -u
CS="compute-sanitizer --error-exitcode 99 --print-limit 5"
run() { echo "=== $*"; timeout 600 "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error|error|Invalid|hazard" | head -12; echo "exit=${PIPESTATUS[0]}"; }
run $CS --tool memcheck python -m pytest tests/test_device_chain_gpu.py -q -m gpu -x -k "full_chain or eedi2_bob or padded_stride"
run $CS --tool memcheck python -m pytest tests/test_comb_detect_gpu.py -q -m gpu -x -k "masks_match"
run $CS --tool memcheck python -m pytest tests/test_unsharp_gpu.py -q -m gpu -x -k "extreme"
run $CS --tool memcheck python -m pytest tests/test_nlmeans_gpu.py -q -m gpu -x -k "out_of_range or 10bit_extreme or (prefilter_modes and 32)"
run $CS --tool racecheck python -m pytest tests/test_device_chain_gpu.py -q -m gpu -x -k "eedi2_bob"
run $CS --tool racecheck python -m pytest tests/test_comb_detect_gpu.py -q -m gpu -x -k "short_clips"
