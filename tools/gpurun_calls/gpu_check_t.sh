#!/bin/bash
# One short gpurun call: memcheck over the comb-detect tests (three-phase mask kernel) and the unsharp tests (forked chroma streams).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkt
mkdir -p $OUT
timeout 100 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_comb_detect_gpu.py -m gpu -x -q > $OUT/memcheck_comb.log 2>&1
echo "memcheck comb-detect rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/memcheck_comb.log | tail -2 | tr '\n' ' ')" | tee $OUT/summary.txt
