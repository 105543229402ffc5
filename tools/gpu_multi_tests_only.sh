#!/bin/bash
# One gpurun --gpus 2 call: the tests that need two real devices (devices=0,1 dealing of NLMeans, decomb, the frame-parallel filters).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/multi2t
mkdir -p $OUT
timeout 600 python -m pytest tests/test_nlmeans_multi_gpu.py tests/test_sharding_gpu.py tests/test_decomb_gpu.py tests/test_unsharp_gpu.py -m gpu -x -q > $OUT/pytest_multi.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_multi.log)" | tee $OUT/summary.txt
grep -c "SKIP\|skipped" $OUT/pytest_multi.log | tee -a $OUT/summary.txt
