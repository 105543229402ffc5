#!/bin/bash
# One gpurun call (1 GPU): the multi-device tests (decomb devices= among them), ncu --set full of the 16-bit v3w kernel and of
# the prefilter variant.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkg
mkdir -p $OUT
timeout 600 python -m pytest tests/test_nlmeans_multi_gpu.py tests/test_nlmeans_gpu.py tests/test_decomb_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest.log)" | tee $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlmeans_v3w -s 4 -c 1 -o $OUT/ncu_v3w -f python bench.py --workload 4k10_nlmeans_medium --steps 1 --warmup 3 --batch 2 --no-cpu-baseline --no-extra --no-copy-only > $OUT/ncu_v3w.log 2>&1
echo "ncu v3w rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlmeans_v3_kernel -s 4 -c 1 -o $OUT/ncu_v3pre -f python bench.py --workload 4k_nlmeans_medium_prefilter --steps 1 --warmup 3 --batch 2 --no-cpu-baseline --no-extra --no-copy-only > $OUT/ncu_v3pre.log 2>&1
echo "ncu v3 prefilter rc=$?" | tee -a $OUT/summary.txt
timeout 300 python bench.py --workload 4k10_nlmeans_medium --steps 8 --warmup 3 --batch 256 --no-cpu-baseline --no-extra > $OUT/bench_4k10.json 2> $OUT/bench_4k10.err
echo "bench 4k10 rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
