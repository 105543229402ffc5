#!/bin/bash
# One gpurun call (1 GPU): unsharp / chroma smooth with the chroma planes on forked streams, the added hqdn3d geometries.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checko
mkdir -p $OUT
timeout 600 python -m pytest tests/test_unsharp_gpu.py tests/test_hqdn3d_gpu.py tests/test_device_chain_gpu.py tests/test_golden_gpu.py tests/test_nlmeans_multi_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest.log)" | tee $OUT/summary.txt
for w in 4k_unsharp 4k_chroma_smooth; do
  timeout 300 python tools/bench_filters.py --only $w --frames 64 --cpu-frames 0 >> $OUT/bench.jsonl 2>> $OUT/bench.err
done
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_hqdn3d_gpu.py -m gpu -x -q -k "extreme or round1" > $OUT/racecheck_hqdn3d.log 2>&1
echo "racecheck hqdn3d rc=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' $OUT/racecheck_hqdn3d.log | tail -2 | tr '\n' ' ')" | tee -a $OUT/summary.txt
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_hqdn3d_gpu.py tests/test_lapsharp_gpu.py -m gpu -x -q -k "extreme or round1 or lapsharp" > $OUT/memcheck.log 2>&1
echo "memcheck hqdn3d+lapsharp rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/memcheck.log | tail -2 | tr '\n' ' ')" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
