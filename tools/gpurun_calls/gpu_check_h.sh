#!/bin/bash
# One gpurun call (1 GPU): unsharp / chroma smooth after the round-2 kernel (parity + throughput), the device chains that use them.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkh
mkdir -p $OUT
timeout 600 python -m pytest tests/test_unsharp_gpu.py tests/test_device_chain_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_nlmeans_multi_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest.log)" | tee $OUT/summary.txt
for w in 4k_unsharp 4k_chroma_smooth; do
  timeout 300 python tools/bench_filters.py --only $w --frames 64 --cpu-frames 0 >> $OUT/bench_unsharp.jsonl 2>> $OUT/bench_unsharp.err
done
echo "unsharp bench rc=$?" | tee -a $OUT/summary.txt
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $OUT/smoke.log)" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
