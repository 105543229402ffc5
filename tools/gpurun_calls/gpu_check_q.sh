#!/bin/bash
# One gpurun call (1 GPU): detelecine with the deferred copy-out (parity, goldens, chains, throughput).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkq
mkdir -p $OUT
timeout 600 python -m pytest tests/test_detelecine_gpu.py tests/test_golden_gpu.py tests/test_device_chain_gpu.py tests/test_abi.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest.log)" | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_abi.py -q > $OUT/pytest_abi.log 2>&1
echo "pytest abi rc=$? $(tail -1 $OUT/pytest_abi.log)" | tee -a $OUT/summary.txt
for i in 1 2; do
  timeout 300 python tools/bench_filters.py --only 4k_detelecine --frames 96 --cpu-frames 0 >> $OUT/bench.jsonl 2>> $OUT/bench.err
done
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_detelecine_gpu.py -m gpu -x -q > $OUT/memcheck.log 2>&1
echo "memcheck detelecine rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/memcheck.log | tail -2 | tr '\n' ' ')" | tee -a $OUT/summary.txt
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $OUT/smoke.log)" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
