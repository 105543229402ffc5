#!/bin/bash
# One gpurun call (1 GPU): EEDI2 after the sorting-network change (parity + throughput), initcheck of the multi-device path,
# wall time of the default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checke
mkdir -p $OUT
timeout 900 python -m pytest tests/test_decomb_gpu.py tests/test_device_chain_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > $OUT/pytest_eedi2.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_eedi2.log)" | tee $OUT/summary.txt
timeout 600 python tools/bench_filters.py --only 4k10_decomb_eedi2bob --frames 48 --cpu-frames 0 > $OUT/bench_eedi2.jsonl 2> $OUT/bench_eedi2.err
echo "eedi2 bench rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/eedi2_launches.csv python tools/bench_filters.py --only 4k10_decomb_eedi2bob --frames 3 --cpu-frames 0 > $OUT/eedi2_launches.log 2>&1
echo "eedi2 launch list rc=$?" | tee -a $OUT/summary.txt
timeout 600 compute-sanitizer --tool initcheck --print-limit 5 python -m pytest tests/test_nlmeans_multi_gpu.py -q -m gpu -x -k "two_handles" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Uninitialized" | head -8 > $OUT/initcheck_multi.txt
echo "initcheck rc=$?" | tee -a $OUT/summary.txt
T0=$(date +%s)
timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$? wall $(( $(date +%s) - T0 )) s" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
