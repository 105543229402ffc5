#!/bin/bash
# One gpurun call (1 GPU): comb-detect with the three-phase mask kernel as default: whole comb/decomb parity, both kernels on interlaced and progressive content.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checks
mkdir -p $OUT
timeout 600 python -m pytest tests/test_comb_detect_gpu.py tests/test_decomb_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_sharding_gpu.py tests/test_device_chain_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest (default) rc=$? $(tail -1 $OUT/pytest.log)" | tee $OUT/summary.txt
HBCU_COMB_MASK=1 timeout 600 python -m pytest tests/test_comb_detect_gpu.py -m gpu -x -q > $OUT/pytest1.log 2>&1
echo "pytest (mask 1) rc=$? $(tail -1 $OUT/pytest1.log)" | tee -a $OUT/summary.txt
for prog in 0 1; do for m in 1 2; do
  HBCU_BENCH_COMB_PROGRESSIVE=$prog HBCU_COMB_MASK=$m timeout 300 python tools/bench_filters.py --only 4k10_comb_detect --frames 64 --cpu-frames 0 >> $OUT/bench_prog${prog}_mask$m.jsonl 2>> $OUT/bench.err
done; done
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/comb_launches.csv python tools/bench_filters.py --only 4k10_comb_detect --frames 8 --cpu-frames 0 > $OUT/ncu_run.log 2>&1
cat $OUT/summary.txt
