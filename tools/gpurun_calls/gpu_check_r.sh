#!/bin/bash
# One gpurun call (1 GPU): three-phase comb-detect mask kernel (HBCU_COMB_MASK=2) against the round-1 kernel: parity, throughput, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkr
mkdir -p $OUT
HBCU_COMB_MASK=2 timeout 600 python -m pytest tests/test_comb_detect_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_sharding_gpu.py -m gpu -x -q > $OUT/pytest2.log 2>&1
echo "pytest (mask 2) rc=$? $(tail -1 $OUT/pytest2.log)" | tee $OUT/summary.txt
for m in 1 2 1 2; do
  HBCU_COMB_MASK=$m timeout 300 python tools/bench_filters.py --only 4k10_comb_detect --frames 64 --cpu-frames 0 >> $OUT/bench_mask$m.jsonl 2>> $OUT/bench.err
done
echo "bench rc=$?" | tee -a $OUT/summary.txt
HBCU_COMB_MASK=2 timeout 300 ncu --set full --clock-control none -k regex:comb_mask_bits -c 1 -o $OUT/comb_mask2 python tools/bench_filters.py --only 4k10_comb_detect --frames 8 --cpu-frames 0 > $OUT/ncu_run.log 2>&1
echo "ncu rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
