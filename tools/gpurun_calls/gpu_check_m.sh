#!/bin/bash
# One gpurun call (1 GPU): comb-detect mask kernel with warp-shared prev/next rows, lapsharp with exact_i2d (A/B against the conversion instructions), ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkm
mkdir -p $OUT
timeout 600 python -m pytest tests/test_comb_detect_gpu.py tests/test_decomb_gpu.py tests/test_lapsharp_gpu.py tests/test_hqdn3d_gpu.py -m gpu -x -q > $OUT/pytest_a.log 2>&1
echo "pytest comb/decomb/lapsharp/hqdn3d rc=$? $(tail -1 $OUT/pytest_a.log)" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_device_chain_gpu.py -m gpu -q > $OUT/pytest_b.log 2>&1
echo "pytest golden/fullsize/chain rc=$? $(tail -1 $OUT/pytest_b.log)" | tee -a $OUT/summary.txt
HBCU_LAP_VARIANT=7 timeout 300 python -m pytest tests/test_lapsharp_gpu.py -m gpu -x -q > $OUT/pytest_lap7.log 2>&1
for w in 4k10_comb_detect 4k_lapsharp 4k10_decomb_yadif; do
  timeout 300 python tools/bench_filters.py --only $w --frames 64 --cpu-frames 0 >> $OUT/bench.jsonl 2>> $OUT/bench.err
done
HBCU_LAP_VARIANT=7 timeout 300 python tools/bench_filters.py --only 4k_lapsharp --frames 64 --cpu-frames 0 >> $OUT/bench_lap7.jsonl 2>> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --set full --clock-control none -k regex:comb_mask_bits -c 1 -o $OUT/comb_mask python tools/bench_filters.py --only 4k10_comb_detect --frames 8 --cpu-frames 0 > $OUT/ncu_run1.log 2>&1
echo "ncu comb mask rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --set full --clock-control none -k regex:lapsharp_kernel -c 1 -o $OUT/lapsharp python tools/bench_filters.py --only 4k_lapsharp --frames 8 --cpu-frames 0 > $OUT/ncu_run2.log 2>&1
echo "ncu lapsharp rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --set full --clock-control none -k regex:unsharp_kernel -c 1 -o $OUT/unsharp python tools/bench_filters.py --only 4k_unsharp --frames 8 --cpu-frames 0 > $OUT/ncu_run3.log 2>&1
echo "ncu unsharp rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/comb_launches.csv python tools/bench_filters.py --only 4k10_comb_detect --frames 8 --cpu-frames 0 > $OUT/ncu_run4.log 2>&1
cat $OUT/summary.txt
